// k_deskew.cu — rotation detection and interpolated resampling.
// CPU semantics: reference imageprocess/deskew.c:48-290 and
// imageprocess/interpolate.c:13-129 (including its quirks: bicubic truncates
// toward zero, bilinear uses the wrong fractional coordinate in its two
// degenerate branches).  Compiled with --fmad=false so that every float
// expression rounds exactly like the C code (reference meson.build:243).
// Replaces backend_cuda_deskew.c + cuda_kernels_deskew.cu:13-200 and the
// cv::cuda::warpAffine path (opencv_ops.cpp:552).
#include "common.cuh"
#include "launch.h"

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

#define ROT_CHUNK 32
#define ROT_ROWS 16
#define ROT_TILE 8

struct RotParams {
  int scan_size;        // params.deskewScanSize (may be -1)
  float scan_depth;     // params.deskewScanDepth
  int nangles;
  int edges[4];         // left, top, right, bottom enabled
  int peak_off;         // u32 offset: peaks[mask][edge][angle]
};

// detect_edge_rotation_peak (deskew.c:48-142) for one (page, mask, edge, angle)
__global__ void k_rot_peaks(DPage *pages, const float *tan_tab, RotParams rp, int mi0) {
  extern __shared__ int2 pts[];
  __shared__ int s_red[8][ROT_CHUNK];
  __shared__ int s_done, s_max, s_dep;
  DPage &pg = pages[blockIdx.z];
  int a = blockIdx.x, mi = mi0 + (blockIdx.y >> 2), e = blockIdx.y & 3;
  if (mi >= pg.mask_count || !rp.edges[e]) return;
  const DImg &im = pg.img;
  DRect mask = pg.masks[mi];
  float m = tan_tab[a];
  int sw = abs(mask.x0 - mask.x1) + 1, sh = abs(mask.y0 - mask.y1) + 1;
  // edge -> shift: left = rightward, top = downward, right = leftward, bottom = upward
  int shx = e == 0 ? 1 : e == 2 ? -1 : 0;
  int shy = e == 1 ? 1 : e == 3 ? -1 : 0;
  int scan = rp.scan_size, maxDepth;
  float X, Y, stepX, stepY;
  if (shy == 0) {
    if (scan == -1) scan = sh;
    scan = min(scan, min(10000, sh));
    maxDepth = sw / 2;
    int half = scan / 2;
    int outer = (int)(fabsf(m) * half);
    int mid = sh / 2;
    int side = shx > 0 ? mask.x0 - outer : mask.x1 + outer;
    X = side + half * m;
    Y = mask.y0 + mid - half;
    stepX = -m; stepY = 1.0f;
  } else {
    if (scan == -1) scan = sw;
    scan = min(scan, min(10000, sw));
    maxDepth = sh / 2;
    int half = scan / 2;
    int outer = (int)(fabsf(m) * half);
    int mid = sw / 2;
    int side = shy > 0 ? mask.x0 - outer : mask.x1 + outer;   // deskew.c:96-97 uses .x here
    X = mask.x0 + mid - half;
    Y = side - (half * m);
    stepX = 1.0f; stepY = -m;
  }
  int maxAbs = (int)(255 * rp.scan_size * rp.scan_depth);   // deskew.c:67
  // deskew.c:108-113: sequential float accumulation; X and Y are independent
  // chains, so two threads (of different warps) run one each
  if (threadIdx.x == 0) {
    for (int k = 0; k < scan; k++) { pts[k].x = (int)X; X += stepX; }
    s_done = 0; s_max = 0; s_dep = 0;
  } else if (threadIdx.x == 32) {
    for (int k = 0; k < scan; k++) { pts[k].y = (int)Y; Y += stepY; }
  }
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  int last = 0, maxDiff = 0, acc = 0, dep = 0;   // thread 0's copies are authoritative
  int mx0 = min(mask.x0, mask.x1), mx1 = max(mask.x0, mask.x1), my0 = min(mask.y0, mask.y1), my1 = max(mask.y0, mask.y1);
  // a sample only counts inside the mask AND inside the image (outside reads white = 0)
  int vx0 = max(mx0, 0), vx1 = min(mx1, im.w - 1), vy0 = max(my0, 0), vy1 = min(my1, im.h - 1);
  bool gray8 = im.fmt == DF_GRAY8;
  for (int base = 0; base < maxDepth; base += ROT_CHUNK) {
    // lane = depth (base + lane): with a horizontal shift the 32 lanes read 32
    // consecutive bytes of one row per sample point
    int d = base + lane;
    int part = 0;
    int ox = d * shx, oy = d * shy;
    for (int k = warp; k < scan; k += nwarp) {
      int2 p = pts[k];
      int x = p.x + ox, y = p.y + oy;
      if (x >= vx0 && x <= vx1 && y >= vy0 && y <= vy1) {
        int v = gray8 ? (int)im.data[(size_t)y * im.pitch + x] : px_darkinv(px_load(im, x, y));
        part += 255 - v;
      }
    }
    s_red[warp][lane] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int dd = 0; dd < ROT_CHUNK; dd++) {
        if (!((acc < maxAbs) && (dep < maxDepth))) { s_done = 1; break; }   // deskew.c:119
        int blackness = 0;
        for (int w = 0; w < nwarp; w++) blackness += s_red[w][dd];
        int diff = blackness - last;
        last = blackness;
        if (diff >= maxDiff) maxDiff = diff;
        acc += blackness;
        dep++;
      }
      if (!((acc < maxAbs) && (dep < maxDepth))) s_done = 1;
      s_max = maxDiff; s_dep = dep;
    }
    __syncthreads();
    if (s_done) break;
  }
  if (threadIdx.x == 0) {
    int peak = (s_dep < maxDepth) ? s_max : 0;   // deskew.c:137-141
    pg.u32[rp.peak_off + ((size_t)mi * 4 + e) * rp.nangles + a] = (unsigned)peak;
  }
}


/* --------------------------------------------------------------------------
 * Horizontal-edge fast path (left / right edges, the default).
 *
 * For these edges the scan line visits consecutive rows Y0, Y0+1, ... and its
 * x coordinate (int)X changes only every ~1/|tan| rows, i.e. the line is a
 * staircase of vertical runs.  With a column prefix sum
 *     C[r][x] = sum_{r' < r} (255 - max channel)(x, Y0 + r')      (rows outside
 *     mask /\ image contribute 0)
 * the blackness of the line at depth d is a sum over RUNS of
 *     C[k_end][x_run + d] - C[k_start][x_run + d]
 * — the same integer as the reference's per-sample sum (deskew.c:121-129), with
 * ~scan*|tan| loads instead of `scan`.  Lanes own consecutive depths, so a warp
 * reads 32 consecutive prefix entries per run.
 * ------------------------------------------------------------------------ */
#define RP_SEG 32
#define ROT_DEPTH_CAP 256   // columns from each scanned edge covered by phase 1 of the prefix table

__device__ __forceinline__ void rot_line_geometry(const DRect &mask, int scan_param, int &scan, int &Y0) {
  int sh = abs(mask.y0 - mask.y1) + 1;
  scan = scan_param;
  if (scan == -1) scan = sh;
  scan = min(scan, min(10000, sh));
  int half = scan / 2, mid = sh / 2;
  Y0 = mask.y0 + mid - half;
}

// Two phases (the blackness of a scan line is only needed until the running total reaches
// deskew.c:67's limit, a few dozen columns past the first ink): phase 1 builds the table for the
// columns within `depth_cap` of the two scanned edges only; a page on which some scan line did not
// stop within that depth is flagged (DPage.rot_more) and redone over the full width in phase 2,
// whose kernels return at once for every other page.
__global__ void __launch_bounds__(32 * RP_SEG, 2) k_rot_colprefix(DPage *pages, int mi, int scan_param, int depth_cap, int outer_max, int phase) {
  // block = 32 lanes x RP_SEG row segments; a lane owns 4 adjacent columns (one
  // 32-bit load per row, one 16-byte store of four running sums)
  __shared__ uint4 tot[RP_SEG][32];
  DPage &pg = pages[blockIdx.y];
  if (mi >= pg.mask_count) return;
  if (phase == 2 && !pg.rot_more[mi]) return;
  const DImg &im = pg.img;
  DRect mask = pg.masks[mi];
  int scan, Y0;
  rot_line_geometry(mask, scan_param, scan, Y0);
  if (scan <= 0) return;
  if ((long long)(scan + 1) * im.w > pg.pre_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&pg.error, DERR_UNSUPPORTED); return; }
  int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  // columns the peak kernels can read: run x in [x0 - 2 outer, x0] + depth (left edge), mirrored for the right edge
  int need_l1, need_r0;
  {
    int mx0 = min(mask.x0, mask.x1), mx1 = max(mask.x0, mask.x1);
    bool limited = phase == 1 && depth_cap < (mx1 - mx0 + 1) / 2;
    need_l1 = limited ? mx0 + depth_cap + 1 : 0x7fffffff;
    need_r0 = limited ? mx1 - depth_cap - 1 : -0x7fffffff;
    (void)outer_max;
    int bx0 = blockIdx.x * 128, bx1 = bx0 + 127;
    if (bx0 > need_l1 && bx1 < need_r0) return;     // whole block outside both strips (block-uniform)
  }
  int x = (blockIdx.x * 32 + lane) * 4;
  int my0 = min(mask.y0, mask.y1), my1 = max(mask.y0, mask.y1);
  int vy0 = max(my0, 0), vy1 = min(my1, im.h - 1);
  int per = (scan + RP_SEG - 1) / RP_SEG;
  int r0 = seg * per, r1 = min(r0 + per, scan);
  bool fast = im.fmt == DF_GRAY8 && (im.pitch & 3) == 0 && ((uintptr_t)im.data & 3) == 0 && (im.w & 3) == 0 &&
              ((uintptr_t)pg.pre & 15) == 0;
  int ncol = min(4, im.w - x);          // columns this lane really owns (<= 0: none)
  {   // only the mask's columns are ever read back (k_rot_peaks_*: vx0..vx1)
    int mx0 = min(mask.x0, mask.x1), mx1 = max(mask.x0, mask.x1);
    if (x + 3 < max(mx0, 0) || x > min(mx1, im.w - 1)) ncol = 0;
    if (x > need_l1 && x + 3 < need_r0) ncol = 0;
  }
  uint4 t = make_uint4(0, 0, 0, 0);
  if (ncol > 0)
    for (int r = r0; r < r1; r++) {
      int y = Y0 + r;
      if (y < vy0 || y > vy1) continue;
      if (fast) {
        unsigned w = ~*(const unsigned *)(im.data + (size_t)y * im.pitch + x);   // 255 - v per byte
        t.x += w & 0xFFu; t.y += (w >> 8) & 0xFFu; t.z += (w >> 16) & 0xFFu; t.w += w >> 24;
      } else {
        unsigned v[4] = {0, 0, 0, 0};
        for (int k = 0; k < ncol; k++) v[k] = 255u - (unsigned)px_darkinv(px_load(im, x + k, y));
        t.x += v[0]; t.y += v[1]; t.z += v[2]; t.w += v[3];
      }
    }
  tot[seg][lane] = t;
  __syncthreads();
  if (ncol <= 0) return;
  uint4 run = make_uint4(0, 0, 0, 0);
  for (int s = 0; s < seg; s++) { uint4 q = tot[s][lane]; run.x += q.x; run.y += q.y; run.z += q.z; run.w += q.w; }
  unsigned *C = pg.pre;
  for (int r = r0; r <= r1; r++) {
    if (r == r1 && r1 != scan) break;   // row `scan` (the grand total) is written by the last segment only
    if (r0 >= r1) break;
    unsigned *dst = C + (size_t)r * im.w + x;
    if (fast) *(uint4 *)dst = run;
    else { unsigned rv[4] = {run.x, run.y, run.z, run.w}; for (int k = 0; k < ncol; k++) dst[k] = rv[k]; }
    if (r == r1) break;
    int y = Y0 + r;
    if (y < vy0 || y > vy1) continue;
    if (fast) {
      unsigned w = ~*(const unsigned *)(im.data + (size_t)y * im.pitch + x);
      run.x += w & 0xFFu; run.y += (w >> 8) & 0xFFu; run.z += (w >> 16) & 0xFFu; run.w += w >> 24;
    } else {
      unsigned v[4] = {0, 0, 0, 0};
      for (int k = 0; k < ncol; k++) v[k] = 255u - (unsigned)px_darkinv(px_load(im, x + k, y));
      run.x += v[0]; run.y += v[1]; run.z += v[2]; run.w += v[3];
    }
  }
}

__global__ void k_rot_peaks_h(DPage *pages, const float *tan_tab, RotParams rp, int mi) {
  extern __shared__ int2 runs[];   // {x of the run, first k of the run}
  __shared__ int s_red[8][ROT_CHUNK];
  __shared__ int s_done, s_max, s_dep, s_nruns;
  DPage &pg = pages[blockIdx.z];
  int a = blockIdx.x, e = blockIdx.y ? 2 : 0;
  if (mi >= pg.mask_count || !rp.edges[e]) return;
  const DImg &im = pg.img;
  DRect mask = pg.masks[mi];
  float m = tan_tab[a];
  int sw = abs(mask.x0 - mask.x1) + 1;
  int shx = e == 0 ? 1 : -1;
  int scan, Y0;
  rot_line_geometry(mask, rp.scan_size, scan, Y0);
  int maxDepth = sw / 2;
  int half = scan / 2;
  int outer = (int)(fabsf(m) * half);
  int side = shx > 0 ? mask.x0 - outer : mask.x1 + outer;
  float X = side + half * m;          // deskew.c:88
  float stepX = -m;
  int maxAbs = (int)(255 * rp.scan_size * rp.scan_depth);
  if (threadIdx.x == 0) {
    int n = 0, prev = 0;
    for (int k = 0; k < scan; k++) {   // deskew.c:108-113, sequential float accumulation
      int xi = (int)X;
      if (k == 0 || xi != prev) { runs[n++] = make_int2(xi, k); prev = xi; }
      X += stepX;
    }
    s_nruns = n; s_done = 0; s_max = 0; s_dep = 0;
  }
  __syncthreads();
  int nruns = s_nruns;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  int last = 0, maxDiff = 0, acc = 0, dep = 0;
  int mx0 = min(mask.x0, mask.x1), mx1 = max(mask.x0, mask.x1);
  int vx0 = max(mx0, 0), vx1 = min(mx1, im.w - 1);
  const unsigned *C = pg.pre;
  bool have = (long long)(scan + 1) * im.w <= pg.pre_cap && scan > 0;
  for (int base = 0; base < maxDepth; base += ROT_CHUNK) {
    int ox = (base + lane) * shx;
    int part = 0;
    if (have)
      for (int r = warp; r < nruns; r += nwarp) {
        int2 rn = runs[r];
        int kend = (r + 1 < nruns) ? runs[r + 1].y : scan;
        int x = rn.x + ox;
        if (x >= vx0 && x <= vx1) part += (int)(C[(size_t)kend * im.w + x] - C[(size_t)rn.y * im.w + x]);
      }
    s_red[warp][lane] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int dd = 0; dd < ROT_CHUNK; dd++) {
        if (!((acc < maxAbs) && (dep < maxDepth))) { s_done = 1; break; }   // deskew.c:119
        int blackness = 0;
        for (int w = 0; w < nwarp; w++) blackness += s_red[w][dd];
        int diff = blackness - last;
        last = blackness;
        if (diff >= maxDiff) maxDiff = diff;
        acc += blackness;
        dep++;
      }
      if (!((acc < maxAbs) && (dep < maxDepth))) s_done = 1;
      s_max = maxDiff; s_dep = dep;
    }
    __syncthreads();
    if (s_done) break;
  }
  if (threadIdx.x == 0) {
    int peak = (s_dep < maxDepth) ? s_max : 0;
    pg.u32[rp.peak_off + ((size_t)mi * 4 + e) * rp.nangles + a] = (unsigned)peak;
  }
}

// Warp form of the same scan: a CTA of RW_ANG warps takes RW_ANG angles of one
// (page, mask, edge).  Phase A: the first RW_ANG lanes of warp 0 run the
// sequential float chains of those angles side by side and record the runs.
// Phase B: one warp per angle with lane = depth; the early exit of deskew.c:119
// and the max-difference tracking become a warp prefix sum and a ballot, so
// there is no block barrier and no single-thread scan inside the depth loop.
#define RW_ANG 8
__global__ void __launch_bounds__(32 * RW_ANG) k_rot_peaks_w(DPage *pages, const float *tan_tab, RotParams rp, int mi, int rstride, int depth_cap, int phase) {
  extern __shared__ int rsm[];          // [RW_ANG][rstride] run x | [RW_ANG][rstride] run start k
  __shared__ int rn[RW_ANG];
  int *rx = rsm, *rk = rsm + RW_ANG * rstride;
  DPage &pg = pages[blockIdx.z];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int e = blockIdx.y ? 2 : 0;
  if (mi >= pg.mask_count || !rp.edges[e]) return;
  if (phase == 2 && !pg.rot_more[mi]) return;
  const DImg &im = pg.img;
  DRect mask = pg.masks[mi];
  int sw = abs(mask.x0 - mask.x1) + 1;
  int shx = e == 0 ? 1 : -1;
  int scan, Y0;
  rot_line_geometry(mask, rp.scan_size, scan, Y0);
  int maxDepth = sw / 2;
  int maxAbs = (int)(255 * rp.scan_size * rp.scan_depth);
  const int depthLimit = phase == 1 ? min(maxDepth, depth_cap) : maxDepth;   // phase 1 looks at the first depth_cap columns only
  if (threadIdx.x < RW_ANG) {
    int t = threadIdx.x, a = blockIdx.x * RW_ANG + t;
    int n = 0;
    bool overflow = false;
    if (a < rp.nangles) {
      float m = tan_tab[a];
      int half = scan / 2;
      int outer = (int)(fabsf(m) * half);
      int side = shx > 0 ? mask.x0 - outer : mask.x1 + outer;
      float X = side + half * m;          // deskew.c:88
      float stepX = -m;
      int prev = 0;
      for (int k = 0; k < scan; k++) {    // deskew.c:108-113, sequential float accumulation
        int xi = (int)X;
        if (k == 0 || xi != prev) {
          if (n < rstride) { rx[t * rstride + n] = xi; rk[t * rstride + n] = k; }
          else overflow = true;
          n++; prev = xi;
        }
        X += stepX;
      }
    }
    rn[t] = overflow ? -1 : n;
  }
  __syncthreads();
  int a = blockIdx.x * RW_ANG + warp;
  if (a >= rp.nangles) return;
  int nruns = rn[warp];
  if (nruns < 0) { if (lane == 0) atomicOr(&pg.error, DERR_UNSUPPORTED); nruns = 0; }
  const int *jx = rx + warp * rstride, *jk = rk + warp * rstride;
  int mx0 = min(mask.x0, mask.x1), mx1 = max(mask.x0, mask.x1);
  int vx0 = max(mx0, 0), vx1 = min(mx1, im.w - 1);
  const unsigned *C = pg.pre;
  bool have = (long long)(scan + 1) * im.w <= pg.pre_cap && scan > 0;
  int last = 0, maxDiff = 0, acc = 0, dep = 0;
  bool stopped = false;     // the reference's loop ended: total reached, or all depths seen
  for (int base = 0; base < depthLimit; base += 32) {
    int ox = (base + lane) * shx;
    int b = 0;
    if (have)
      for (int r = 0; r < nruns; r++) {
        int x = jx[r] + ox;
        int kend = (r + 1 < nruns) ? jk[r + 1] : scan;
        if (x >= vx0 && x <= vx1) b += (int)(C[(size_t)kend * im.w + x] - C[(size_t)jk[r] * im.w + x]);
      }
    // running total before each depth step (deskew.c:119 tests it before the step)
    int incl = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    bool proc = (acc + incl - b < maxAbs) && (base + lane < depthLimit);
    unsigned pm = __ballot_sync(0xffffffffu, proc);
    int nproc = __popc(pm);            // the processed steps are a prefix: blackness >= 0
    int prevb = __shfl_up_sync(0xffffffffu, b, 1);
    if (lane == 0) prevb = last;
    int diff = proc ? b - prevb : -2147483647 - 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff = max(diff, __shfl_xor_sync(0xffffffffu, diff, o));
    if (nproc > 0) {
      maxDiff = max(maxDiff, diff);    // `if (diff >= maxDiff) maxDiff = diff` from 0
      acc += __shfl_sync(0xffffffffu, incl, nproc - 1);
      last = __shfl_sync(0xffffffffu, b, nproc - 1);
      dep += nproc;
    }
    if (nproc < 32) break;
  }
  stopped = acc >= maxAbs || dep >= maxDepth;     // deskew.c:119 would have left its loop here too
  if (lane == 0) {
    if (phase == 1 && !stopped) atomicOr((unsigned *)&pg.rot_more[mi], 1u);   // did not end within depth_cap: redo in phase 2
    int peak = (dep < maxDepth) ? maxDiff : 0;   // deskew.c:137-141
    pg.u32[rp.peak_off + ((size_t)mi * 4 + e) * rp.nangles + a] = (unsigned)peak;
  }
}

struct RotFinalParams {
  int nangles;
  int edges[4];
  int peak_off;
  float deviation;
  int use_table;         // 1: pair table indexed [vi][vj], v = sign*nangles + angle
  int mi_first, mi_count; // masks to finalize (the engine goes mask by mask, sheet_stages.c:406-413)
};

// detect_edge_rotation's argmax (deskew.c:156-168) + detect_rotation_cpu's
// average / deviation test (deskew.c:218-240)
__global__ void k_rot_finalize(DPage *pages, int npages, const float *rot_tab, const float *pair_tab, RotFinalParams fp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int p = t / D_MAX_MASKS, mi = t % D_MAX_MASKS;
  if (p >= npages) return;
  DPage &pg = pages[p];
  if (mi < fp.mi_first || mi >= fp.mi_first + fp.mi_count) return;
  if (mi == 0) pg.mask_count_deskew = pg.mask_count;
  if (mi >= pg.mask_count) return;
  pg.masks_deskew[mi] = pg.masks[mi];
  float rot[4];
  int vidx[4];
  int count = 0;
  for (int e = 0; e < 4; e++) {
    if (!fp.edges[e]) { pg.rot_angle_idx[mi][e] = -2; continue; }
    const unsigned *pk = pg.u32 + fp.peak_off + ((size_t)mi * 4 + e) * fp.nangles;
    int best = 0, besti = -1;
    for (int a = 0; a < fp.nangles; a++) {
      int v = (int)pk[a];
      if (v > best) { best = v; besti = a; }
    }
    pg.rot_angle_idx[mi][e] = besti;
    float r = besti >= 0 ? rot_tab[besti] : 0.0f;
    bool neg = (e == 1 || e == 3);   // top / bottom results are negated (deskew.c:204,222)
    rot[count] = neg ? -r : r;
    // table coordinate: angle index (no winner = angle 0, rotation 0.0) and sign
    vidx[count] = (neg ? fp.nangles : 0) + (besti >= 0 ? besti : 0);
    count++;
  }
  float result = 0.0f, s = 0.0f, c = 1.0f;
  if (count > 0) {
    if (fp.use_table && count <= 2) {
      int i = vidx[0], j = count == 2 ? vidx[1] : vidx[0];
      const float *en = pair_tab + ((size_t)i * (2 * fp.nangles) + j) * 4 + (count == 2 ? 0 : 0);
      // entry = {rotation(count=2), sin, cos, rotation(count=1 uses i==j which is identical)}
      result = en[0]; s = en[1]; c = en[2];
    } else {
      float total = 0.0f;
      for (int i = 0; i < count; i++) total += rot[i];
      float average = total / count;
      total = 0.0f;
      for (int i = 0; i < count; i++) { float d = rot[i] - average; total += d * d; }
      float deviation = sqrtf(total);
      result = (deviation <= fp.deviation) ? average : 0.0f;
      // no host libm here: double-precision evaluation rounded once (may differ
      // from glibc sinf/cosf by 1 ulp; only reachable with 3-4 scan edges)
      s = (float)sin(-(double)result); c = (float)cos(-(double)result);
    }
  }
  pg.rotation[mi] = result;
  pg.rot_sin[mi] = s; pg.rot_cos[mi] = c;
  pg.rot_apply[mi] = result != 0.0f;
}

// the float tail of detect_rotation_cpu (deskew.c:218-240) evaluated by the host's libm
// (3-4 scan edges: no pair table): tab[page] = {rotation, sinf(-rotation), cosf(-rotation), -}
__global__ void k_rot_set(DPage *pages, int npages, int mi, const float *tab) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  if (mi >= pg.mask_count) return;
  float r = tab[p * 4 + 0];
  pg.rotation[mi] = r; pg.rot_sin[mi] = tab[p * 4 + 1]; pg.rot_cos[mi] = tab[p * 4 + 2];
  pg.rot_apply[mi] = r != 0.0f;
}

// ---- interpolation (interpolate.c:13-129) ----------------------------------
__device__ __forceinline__ int clip_u8(int a) { return a < 0 ? 0 : a > 255 ? 255 : a; }

__device__ __forceinline__ int cubic_scale(float f, int a, int b, int c, int d) {   // interpolate.c:24-32
  int result = (int)(b + 0.5f * f * (c - a + f * (2.0f * a - 5.0f * b + 4.0f * c - d + f * (3.0f * (b - c) + d - a))));
  return clip_u8(result);
}

// exact uint8/small-int -> float without the (slow) conversion pipe: 2^23 + b has b in its mantissa
__device__ __forceinline__ float u8f(unsigned b) { return __int_as_float(0x4B000000u | b) - 8388608.0f; }
// cubic_scale on operands that are already floats holding the same integers:
// (float)(c - a) == fc - fa and (float)(b - c) == fb - fc exactly, so every
// rounding step of interpolate.c:24-32 is reproduced
__device__ __forceinline__ unsigned cubic_scale_f(float f, float a, float b, float c, float d) {
  int result = __float2int_rz(b + 0.5f * f * ((c - a) + f * (2.0f * a - 5.0f * b + 4.0f * c - d + f * (3.0f * (b - c) + d - a))));
  return (unsigned)clip_u8(result);
}
// cubic_scale on four taps packed in one word (a = byte 0 ... d = byte 3).  The three
// integer combinations of interpolate.c:24-32 — c - a, 2a - 5b + 4c - d and
// 3(b - c) + d - a — are exact in float (|v| < 2^24), so any way of forming them gives
// the reference's operands: one byte dot product each, accumulated onto the bit
// pattern of 1.5 * 2^23 so that a single float subtraction finishes the int -> float
// conversion.  The float steps that round (the f-multiplications and the sums with
// them) keep the reference's order.  `hf` is 0.5f * f.
__device__ __forceinline__ float dot_to_float(unsigned taps, int weights) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(weights), "r"(0x4B400000));
  return __int_as_float(d) - 12582912.0f;
}
__device__ __forceinline__ unsigned cubic_scale_w(float f, float hf, unsigned taps) {
  float R = dot_to_float(taps, 0x000100FF);          // c - a
  float P = dot_to_float(taps, (int)0xFF04FB02u);    // 2a - 5b + 4c - d
  float Q = dot_to_float(taps, 0x01FD03FF);          // -a + 3b - 3c + d
  float B = dot_to_float(taps, 0x00000100);          // b
  float v = B + hf * (R + f * (P + f * Q));
  // (int) then av_clip_uint8 without the conversion pipe: clamp, then add 2^23 rounding
  // toward zero, which leaves the truncated value in the low mantissa bits
  v = fminf(fmaxf(v, 0.0f), 255.0f);
  return __float_as_uint(__fadd_rz(v, 8388608.0f)) & 0xFFu;
}
__device__ __forceinline__ int linear_scale(float x, int a, int b) {   // interpolate.c:62-64
  return (int)(uint8_t)(int)((1.0f - x) * a + x * b);
}

template <bool GRAY>
__device__ __forceinline__ Px interp_cubic(const DImg &im, float fx, float fy) {
  int px = (int)fx, py = (int)fy;
  float ffx = fx - px, ffy = fy - py;
  int rr[4], gg[4], bb[4];
#pragma unroll
  for (int i = -1; i < 3; i++) {
    Px q0 = px_get(im, px - 1, py + i), q1 = px_get(im, px, py + i), q2 = px_get(im, px + 1, py + i), q3 = px_get(im, px + 2, py + i);
    rr[i + 1] = cubic_scale(ffx, q0.r, q1.r, q2.r, q3.r);
    if (!GRAY) {
      gg[i + 1] = cubic_scale(ffx, q0.g, q1.g, q2.g, q3.g);
      bb[i + 1] = cubic_scale(ffx, q0.b, q1.b, q2.b, q3.b);
    }
  }
  Px o;
  o.r = cubic_scale(ffy, rr[0], rr[1], rr[2], rr[3]);
  if (GRAY) { o.g = o.r; o.b = o.r; }
  else { o.g = cubic_scale(ffy, gg[0], gg[1], gg[2], gg[3]); o.b = cubic_scale(ffy, bb[0], bb[1], bb[2], bb[3]); }
  return o;
}

__device__ __forceinline__ Px lin_px(float f, Px a, Px b) {
  return Px{linear_scale(f, a.r, b.r), linear_scale(f, a.g, b.g), linear_scale(f, a.b, b.b)};
}

__device__ Px interp_linear(const DImg &im, float fx, float fy) {   // interpolate.c:76-117
  int p1x = (int)floorf(fx), p1y = (int)floorf(fy);
  int p2x = (int)ceil((double)fx), p2y = (int)ceilf(fy);
  if (!in_img(im, p2x, p2y)) return px_get(im, p1x, p1y);
  if (p1x == p2x && p1y == p2y) return px_get(im, p1x, p1y);
  if (p1x == p2x) return lin_px(fx - p1x, px_get(im, p1x, p1y), px_get(im, p2x, p2y));
  if (p1y == p2y) return lin_px(fy - p1y, px_get(im, p1x, p1y), px_get(im, p2x, p2y));
  Px a = px_get(im, p1x, p1y), b = px_get(im, p2x, p1y), c = px_get(im, p1x, p2y), d = px_get(im, p2x, p2y);
  Px h1 = lin_px(fx - p1x, a, b), h2 = lin_px(fx - p1x, c, d);
  return lin_px(fy - p1y, h1, h2);
}

__device__ __forceinline__ Px interp_any(const DImg &im, float fx, float fy, int type, bool gray) {
  if (type == 0) return px_get(im, (int)roundf(fx), (int)roundf(fy));   // interpolate.c:13-18
  if (type == 1) return interp_linear(im, fx, fy);
  return gray ? interp_cubic<true>(im, fx, fy) : interp_cubic<false>(im, fx, fy);
}


/* Ink map: which INK_CELL x INK_CELL cells of the image are pure white.  A target tile whose
 * source footprint (plus the interpolation taps) only touches white cells — or
 * lies outside the image, which reads as white — is white after bicubic
 * interpolation (all 16 taps equal => every term of interpolate.c:24-32 cancels
 * exactly), so rotate() can write it without touching a pixel.  Cells of 8 pixels
 * are fine enough to catch the gaps between text lines. */
#define INK_CELL D_INK_CELL
static_assert(INK_CELL == 8, "k_inkmap reads a cell row as one 8-byte (GRAY8) / 24-byte (RGB24) piece");
__global__ void k_inkmap(DPage *pages) {
  DPage &pg = pages[blockIdx.y];
  const DImg &im = pg.img;
  int ncx = (im.w + INK_CELL - 1) / INK_CELL, ncy = (im.h + INK_CELL - 1) / INK_CELL;
  int bpp = im.fmt == DF_GRAY8 ? 1 : im.fmt == DF_RGB24 ? 3 : 0;
  bool ok = bpp && (im.pitch & 7) == 0 && ((uintptr_t)im.data & 7) == 0 && ncx * ncy <= pg.ink_cap && pg.ink;
  if (blockIdx.x == 0 && threadIdx.x == 0) { pg.ink_ncx = ncx; pg.ink_ncy = ncy; pg.ink_ok = ok; }
  if (!ok) return;
  for (int cy = blockIdx.x; cy < ncy; cy += gridDim.x) {
    int y0 = cy * INK_CELL, y1 = min(y0 + INK_CELL, im.h);
    for (int c = threadIdx.x; c < ncx; c += blockDim.x) {
      bool white = true;
      int x0 = c * INK_CELL;
      if (x0 + INK_CELL <= im.w) {
        for (int y = y0; y < y1 && white; y++) {
          const uint2 *q = (const uint2 *)(im.data + (size_t)y * im.pitch + (size_t)x0 * bpp);
          for (int v = 0; v < bpp; v++) { uint2 t = q[v]; white = white && ((t.x & t.y) == 0xFFFFFFFFu); }
        }
      } else {
        for (int y = y0; y < y1 && white; y++)
          for (int x = x0; x < im.w && white; x++) { Px p = px_load(im, x, y); white = (p.r & p.g & p.b) == 255; }
      }
      pg.ink[cy * ncx + c] = white ? 1 : 0;
    }
  }
}

// warp-collective: true when the source footprint of the target tile
// [xa..xb] x [ya..yb] (mask-local coordinates) is entirely white
__device__ bool rot_tile_white(const DPage &pg, const DImg &im, int xa, int xb, int ya, int yb, float scx, float scy,
                               float tcx, float tcy, float sinval, float cosval, int lane) {
  int cxp = (lane & 1) ? xb : xa, cyp = (lane & 2) ? yb : ya;
  float sx = scx + (cxp - tcx) * cosval + (cyp - tcy) * sinval;
  float sy = scy + (cyp - tcy) * cosval - (cxp - tcx) * sinval;
  float mnx = sx, mxx = sx, mny = sy, mxy = sy;
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  mnx = __shfl_sync(0xffffffffu, mnx, 0); mxx = __shfl_sync(0xffffffffu, mxx, 0);
  mny = __shfl_sync(0xffffffffu, mny, 0); mxy = __shfl_sync(0xffffffffu, mxy, 0);
  if (!(fabsf(mnx) < 1e7f && fabsf(mxx) < 1e7f && fabsf(mny) < 1e7f && fabsf(mxy) < 1e7f)) return false;
  // taps reach from (int)src - 1 to (int)src + 2; two pixels of slack for rounding
  int bx0 = (int)floorf(mnx) - 3, bx1 = (int)floorf(mxx) + 4, by0 = (int)floorf(mny) - 3, by1 = (int)floorf(mxy) + 4;
  if (bx1 < 0 || by1 < 0 || bx0 >= im.w || by0 >= im.h) return true;   // entirely outside: reads as white
  int cx0 = max(bx0, 0) / INK_CELL, cx1 = min(bx1, im.w - 1) / INK_CELL;
  int cy0 = max(by0, 0) / INK_CELL, cy1 = min(by1, im.h - 1) / INK_CELL;
  int nx = cx1 - cx0 + 1, n = nx * (cy1 - cy0 + 1);
  if (n > 32) return false;
  bool w = true;
  if (lane < n) w = pg.ink[(cy0 + lane / nx) * pg.ink_ncx + cx0 + lane % nx] != 0;
  return __all_sync(0xffffffffu, w);
}

// rotate() (deskew.c:253-274) of mask `mi` into aux (mask-sized); the copy
// back is a DCopyJob prepared here.
__global__ void k_rotate(DPage *pages, int mi, int interp, DCopyJob *back_jobs) {
  DPage &pg = pages[blockIdx.z];
  bool active = mi < pg.mask_count && pg.rot_apply[mi];
  DRect mask = pg.masks[mi];
  int w = abs(mask.x0 - mask.x1) + 1, h = abs(mask.y0 - mask.y1) + 1;
  DImg aux = pg.aux;
  int bpp = bytes_pp(aux.fmt);
  int pitch = bpp ? ((w * bpp + 15) & ~15) : (((w + 7) / 8 + 15) & ~15);
  if ((long long)pitch * h > (long long)aux.pitch * aux.h) { active = false; if (mi < pg.mask_count && pg.rot_apply[mi] && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicOr(&pg.error, DERR_UNSUPPORTED); }
  aux.w = w; aux.h = h; aux.pitch = pitch;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    DCopyJob j; j.src = aux; j.dst = pg.img; j.area = DRect{0, 0, w - 1, h - 1};
    j.tx = mask.x0; j.ty = mask.y0; j.enabled = active; j.pad = 0;
    back_jobs[blockIdx.z] = j;
  }
  if (!active) return;
  const DImg &im = pg.img;
  // center_of_rectangle (primitives.c:136-145) of the (normalised) mask / target
  int nx0 = min(mask.x0, mask.x1), ny0 = min(mask.y0, mask.y1);
  float scx = nx0 + w / 2.0f, scy = ny0 + h / 2.0f;
  float tcx = 0 + w / 2.0f, tcy = 0 + h / 2.0f;
  float sinval = pg.rot_sin[mi], cosval = pg.rot_cos[mi];
  bool gray = im.fmt != DF_RGB24;
  bool fast = im.fmt == DF_GRAY8 && interp == 2;
  bool pitch4 = ((im.pitch & 3) == 0);
  // a block owns ROT_ROWS consecutive target rows x 128 columns (one block per
  // row would be launch-bound: ~70 k blocks per A4 page); each warp first tries to
  // dispose of its 32 x ROT_TILE tile through the ink map
  bool can_skip = pg.ink_ok && interp == 2 && (im.fmt == DF_GRAY8 || im.fmt == DF_RGB24) && (aux.fmt == im.fmt);
  int lane = threadIdx.x & 31;
  int yblock = blockIdx.y * ROT_ROWS, yend = min(h, yblock + ROT_ROWS);
  for (int xs = blockIdx.x * blockDim.x; xs < w; xs += gridDim.x * blockDim.x) {
  int x = xs + threadIdx.x;
  int xwa = xs + (threadIdx.x & ~31), xwb = min(xwa + 31, w - 1);   // this warp's columns
  for (int yt = yblock; yt < yend; yt += ROT_TILE) {
  int yte = min(yt + ROT_TILE, yend);
  if (can_skip && xwa < w && rot_tile_white(pg, im, xwa, xwb, yt, yte - 1, scx, scy, tcx, tcy, sinval, cosval, lane)) {
    if (x < w) {
      int nb = aux.fmt == DF_RGB24 ? 3 : 1;
      for (int y = yt; y < yte; y++) {
        uint8_t *o = aux.data + (size_t)y * aux.pitch + (size_t)x * nb;
        o[0] = 255; if (nb == 3) { o[1] = 255; o[2] = 255; }
      }
    }
    continue;
  }
  if (x >= w) continue;
  for (int y = yt; y < yte; y++) {
  uint8_t *orow = aux.data + (size_t)y * aux.pitch;
  {
    float xf = u8f((unsigned)x), yf = u8f((unsigned)y);   // == (float)x, (float)y for 0 <= v < 2^23
    float srcX = scx + (xf - tcx) * cosval + (yf - tcy) * sinval;
    float srcY = scy + (yf - tcy) * cosval - (xf - tcx) * sinval;
    if (fast) {
      // (int)srcX for 1 <= srcX < 2^23 by the 2^23 trick (anything else fails the
      // range test below and takes the general path): no conversion instructions
      float tX = __fadd_rz(srcX, 8388608.0f), tY = __fadd_rz(srcY, 8388608.0f);
      int px = __float_as_int(tX) - 0x4B000000, py = __float_as_int(tY) - 0x4B000000;
      if ((unsigned)(px - 1) < (unsigned)(im.w - 3) && (unsigned)(py - 1) < (unsigned)(im.h - 3)) {
        // the 4x4 taps as four packed words (two aligned loads + funnel shift per row)
        const uint8_t *p0 = im.data + (size_t)(py - 1) * im.pitch + (px - 1);
        unsigned rw[4];
        if (pitch4) {   // rows share their alignment
          const unsigned *wp = (const unsigned *)((uintptr_t)p0 & ~(uintptr_t)3);
          unsigned sh = ((unsigned)(uintptr_t)p0 & 3u) * 8u;
          int wpitch = im.pitch >> 2;
#pragma unroll
          for (int i = 0; i < 4; i++) rw[i] = __funnelshift_r(wp[i * wpitch], wp[i * wpitch + 1], sh);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint8_t *pr = p0 + (size_t)i * im.pitch;
            const unsigned *wp = (const unsigned *)((uintptr_t)pr & ~(uintptr_t)3);
            rw[i] = __funnelshift_r(wp[0], wp[1], ((unsigned)(uintptr_t)pr & 3u) * 8u);
          }
        }
        unsigned o = rw[0] & 0xFFu;
        // all 16 taps equal: every cubic term cancels exactly and the result is that value
        if (!(rw[0] == o * 0x01010101u && rw[1] == rw[0] && rw[2] == rw[0] && rw[3] == rw[0])) {
          float fx = srcX - (tX - 8388608.0f), fy = srcY - (tY - 8388608.0f);   // srcX - (float)px
          float hfx = 0.5f * fx;
          unsigned r4[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            unsigned b0 = rw[i] & 0xFFu;
            // a row of four equal taps interpolates to that value exactly; a row equal
            // to the one above (vertical strokes) has the same result as that row
            if (rw[i] == b0 * 0x01010101u) r4[i] = b0;
            else if (i > 0 && rw[i] == rw[i - 1]) r4[i] = r4[i - 1];
            else r4[i] = cubic_scale_w(fx, hfx, rw[i]);
          }
          // four equal row results: the vertical pass returns that value exactly
          if (r4[1] == r4[0] && r4[2] == r4[0] && r4[3] == r4[0]) o = r4[0];
          else {
            unsigned lo = __byte_perm(r4[0], r4[1], 0x0040), hi = __byte_perm(r4[2], r4[3], 0x0040);
            o = cubic_scale_w(fy, 0.5f * fy, __byte_perm(lo, hi, 0x5410));
          }
        }
        orow[x] = (uint8_t)o;
        continue;
      }
    }
    Px o = interp_any(im, srcX, srcY, interp, gray);
    px_store(aux, x, y, o.r, o.g, o.b);
  }
  }
  }
  }
}

// ---- packed fp32x2 arithmetic (sm_100a FMUL2 / FADD2): two IEEE single operations per issue
// slot, each rounded exactly like the scalar instruction.  Separate mul and add instructions
// with explicit .rn: nothing here may be contracted into an FMA (checked in the SASS: no FFMA2).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(unsigned lo, unsigned hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, unsigned &lo, unsigned &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, which would
// round once instead of twice.  The product therefore passes through an integer XOR with a
// run-time zero (kernel argument): one LOP3 per packed multiplication keeps the two roundings.
__device__ __forceinline__ u64 fmul2(u64 a, u64 b, unsigned zero) {
  u64 r;
  asm volatile("{\n\t.reg .b32 lo, hi;\n\tmul.rn.f32x2 %0, %1, %2;\n\tmov.b64 {lo, hi}, %0;\n\txor.b32 lo, lo, %3;\n\tmov.b64 %0, {lo, hi};\n\t}"
               : "=l"(r) : "l"(a), "l"(b), "r"(zero));
  return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fadd2_rz(u64 a, u64 b) { u64 r; asm volatile("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned dp4(unsigned taps, int weights) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(weights), "r"(0x4B400000));
  return (unsigned)d;
}
// cubic_scale_w for two independent tap words at once: (f.lo, ta) and (f.hi, tb)
__device__ __forceinline__ void cubic_scale_w2(u64 f2, u64 hf2, unsigned ta, unsigned tb, unsigned &oa, unsigned &ob, unsigned zero) {
  const u64 M = pk2(0xCB400000u, 0xCB400000u);            // -1.5 * 2^23 twice
  u64 R = fadd2(pk2(dp4(ta, 0x000100FF), dp4(tb, 0x000100FF)), M);                    // c - a
  u64 P = fadd2(pk2(dp4(ta, (int)0xFF04FB02u), dp4(tb, (int)0xFF04FB02u)), M);        // 2a - 5b + 4c - d
  u64 Q = fadd2(pk2(dp4(ta, 0x01FD03FF), dp4(tb, 0x01FD03FF)), M);                    // -a + 3b - 3c + d
  u64 B = fadd2(pk2(dp4(ta, 0x00000100), dp4(tb, 0x00000100)), M);                    // b
  u64 v = fadd2(B, fmul2(hf2, fadd2(R, fmul2(f2, fadd2(P, fmul2(f2, Q, zero)), zero)), zero));
  unsigned va, vb;
  upk2(v, va, vb);
  float ca = fminf(fmaxf(__uint_as_float(va), 0.0f), 255.0f), cb = fminf(fmaxf(__uint_as_float(vb), 0.0f), 255.0f);
  u64 t = fadd2_rz(pk2(__float_as_uint(ca), __float_as_uint(cb)), pk2(0x4B000000u, 0x4B000000u));   // + 2^23, toward zero
  upk2(t, va, vb);
  oa = va & 0xFFu; ob = vb & 0xFFu;
}

// cubic_scale_w2 with the two results clipped to 0..255 in the two 16-bit halves of one word.
//  * each product is fma(a, b, -0.0) == mul.rn(a, b) (same single rounding, same signed zeros): an FMA
//    cannot be contracted with the addition that follows, so no integer fence is needed.  `nz2` holds
//    -0.0 twice and comes from a kernel argument so that ptxas cannot fold the addend away;
//  * (int)v then av_clip_uint8: v + 1.5 * 2^23 rounded toward zero leaves floor(v) as a two's complement
//    number in the low mantissa bits (floor(v) == (int)v for v >= 0; below zero both clip to 0), and one
//    packed min-with-relu clips both halves.
__device__ __forceinline__ u64 fmul2z(u64 a, u64 b, u64 nz2) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz2)); return r; }
__device__ __forceinline__ unsigned cubic_scale_w2p(u64 f2, u64 hf2, unsigned ta, unsigned tb, u64 nz2) {
  const u64 M = pk2(0xCB400000u, 0xCB400000u);            // -1.5 * 2^23 twice
  u64 R = fadd2(pk2(dp4(ta, 0x000100FF), dp4(tb, 0x000100FF)), M);                    // c - a
  u64 P = fadd2(pk2(dp4(ta, (int)0xFF04FB02u), dp4(tb, (int)0xFF04FB02u)), M);        // 2a - 5b + 4c - d
  u64 Q = fadd2(pk2(dp4(ta, 0x01FD03FF), dp4(tb, 0x01FD03FF)), M);                    // -a + 3b - 3c + d
  u64 B = fadd2(pk2(__byte_perm(ta, 0x4B400000u, 0x7651), __byte_perm(tb, 0x4B400000u, 0x7651)), M);   // b (byte 1 under the bias)
  u64 v = fadd2(B, fmul2z(hf2, fadd2(R, fmul2z(f2, fadd2(P, fmul2z(f2, Q, nz2)), nz2)), nz2));
  u64 t = fadd2_rz(v, pk2(0x4B400000u, 0x4B400000u));
  unsigned lo, hi;
  upk2(t, lo, hi);
  return __vimin_s16x2_relu(__byte_perm(lo, hi, 0x5410), 0x00FF00FFu);
}

// Sheet-engine form of deskew() (deskew.c:276-290) for mask `mi` of every page: ONE sweep
// over the whole sheet from the working buffer into the slot's other buffer,
//     dst(X,Y) = (X,Y) inside the pasted rectangle [mask.vertex[0], + size) ? rotate(...) : src(X,Y)
// so there is no mask-sized temporary and no copy back (4 M bytes -> M + S).  Tiles that do
// not touch the rectangle (and whole pages whose detected rotation is 0: the reference skips
// deskew() for them, sheet_stages.c:410) are copied 16 bytes per thread.
// GRAY8 + cubic fast path: the cubic arithmetic is evaluated unconditionally (the "equal
// taps" identities are exact, so computing the formula gives the same value) — the only
// branch is warp-uniform: a warp whose 32 pixels all sit on constant 4x4 neighbourhoods
// skips the arithmetic.  Divergent per-row shortcuts cost more issue slots than they save.
// true when the source footprint (plus the interpolation taps) of the target tile
// [xa..xb] x [ya..yb] (coordinates of the rotated image) only touches pure-white ink cells or lies
// outside the image.  Warp-collective.  The footprint is the bounding box of the four
// rotated corners, with two pixels of slack for float rounding.
__device__ __forceinline__ bool tile_white(const uint8_t *ink, int ink_ncx, int W, int H, int xa, int xb, int ya, int yb,
                                           float scx, float scy, float tcx, float tcy, float sinval, float cosval, int lane) {
  float ax = (xa - tcx) * cosval, bx = (xb - tcx) * cosval, ay = (ya - tcy) * sinval, by = (yb - tcy) * sinval;
  float cy = (ya - tcy) * cosval, dy = (yb - tcy) * cosval, cx = (xa - tcx) * sinval, dx = (xb - tcx) * sinval;
  float mnx = scx + fminf(ax, bx) + fminf(ay, by), mxx = scx + fmaxf(ax, bx) + fmaxf(ay, by);
  float mny = scy + fminf(cy, dy) - fmaxf(cx, dx), mxy = scy + fmaxf(cy, dy) - fminf(cx, dx);
  if (!(fabsf(mnx) < 1e7f && fabsf(mxx) < 1e7f && fabsf(mny) < 1e7f && fabsf(mxy) < 1e7f)) return false;
  // taps reach from (int)src - 1 to (int)src + 2
  int bx0 = (int)floorf(mnx) - 3, bx1 = (int)floorf(mxx) + 4, by0 = (int)floorf(mny) - 3, by1 = (int)floorf(mxy) + 4;
  if (bx1 < 0 || by1 < 0 || bx0 >= W || by0 >= H) return true;   // entirely outside: reads as white
  int cx0 = max(bx0, 0) / INK_CELL, cx1 = min(bx1, W - 1) / INK_CELL;
  int cy0 = max(by0, 0) / INK_CELL, cy1 = min(by1, H - 1) / INK_CELL;
  // lanes form an 8 x 4 grid of cells (a 32 x 8 tile rotated by a few degrees touches 6-7 x 3-4 cells)
  int nx = cx1 - cx0 + 1, ny = cy1 - cy0 + 1;
  if (nx > 8 || ny > 4) return false;
  int lx = lane & 7, ly = lane >> 3;
  bool wh = true;
  if (lx < nx && ly < ny) wh = ink[(cy0 + ly) * ink_ncx + cx0 + lx] != 0;
  return __all_sync(0xffffffffu, wh);
}

__global__ void __launch_bounds__(128) k_rotate_sheet(DPage *pages, int mi, int interp, unsigned zero) {
  const DPage &pg = pages[blockIdx.z];
  const DImg im = pg.img;                 // by value: no reloads of the descriptor behind the stores below
  DImg out = im;
  out.data = pg.other;
  const int W = im.w, H = im.h, bpp = bytes_pp(im.fmt);
  const bool active = mi < pg.mask_count && pg.rot_apply[mi];
  const DRect mask = pg.masks[mi];
  const int w = abs(mask.x0 - mask.x1) + 1, h = abs(mask.y0 - mask.y1) + 1;
  const int ox = mask.x0, oy = mask.y0;   // copy_rectangle(rotated, source, full, mask.vertex[0]) (deskew.c:283)
  const int X0 = blockIdx.x * 128, Y0 = blockIdx.y * ROT_ROWS;
  if (X0 >= W || Y0 >= H) return;
  const int X1 = min(X0 + 127, W - 1), Y1 = min(Y0 + ROT_ROWS - 1, H - 1);
  const bool touches = active && X1 >= ox && X0 < ox + w && Y1 >= oy && Y0 < oy + h;
  if (!touches) {
    int rows = Y1 - Y0 + 1, rowb = (X1 - X0 + 1) * bpp;
    if (bpp > 0 && (im.pitch & 15) == 0 && (((uintptr_t)im.data | (uintptr_t)out.data) & 15) == 0) {
      int nch = (rowb + 15) >> 4;      // a row's pitch padding may be copied along
      for (int i = threadIdx.x; i < nch * rows; i += blockDim.x) {
        int r = i / nch, c = i - r * nch;
        size_t off = (size_t)(Y0 + r) * im.pitch + (size_t)X0 * bpp + ((size_t)c << 4);
        *(uint4 *)(out.data + off) = *(const uint4 *)(im.data + off);
      }
    } else {
      for (int r = 0; r < rows; r++)
        for (int x = X0 + threadIdx.x; x <= X1; x += blockDim.x) { Px p = px_load(im, x, Y0 + r); px_store(out, x, Y0 + r, p.r, p.g, p.b); }
    }
    return;
  }
  // center_of_rectangle (primitives.c:136-145) of the (normalised) mask / of the target
  const int nx0 = min(mask.x0, mask.x1), ny0 = min(mask.y0, mask.y1);
  const float scx = nx0 + w / 2.0f, scy = ny0 + h / 2.0f;
  const float tcx = 0 + w / 2.0f, tcy = 0 + h / 2.0f;
  const float sinval = pg.rot_sin[mi], cosval = pg.rot_cos[mi];
  const bool gray = im.fmt != DF_RGB24;
  const bool fast = im.fmt == DF_GRAY8 && interp == 2 && (im.pitch & 3) == 0 && ((uintptr_t)im.data & 3) == 0;
  const bool can_skip = pg.ink_ok && interp == 2 && (im.fmt == DF_GRAY8 || im.fmt == DF_RGB24);
  const uint8_t *const ink = pg.ink;
  const int ink_ncx = pg.ink_ncx;
  const int lane = threadIdx.x & 31;
  const int Xw = X0 + (threadIdx.x & ~31);          // this warp's 32 sheet columns
  if (Xw >= W) return;                              // whole warp (warp-uniform)
  const int X = Xw + lane;
  const int x = X - ox;                             // column in the rotated image
  const bool colin = X < W && x >= 0 && x < w;
  const int pitch = im.pitch;
  const uint8_t *const src = im.data;
  // the x-dependent halves of the source coordinates: srcX = (scx + (x - tcx) cos) + (y - tcy) sin,
  // srcY = (scy + (y - tcy) cos) - (x - tcx) sin — evaluated in the reference's order below
  const float xr = u8f((unsigned)(colin ? x : 0)) - tcx;      // (float)x - tcx
  const float xc = xr * cosval, xs = xr * sinval;
  for (int Yt = Y0; Yt <= Y1; Yt += ROT_TILE) {
    const int Yte = min(Yt + ROT_TILE - 1, Y1);
    // the warp's tile entirely inside the pasted rectangle and its source footprint pure white?
    const int Xwe = min(Xw + 31, W - 1);
    const bool tile_in = Xw >= ox && Xwe < ox + w && Yt >= oy && Yte < oy + h;
    if (can_skip && tile_in &&
        tile_white(ink, ink_ncx, W, H, Xw - ox, Xwe - ox, Yt - oy, Yte - oy, scx, scy, tcx, tcy, sinval, cosval, lane)) {
      if (X < W) {
        uint8_t *o = out.data + (size_t)Yt * pitch + (size_t)X * bpp;
        for (int Y = Yt; Y <= Yte; Y++, o += pitch) { o[0] = 255; if (bpp == 3) { o[1] = 255; o[2] = 255; } }
      }
      continue;
    }
    if (fast && tile_in && Yte == Yt + ROT_TILE - 1) {
      // ---- GRAY8 + cubic, every pixel of the 32 x 8 tile inside the rotated image: two rows per step
      uint8_t *orow = out.data + (size_t)Yt * pitch + X;
#pragma unroll 1
      for (int Ya = Yt; Ya < Yt + ROT_TILE; Ya += 2, orow += 2 * (size_t)pitch) {
        float srcX[2], srcY[2], tX[2], tY[2];
        unsigned rw[2][4];
        bool ok[2];
        bool need = false;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const float yr = u8f((unsigned)(Ya + k - oy)) - tcy;        // (float)y - tcy
          srcX[k] = (scx + xc) + yr * sinval;
          srcY[k] = (scy + yr * cosval) - xs;
          // (int)srcX for 1 <= srcX < 2^23 by the 2^23 trick (anything else fails the range test)
          tX[k] = __fadd_rz(srcX[k], 8388608.0f); tY[k] = __fadd_rz(srcY[k], 8388608.0f);
          const int px = __float_as_int(tX[k]) - 0x4B000000, py = __float_as_int(tY[k]) - 0x4B000000;
          ok[k] = (unsigned)(px - 1) < (unsigned)(W - 3) && (unsigned)(py - 1) < (unsigned)(H - 3);
          const int off = ok[k] ? (py - 1) * pitch + (px - 1) : 0;
          const unsigned sh = ((unsigned)off & 3u) * 8u;
          const uint8_t *p0 = src + (off & ~3);
#pragma unroll
          for (int i = 0; i < 4; i++, p0 += pitch) {
            const unsigned *wp = (const unsigned *)p0;
            rw[k][i] = __funnelshift_r(wp[0], wp[1], sh);
          }
          const unsigned b0 = rw[k][0] & 0xFFu;
          const bool uni = rw[k][0] == b0 * 0x01010101u && rw[k][1] == rw[k][0] && rw[k][2] == rw[k][0] && rw[k][3] == rw[k][0];
          need = need || (ok[k] && !uni);
        }
        unsigned o0 = rw[0][0] & 0xFFu, o1 = rw[1][0] & 0xFFu;   // all 16 taps equal: every cubic term cancels exactly
        if (__any_sync(0xffffffffu, need)) {
          // srcX - (float)px; the four rows of a pixel share fx, the two pixels' vertical passes pair up
          float fx0 = srcX[0] - (tX[0] - 8388608.0f), fy0 = srcY[0] - (tY[0] - 8388608.0f);
          float fx1 = srcX[1] - (tX[1] - 8388608.0f), fy1 = srcY[1] - (tY[1] - 8388608.0f);
          u64 f2a = pk2(__float_as_uint(fx0), __float_as_uint(fx0)), h2a = pk2(__float_as_uint(0.5f * fx0), __float_as_uint(0.5f * fx0));
          u64 f2b = pk2(__float_as_uint(fx1), __float_as_uint(fx1)), h2b = pk2(__float_as_uint(0.5f * fx1), __float_as_uint(0.5f * fx1));
          unsigned a0, a1, a2, a3, b0, b1, b2, b3;
          cubic_scale_w2(f2a, h2a, rw[0][0], rw[0][1], a0, a1, zero);
          cubic_scale_w2(f2a, h2a, rw[0][2], rw[0][3], a2, a3, zero);
          cubic_scale_w2(f2b, h2b, rw[1][0], rw[1][1], b0, b1, zero);
          cubic_scale_w2(f2b, h2b, rw[1][2], rw[1][3], b2, b3, zero);
          unsigned ca = __byte_perm(__byte_perm(a0, a1, 0x0040), __byte_perm(a2, a3, 0x0040), 0x5410);
          unsigned cb = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
          u64 fy2 = pk2(__float_as_uint(fy0), __float_as_uint(fy1)), hy2 = pk2(__float_as_uint(0.5f * fy0), __float_as_uint(0.5f * fy1));
          cubic_scale_w2(fy2, hy2, ca, cb, o0, o1, zero);
        }
        if (ok[0] && ok[1]) { orow[0] = (uint8_t)o0; orow[pitch] = (uint8_t)o1; continue; }
        // taps reaching outside the image: the general path (reads outside = white)
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (ok[k]) { orow[(size_t)k * pitch] = (uint8_t)(k ? o1 : o0); continue; }
          Px q = interp_any(im, srcX[k], srcY[k], interp, gray);
          px_store(out, X, Ya + k, q.r, q.g, q.b);
        }
      }
      continue;
    }
    // ---- tiles on the rectangle's edge, other formats / interpolations: pixel by pixel
    for (int Y = Yt; Y <= Yte; Y++) {
      const int y = Y - oy;
      if (colin && y >= 0 && y < h) {
        float xf = u8f((unsigned)x), yf = u8f((unsigned)y);   // == (float)x, (float)y for 0 <= v < 2^23
        float sX = scx + (xf - tcx) * cosval + (yf - tcy) * sinval;
        float sY = scy + (yf - tcy) * cosval - (xf - tcx) * sinval;
        Px q = interp_any(im, sX, sY, interp, gray);
        px_store(out, X, Y, q.r, q.g, q.b);
      } else if (X < W) {
        Px q = px_load(im, X, Y);
        px_store(out, X, Y, q.r, q.g, q.b);
      }
    }
  }
}

// ---- GRAY8 + bicubic form of the sweep above: four target pixels per lane --------------------
// A warp owns a 128 x 8 tile of the sheet (lane = four consecutive columns, one aligned 32-bit
// store per row), a CTA four such tiles stacked (128 x 32).  Per tile, once:
//   * the lane's four columns give the x-halves of the source coordinates,
//       srcX = (scx + (x - tcx) cos) + (y - tcy) sin,  srcY = (scy + (y - tcy) cos) - (x - tcx) sin
//     (the reference's evaluation order, deskew.c:263-268); a row then costs one add per coordinate;
//   * the source bounding box of (tile ∩ pasted rectangle) decides whether every tap of every
//     pixel lies inside the image (then no pixel needs a range test; otherwise the tile takes
//     the general per-pixel path) — columns of the tile outside the rectangle are evaluated at
//     the nearest column inside it and replaced by the sheet's own pixel afterwards, so tiles on
//     the rectangle's edge run the same code;
//   * the ink map decides whether the source footprint is pure white: lane = one column of ink
//     cells, which only looks at the cell rows the tile's rows can reach at that source column
//     (srcY = scy + (y - tcy) / cos - tan * (srcX - scx)), not at the whole bounding box.
// Per row and pixel: two aligned loads + a funnel shift per tap row, "all 16 taps equal" (every
// cubic term cancels exactly) decides per warp row whether the arithmetic runs at all.
#define RS_TW 128
#define RS_TH 8
// The four pixels of one lane on target row y (coordinates of the rotated image), GRAY8 + cubic, every tap inside the
// image: bX[j] = scx + ((float)x_j - tcx) * cos, xs[j] = ((float)x_j - tcx) * sin.  Warp-collective (one vote decides
// whether any lane's taps differ, i.e. whether the cubic arithmetic runs at all).
__device__ __forceinline__ unsigned rot_quad_g8c(const uint8_t *src, int pitch, const float (&bX)[4], const float (&xs)[4], int y,
                                                 float tcy, float scy, float sinval, float cosval, u64 nz2) {
  const float yr = u8f((unsigned)y) - tcy;                        // (float)y - tcy
  const float ysin = yr * sinval, ycos = scy + yr * cosval;
  unsigned rw[4][4], o[4];
  float fx[4], fy[4];
  bool need = false;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float sX = bX[j] + ysin, sY = ycos - xs[j];
    // (int)srcX for 1 <= srcX < 2^23 by the 2^23 trick (the tile's bounding box guarantees the range)
    const float tX = __fadd_rz(sX, 8388608.0f), tY = __fadd_rz(sY, 8388608.0f);
    const int px = __float_as_int(tX) - 0x4B000000, py = __float_as_int(tY) - 0x4B000000;
    fx[j] = sX - (tX - 8388608.0f); fy[j] = sY - (tY - 8388608.0f);   // srcX - (float)px
    const int off = (py - 1) * pitch + (px - 1);
    const unsigned sh = ((unsigned)off & 3u) * 8u;
    const uint8_t *p0 = src + (off & ~3);
#pragma unroll
    for (int i = 0; i < 4; i++, p0 += pitch) {
      const unsigned *wp = (const unsigned *)p0;
      rw[j][i] = __funnelshift_r(wp[0], wp[1], sh);
    }
    o[j] = rw[j][0] & 0xFFu;          // all 16 taps equal: every cubic term cancels exactly
    need = need || !(rw[j][0] == o[j] * 0x01010101u && rw[j][1] == rw[j][0] && rw[j][2] == rw[j][0] && rw[j][3] == rw[j][0]);
  }
  unsigned v01 = o[0] | (o[1] << 16), v23 = o[2] | (o[3] << 16);
  if (__any_sync(0xffffffffu, need)) {
    unsigned c4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned fb = __float_as_uint(fx[j]), hb = __float_as_uint(0.5f * fx[j]);
      const u64 f2 = pk2(fb, fb), h2 = pk2(hb, hb);
      const unsigned r01 = cubic_scale_w2p(f2, h2, rw[j][0], rw[j][1], nz2), r23 = cubic_scale_w2p(f2, h2, rw[j][2], rw[j][3], nz2);
      c4[j] = __byte_perm(r01, r23, 0x6420);            // the four row results as one tap word
    }
    v01 = cubic_scale_w2p(pk2(__float_as_uint(fy[0]), __float_as_uint(fy[1])),
                          pk2(__float_as_uint(0.5f * fy[0]), __float_as_uint(0.5f * fy[1])), c4[0], c4[1], nz2);
    v23 = cubic_scale_w2p(pk2(__float_as_uint(fy[2]), __float_as_uint(fy[3])),
                          pk2(__float_as_uint(0.5f * fy[2]), __float_as_uint(0.5f * fy[3])), c4[2], c4[3], nz2);
  }
  return __byte_perm(v01, v23, 0x6420);
}

__global__ void __launch_bounds__(128) k_rotate_sheet_g8c(DPage *pages, int mi, unsigned zero) {
  const DPage &pg = pages[blockIdx.z];
  const DImg im = pg.img;                 // by value: no reloads of the descriptor behind the stores below
  DImg out = im;
  out.data = pg.other;
  const int W = im.w, H = im.h, pitch = im.pitch;
  const int X0 = blockIdx.x * RS_TW, Yb0 = blockIdx.y * (4 * RS_TH);
  if (X0 >= W || Yb0 >= H) return;
  const bool active = mi < pg.mask_count && pg.rot_apply[mi];
  const DRect mask = pg.masks[mi];
  const int w = abs(mask.x0 - mask.x1) + 1, h = abs(mask.y0 - mask.y1) + 1;
  const int ox = mask.x0, oy = mask.y0;   // copy_rectangle(rotated, source, full, mask.vertex[0]) (deskew.c:283)
  const int X1 = min(X0 + RS_TW - 1, W - 1), Yb1 = min(Yb0 + 4 * RS_TH - 1, H - 1);
  const uint8_t *const src = im.data;
  uint8_t *const dst = out.data;
  __builtin_assume(__isGlobal(src));      // LDG / STG instead of generic accesses
  __builtin_assume(__isGlobal(dst));
  const bool aligned = im.fmt == DF_GRAY8 && (pitch & 15) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  const bool touches = active && X1 >= ox && X0 < ox + w && Yb1 >= oy && Yb0 < oy + h;
  if (!touches && aligned) {              // the CTA's 128 x 32 pixels are copied, 16 bytes per thread
    const int rows = Yb1 - Yb0 + 1, nch = (X1 - X0 + 16) >> 4;   // a row's pitch padding may be copied along
    for (int i = threadIdx.x; i < nch * rows; i += blockDim.x) {
      const int r = i / nch, c = i - r * nch;
      const size_t off = (size_t)(Yb0 + r) * pitch + (size_t)X0 + ((size_t)c << 4);
      *(uint4 *)(dst + off) = *(const uint4 *)(src + off);
    }
    return;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Yt = Yb0 + wid * RS_TH;
  if (Yt > Yb1) return;                   // whole warp
  const int Yte = min(Yt + RS_TH - 1, Yb1);
  const int X = X0 + 4 * lane;            // the lane's columns X .. X + 3
  const bool lane_on = X < pitch;         // its word lies inside the row (pitch is a multiple of 16)
  // center_of_rectangle (primitives.c:136-145) of the (normalised) mask / of the target
  const int nx0 = min(mask.x0, mask.x1), ny0 = min(mask.y0, mask.y1);
  const float scx = nx0 + w / 2.0f, scy = ny0 + h / 2.0f;
  const float tcx = 0 + w / 2.0f, tcy = 0 + h / 2.0f;
  const float sinval = pg.rot_sin[mi], cosval = pg.rot_cos[mi];
  const u64 nz2 = pk2(0x80000000u | zero, 0x80000000u | zero);      // -0.0 twice (`zero` is 0 at run time)
  enum { M_COPY, M_WHITE, M_FAST, M_SLOW };
  int mode = M_SLOW;
  int xa = 0, xb = 0;
  const bool t_touch = touches && Yte >= oy && Yt < oy + h;
  const bool tile_in = X0 >= ox && X0 + RS_TW - 1 < ox + w && X0 + RS_TW - 1 < W && Yt >= oy && Yte < oy + h;
  if (aligned && !t_touch) mode = M_COPY;
  else if (aligned) {
    // the part of the tile inside the pasted rectangle, in coordinates of the rotated image
    xa = max(X0, ox) - ox; xb = min(X1, ox + w - 1) - ox;
    const int ya = max(Yt, oy) - oy, yb = min(Yte, oy + h - 1) - oy;
    const float fxa = xa - tcx, fxb = xb - tcx, fya = ya - tcy, fyb = yb - tcy;
    const float ax = fxa * cosval, bx = fxb * cosval, ay = fya * sinval, by = fyb * sinval;
    const float cy = fya * cosval, dy = fyb * cosval, cx = fxa * sinval, dx = fxb * sinval;
    const float mnx = scx + fminf(ax, bx) + fminf(ay, by), mxx = scx + fmaxf(ax, bx) + fmaxf(ay, by);
    const float mny = scy + fminf(cy, dy) - fmaxf(cx, dx), mxy = scy + fmaxf(cy, dy) - fminf(cx, dx);
    if (fabsf(mnx) < 1e7f && fabsf(mxx) < 1e7f && fabsf(mny) < 1e7f && fabsf(mxy) < 1e7f) {
      // taps reach from (int)src - 1 to (int)src + 2.  The bounding box is evaluated in another order than the
      // pixels' coordinates (differences of the order of 1e-3 pixels): two pixels of slack for the range guarantee,
      // one for the white test
      const int bx0 = (int)floorf(mnx) - 3, bx1 = (int)floorf(mxx) + 4, by0 = (int)floorf(mny) - 3, by1 = (int)floorf(mxy) + 4;
      const bool safe = bx0 >= 0 && by0 >= 0 && bx1 < W && by1 < H;
      bool white = false;
      if (pg.ink_ok && fabsf(cosval) > 0.5f) {
        if (bx1 < 0 || by1 < 0 || bx0 >= W || by0 >= H) white = true;   // entirely outside: reads as white
        else {
          const int cx0 = max(bx0 + 1, 0) >> 3, cx1 = min(bx1 - 1, W - 1) >> 3;
          static_assert(D_INK_CELL == 8, "cell index = coordinate >> 3");
          if (cx1 - cx0 < 32) {
            bool wh = true;
            if (cx0 + lane <= cx1) {
              // pixels whose taps touch cell column c have (int)srcX in [8c - 2, 8c + 8]
              const int c = cx0 + lane;
              const float sxl = fmaxf((float)(8 * c - 3), mnx), sxh = fminf((float)(8 * c + 10), mxx);
              const float icos = 1.0f / cosval, tanv = sinval * icos;
              const float e0 = scy + fya * icos, e1 = scy + fyb * icos;
              const float t0 = tanv * (sxl - scx), t1 = tanv * (sxh - scx);
              const float sylo = fminf(e0, e1) - fmaxf(t0, t1), syhi = fmaxf(e0, e1) - fminf(t0, t1);
              const int ry0 = max((int)floorf(sylo) - 2, by0), ry1 = min((int)floorf(syhi) + 3, by1);
              if (ry1 >= 0 && ry0 < H && ry0 <= ry1) {
                const int cyl = max(ry0, 0) >> 3, cyh = min(ry1, H - 1) >> 3;
                if (cyh - cyl > 7) wh = false;
                else {
                  const uint8_t *q = pg.ink + (size_t)cyl * pg.ink_ncx + c;
                  for (int cc = cyl; cc <= cyh; cc++, q += pg.ink_ncx) wh = wh && (*q != 0);
                }
              }
            }
            white = __all_sync(0xffffffffu, wh);
          }
        }
      }
      if (white) mode = M_WHITE;
      else if (safe) mode = M_FAST;
    }
  }
  if (mode == M_COPY) {
    if (lane_on)
      for (int Y = Yt; Y <= Yte; Y++) {
        const size_t off = (size_t)Y * pitch + X;
        *(unsigned *)(dst + off) = *(const unsigned *)(src + off);
      }
    return;
  }
  if (mode == M_SLOW) {
    // taps outside the image (reads = white), unaligned buffers: pixel by pixel
    for (int Y = Yt; Y <= Yte; Y++) {
      const int y = Y - oy;
      for (int j = 0; j < 4; j++) {
        const int Xj = X + j, x = Xj - ox;
        if (Xj >= W) break;
        if (active && x >= 0 && x < w && y >= 0 && y < h) {
          const float xf = u8f((unsigned)x), yf = u8f((unsigned)y);   // == (float)x, (float)y for 0 <= v < 2^23
          const float sX = scx + (xf - tcx) * cosval + (yf - tcy) * sinval;
          const float sY = scy + (yf - tcy) * cosval - (xf - tcx) * sinval;
          const Px q = interp_cubic<true>(im, sX, sY);
          px_store(out, Xj, Y, q.r, q.g, q.b);
        } else {
          const Px q = px_load(im, Xj, Y);
          px_store(out, Xj, Y, q.r, q.g, q.b);
        }
      }
    }
    return;
  }
  if (mode == M_FAST && tile_in && Yte - Yt == RS_TH - 1) {
    // A full tile inside the pasted rectangle: the warp takes it as 4 x 2 blocks of 32 x 4 pixels (8 lanes along x,
    // 4 rows) instead of eight rows of 128.  The vote that skips the cubic arithmetic then covers a compact block —
    // on text pages 36 % of the blocks hold a pixel with unequal taps against 46 % of the 128-pixel rows.
    const int lx = lane & 7, ly = lane >> 3;
#pragma unroll 1
    for (int bq = 0; bq < RS_TW / 32; bq++) {
      const int Xq = X0 + 32 * bq + 4 * lx;
      float qbX[4], qxs[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float xr = u8f((unsigned)(Xq + j - ox)) - tcx;            // (float)x - tcx
        qbX[j] = scx + xr * cosval;
        qxs[j] = xr * sinval;
      }
#pragma unroll 1
      for (int hq = 0; hq < RS_TH / 4; hq++) {
        const int Y = Yt + 4 * hq + ly;
        const unsigned vals = rot_quad_g8c(src, pitch, qbX, qxs, Y - oy, tcy, scy, sinval, cosval, nz2);
        *(unsigned *)(dst + (size_t)Y * pitch + Xq) = vals;
      }
    }
    return;
  }
  // ---- M_WHITE / M_FAST
  unsigned xmask = 0;                     // bytes of the lane's word that lie inside the pasted rectangle
  float bX[4], xs[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int Xj = X + j, xj = Xj - ox;
    if (Xj < W && xj >= 0 && xj < w) xmask |= 0xFFu << (8 * j);
    const float xr = u8f((unsigned)min(max(xj, xa), xb)) - tcx;      // (float)x - tcx
    bX[j] = scx + xr * cosval;
    xs[j] = xr * sinval;
  }
  for (int Y = Yt; Y <= Yte; Y++) {
    const size_t rowoff = (size_t)Y * pitch + X;
    const int y = Y - oy;
    if (y < 0 || y >= h) {                // warp-uniform: a row of the tile outside the rectangle
      if (lane_on) *(unsigned *)(dst + rowoff) = *(const unsigned *)(src + rowoff);
      continue;
    }
    unsigned vals = 0xFFFFFFFFu;
    if (mode == M_FAST) vals = rot_quad_g8c(src, pitch, bX, xs, y, tcy, scy, sinval, cosval, nz2);
    if (!lane_on) continue;
    if (!tile_in) {
      const unsigned sw = *(const unsigned *)(src + rowoff);
      vals = (vals & xmask) | (sw & ~xmask);
    }
    *(unsigned *)(dst + rowoff) = vals;
  }
}

// stretch_frame (blit.c:209-228)
__global__ void k_stretch(DImg src, DImg dst, float hr, float vr, int interp) {
  int y = blockIdx.y;
  if (y >= dst.h) return;
  bool gray = src.fmt != DF_RGB24;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < dst.w; x += gridDim.x * blockDim.x) {
    Px o = interp_any(src, x * hr, y * vr, interp, gray);
    px_store(dst, x, y, o.r, o.g, o.b);
  }
}

// stretch_and_replace of every page's working sheet into its other buffer (sheet_stages.c:216-222, :516-522)
__global__ void k_stretch_pages(DPage *pages, int dw, int dh, int dpitch, float hr, float vr, int interp) {
  const DPage &pg = pages[blockIdx.z];
  DImg src = pg.img, dst = pg.img;
  dst.data = pg.other; dst.w = dw; dst.h = dh; dst.pitch = dpitch;
  bool gray = src.fmt != DF_RGB24;
  for (int y = blockIdx.y; y < dh; y += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < dw; x += gridDim.x * blockDim.x) {
      Px o = interp_any(src, x * hr, y * vr, interp, gray);
      px_store(dst, x, y, o.r, o.g, o.b);
    }
}

extern "C" {
void b200k_stretch_pages(cudaStream_t st, DPage *pages, int npages, int sw, int sh, int dw, int dh, int dpitch, int interp) {
  if (npages <= 0 || dw <= 0 || dh <= 0) return;
  float hr = (float)sw / (float)dw, vr = (float)sh / (float)dh;   /* blit.c:213-216 */
  dim3 g(min(cdiv(dw, 128), 32u), min((unsigned)dh, 2048u), npages);
  k_stretch_pages<<<g, 128, 0, st>>>(pages, dw, dh, dpitch, hr, vr, interp);
}
int b200k_rot_peaks(cudaStream_t st, DPage *pages, int npages, int mi_first, int mi_count, const float *tan_tab_dev,
                    int nangles, int scan_size_param, float scan_depth, const int edges[4],
                    int peak_off, int scan_cap, int maxw, int use_prefix, int run_cap) {
  if (npages <= 0 || mi_count <= 0 || nangles <= 0) return 0;
  RotParams rp;
  rp.scan_size = scan_size_param; rp.scan_depth = scan_depth; rp.nangles = nangles; rp.peak_off = peak_off;
  for (int i = 0; i < 4; i++) rp.edges[i] = edges[i];
  size_t sm = (size_t)scan_cap * sizeof(int2);
  if (sm > 200 * 1024) return -1;
  if (sm > 40 * 1024) {
    cudaFuncSetAttribute(k_rot_peaks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_rot_peaks_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  }
  bool horiz = edges[0] || edges[2], vert = edges[1] || edges[3];
  if (use_prefix && horiz) {
    for (int mi = mi_first; mi < mi_first + mi_count; mi++) {
      if (run_cap > 0) {
        int rstride = run_cap | 1;   // odd stride: the lanes of phase A hit different banks
        size_t smw = (size_t)(2 * RW_ANG * rstride) * sizeof(int);
        for (int phase = 1; phase <= 2; phase++) {
          k_rot_colprefix<<<dim3(cdiv(maxw, 128), npages), 32 * RP_SEG, 0, st>>>(pages, mi, scan_size_param, ROT_DEPTH_CAP, 0, phase);
          k_rot_peaks_w<<<dim3(cdiv(nangles, RW_ANG), 2, npages), 32 * RW_ANG, smw, st>>>(pages, tan_tab_dev, rp, mi, rstride, ROT_DEPTH_CAP, phase);
        }
      } else {
        k_rot_colprefix<<<dim3(cdiv(maxw, 128), npages), 32 * RP_SEG, 0, st>>>(pages, mi, scan_size_param, 0, 0, 0);
        k_rot_peaks_h<<<dim3(nangles, 2, npages), 64, sm, st>>>(pages, tan_tab_dev, rp, mi);
      }
    }
    rp.edges[0] = 0; rp.edges[2] = 0;   // the sampling kernel below only does what is left
  }
  if ((!use_prefix && horiz) || vert) {
    dim3 g(nangles, mi_count * 4, npages);
    k_rot_peaks<<<g, 256, sm, st>>>(pages, tan_tab_dev, rp, mi_first);
  }
  return 0;
}
void b200k_rot_finalize(cudaStream_t st, DPage *pages, int npages, const float *rot_tab_dev,
                        const float *pair_tab_dev, int nangles, const int edges[4], int peak_off,
                        float deviation, int mi_first, int mi_count) {
  if (npages <= 0) return;
  RotFinalParams fp;
  fp.nangles = nangles; fp.peak_off = peak_off; fp.deviation = deviation; fp.use_table = pair_tab_dev != NULL;
  fp.mi_first = mi_first; fp.mi_count = mi_count;
  for (int i = 0; i < 4; i++) fp.edges[i] = edges[i];
  k_rot_finalize<<<cdiv(npages * D_MAX_MASKS, 64), 64, 0, st>>>(pages, npages, rot_tab_dev, pair_tab_dev, fp);
}
void b200k_rot_set_sincos(cudaStream_t st, DPage *pages, int npages, int mi, const float *tab_dev) {
  if (npages <= 0) return;
  k_rot_set<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, mi, tab_dev);
}
void b200k_rotate(cudaStream_t st, DPage *pages, int npages, int mi, int interp, int maxw, int maxh,
                  DCopyJob *back_jobs) {
  if (npages <= 0 || maxw <= 0 || maxh <= 0) return;
  if (interp == 2) k_inkmap<<<dim3(min(cdiv(maxh, INK_CELL), 256u), npages), 128, 0, st>>>(pages);
  dim3 g(min(cdiv(maxw, 128), 64u), cdiv(maxh, ROT_ROWS), npages);
  k_rotate<<<g, 128, 0, st>>>(pages, mi, interp, back_jobs);
}
void b200k_rotate_sheet(cudaStream_t st, DPage *pages, int npages, int mi, int interp, int fmt, int maxw, int maxh, int ink_fresh) {
  if (npages <= 0 || maxw <= 0 || maxh <= 0) return;
  if (interp == 2 && !ink_fresh) k_inkmap<<<dim3(min(cdiv(maxh, INK_CELL), 256u), npages), 128, 0, st>>>(pages);
  if (interp == 2 && fmt == DF_GRAY8) {
    k_rotate_sheet_g8c<<<dim3(cdiv(maxw, RS_TW), cdiv(maxh, 4 * RS_TH), npages), 128, 0, st>>>(pages, mi, 0u);
    return;
  }
  dim3 g(cdiv(maxw, 128), cdiv(maxh, ROT_ROWS), npages);
  k_rotate_sheet<<<g, 128, 0, st>>>(pages, mi, interp, 0u);
}
void b200k_stretch(cudaStream_t st, DImg src, DImg dst, float hr, float vr, int interp) {
  if (dst.w <= 0 || dst.h <= 0) return;
  dim3 g(min(cdiv(dst.w, 128), 64u), dst.h, 1);
  k_stretch<<<g, 128, 0, st>>>(src, dst, hr, vr, interp);
}
}
