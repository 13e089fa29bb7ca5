// k_stats.cu — one-pass image statistics every detector consumes:
//   * band line sums (column sums over a row band / row sums over a column
//     band) of gray, max-channel or "gray in [lo,hi]" indicators — feeds
//     detect_masks (masks.c:54-100 via blit.c:91-106), detect_border
//     (masks.c:410-448 via blit.c:148-167) and the blackfilter scan
//     (filters.c:49-104 via blit.c:131-146);
//   * rectangle counts for blurfilter (filters.c:176-205);
//   * g x g cell statistics for grayfilter (filters.c:377-389).
// Replaces reference cuda_kernels_masks.cu:13-158 (one 64-bit atomic per pixel)
// and the NPP integral images (npp_integral.c).
#include "common.cuh"
#include "launch.h"
#include "swar.h"

__device__ __forceinline__ unsigned stat_of(Px p, int stat, int lo, int hi) {
  if (stat == ST_GRAY) return (unsigned)px_gray(p);
  if (stat == ST_MAXCH) return (unsigned)px_darkinv(p);
  int g = px_gray(p);
  return (g >= lo && g <= hi) ? 1u : 0u;
}


// ---- GRAY8 helpers: four pixels per 32-bit word ---------------------------------
__device__ __forceinline__ unsigned sum4(unsigned w) { return __vsadu4(w, 0u); }          // b0+b1+b2+b3
// number of bytes of w that lie in [lo, hi] (0 <= lo <= hi <= 255)
__device__ __forceinline__ unsigned count4_range(unsigned w, unsigned lo4, unsigned hi4) {
  unsigned ge = __vcmpgeu4(w, lo4), le = __vcmpleu4(w, hi4);
  return (unsigned)__popc(ge & le) >> 3;
}
// three-instruction byte compares: lt4 / range4_bit7 (swar.h)
// LO0: the lower bound is 0 (every caller on the sheet path: "dark" = gray in [0, threshold]) and costs nothing
template <bool LO0>
__device__ __forceinline__ unsigned range4_bit7_t(unsigned w, Lt4 loT, Lt4 hiT) { return LO0 ? lt4(w, hiT) : range4_bit7(w, loT, hiT); }
// Applies f(word, nvalid_mask) over the bytes [p, p+n): aligned 32-bit loads,
// `keep` has 0xFF in the byte lanes that belong to the run.
template <typename F>
__device__ __forceinline__ void for_words(const uint8_t *p, int n, int lane, int nlanes, F f) {
  if (n <= 0) return;
  uintptr_t a = (uintptr_t)p;
  const unsigned *w0 = (const unsigned *)(a & ~(uintptr_t)3);
  int lead = (int)(a & 3);                  // bytes of the first word that precede the run
  int nw = (lead + n + 3) >> 2;
  for (int i = lane; i < nw; i += nlanes) {
    unsigned keep = 0xFFFFFFFFu;
    if (i == 0 && lead) keep &= 0xFFFFFFFFu << (8 * lead);
    int end = lead + n - 4 * i;             // valid bytes of this word counted from its start
    if (end < 4) keep &= 0xFFFFFFFFu >> (8 * (4 - end));
    f(w0[i], keep);
  }
}

__global__ void k_zero_u32(DPage *pages, int off, int n) {
  unsigned *p = pages[blockIdx.y].u32 + off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0u;
}

// Column sums: thread = column, block = 256 columns x ROWS rows; partial sums
// merged with one atomic per (block, column).  blockIdx.z = page*njobs + job.
#define LS_ROWS 128
__global__ void k_linesum_cols(DPage *pages, const DLineJob *jobs, int njobs, int stat, int lo, int hi) {
  int page = blockIdx.z / njobs, job = blockIdx.z % njobs;
  const DLineJob j = jobs[job];
  if (j.axis != 0 || j.xa > j.xb || j.ya > j.yb) return;
  const DImg &im = pages[page].img;
  int x = j.xa + blockIdx.x * blockDim.x + threadIdx.x;
  int y0 = j.ya + blockIdx.y * LS_ROWS;
  if (x > j.xb || y0 > j.yb) return;
  int y1 = min(y0 + LS_ROWS - 1, j.yb);
  unsigned acc = 0;
  if (im.fmt == DF_GRAY8 && stat != ST_COUNT_GRAY_RANGE) {
    const uint8_t *p = im.data + (size_t)y0 * im.pitch + x;
    for (int y = y0; y <= y1; y++, p += im.pitch) acc += *p;
  } else {
    for (int y = y0; y <= y1; y++) acc += stat_of(px_load(im, x, y), stat, lo, hi);
  }
  atomicAdd(pages[page].u32 + j.out_off + (x - j.xa), acc);
}

// GRAY8 column sums, four adjacent columns per thread (one aligned 32-bit load per row), for jobs
// whose columns start on a 4-byte boundary.  With make_ink (the job covers the whole image, whose
// width is a multiple of 8) the pass also leaves the page's ink map behind — which 8 x 8 cells
// are pure white (see k_inkmap in k_deskew.cu) — so that the rotation that follows needs no pass
// of its own over the sheet.
__global__ void __launch_bounds__(128) k_linesum_cols4(DPage *pages, const DLineJob *jobs, int njobs, int make_ink) {
  int page = blockIdx.z / njobs, job = blockIdx.z % njobs;
  const DLineJob j = jobs[job];
  if (j.axis != 0 || j.xa > j.xb || j.ya > j.yb) return;
  DPage &pg = pages[page];
  const DImg im = pg.img;
  int x = j.xa + 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  int y0 = j.ya + blockIdx.y * LS_ROWS;
  if (y0 > j.yb) return;
  int y1 = min(y0 + LS_ROWS - 1, j.yb);
  const int ncx = (im.w + 7) >> 3, ncy = (im.h + 7) >> 3;
  bool ink = make_ink && pg.ink && ncx * ncy <= pg.ink_cap;
  if (ink && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { pg.ink_ncx = ncx; pg.ink_ncy = ncy; pg.ink_ok = 1; }
  bool live = x <= j.xb;                         // the job's width is a multiple of 4
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0, all = 0xFFFFFFFFu;
  const uint8_t *p = im.data + (size_t)y0 * im.pitch + (live ? x : j.xa);
  for (int y = y0; y <= y1; y++, p += im.pitch) {
    unsigned w = *(const unsigned *)p;
    a0 += w & 0xFFu; a1 += (w >> 8) & 0xFFu; a2 += (w >> 16) & 0xFFu; a3 += w >> 24;
    if (ink) {
      all &= w;
      if ((y & 7) == 7 || y == y1) {
        // a cell = 8 columns = this thread and its neighbour (x is a multiple of 4, cells start at multiples of 8)
        unsigned mine = all == 0xFFFFFFFFu, other = __shfl_xor_sync(0xffffffffu, mine, 1);
        if (live && !(threadIdx.x & 1)) pg.ink[(y >> 3) * ncx + (x >> 3)] = (uint8_t)(mine & other);
        all = 0xFFFFFFFFu;
      }
    }
  }
  if (live) {
    unsigned *o = pg.u32 + j.out_off + (x - j.xa);
    atomicAdd(o, a0); atomicAdd(o + 1, a1); atomicAdd(o + 2, a2); atomicAdd(o + 3, a3);
  }
}

// The same with sixteen columns per thread (one aligned 16-byte load per row; two 16-bit partial sums per
// register: LS_ROWS * 255 < 65536), for jobs whose columns start on a 16-byte boundary.  A thread owns two
// ink cells, so no exchange is needed.
#define LS_ROWS16 128  /* rows per block (32 measured slower: four times the atomics on the same 2480 sums) */
__global__ void __launch_bounds__(64) k_linesum_cols16(DPage *pages, const DLineJob *jobs, int njobs, int make_ink) {
  int page = blockIdx.z / njobs, job = blockIdx.z % njobs;
  const DLineJob j = jobs[job];
  if (j.axis != 0 || j.xa > j.xb || j.ya > j.yb) return;
  DPage &pg = pages[page];
  const DImg im = pg.img;
  int x = j.xa + 16 * (blockIdx.x * blockDim.x + threadIdx.x);
  int y0 = j.ya + blockIdx.y * LS_ROWS16;
  if (y0 > j.yb || x > j.xb) return;
  int y1 = min(y0 + LS_ROWS16 - 1, j.yb);
  static_assert(LS_ROWS16 % 8 == 0 && LS_ROWS16 * 255 < 65536, "whole ink cells per block, 16-bit partial sums");
  const int ncx = (im.w + 7) >> 3, ncy = (im.h + 7) >> 3;
  bool ink = make_ink && pg.ink && ncx * ncy <= pg.ink_cap;
  if (ink && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { pg.ink_ncx = ncx; pg.ink_ncy = ncy; pg.ink_ok = 1; }
  unsigned lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};      // lo[k]: bytes 0 and 2 of word k, hi[k]: bytes 1 and 3
  uint4 all = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  const uint8_t *p = im.data + (size_t)y0 * im.pitch + x;
  // a block's rows are one dependent chain per thread (every block of the grid is resident at once), so the time
  // is rows / loads in flight x DRAM latency: eight rows (= one row of ink cells) per batch, loads first
  int y = y0;
  if (!ink || (y0 & 7) == 0)
    for (; y + 7 <= y1; y += 8, p += 8 * (size_t)im.pitch) {
      uint4 w[8];
#pragma unroll
      for (int k = 0; k < 8; k++) w[k] = __ldg((const uint4 *)(p + (size_t)k * im.pitch));
#pragma unroll
      for (int k = 0; k < 8; k++) {
        lo[0] += w[k].x & 0x00FF00FFu; hi[0] += (w[k].x >> 8) & 0x00FF00FFu;
        lo[1] += w[k].y & 0x00FF00FFu; hi[1] += (w[k].y >> 8) & 0x00FF00FFu;
        lo[2] += w[k].z & 0x00FF00FFu; hi[2] += (w[k].z >> 8) & 0x00FF00FFu;
        lo[3] += w[k].w & 0x00FF00FFu; hi[3] += (w[k].w >> 8) & 0x00FF00FFu;
        all.x &= w[k].x; all.y &= w[k].y; all.z &= w[k].z; all.w &= w[k].w;
      }
      if (ink) {
        uint8_t *c = pg.ink + (y >> 3) * ncx + (x >> 3);
        c[0] = (uint8_t)((all.x & all.y) == 0xFFFFFFFFu);
        c[1] = (uint8_t)((all.z & all.w) == 0xFFFFFFFFu);
      }
      all = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
  for (; y <= y1; y++, p += im.pitch) {     // what is left (and jobs whose blocks do not start on a cell row)
    uint4 w = *(const uint4 *)p;
    lo[0] += w.x & 0x00FF00FFu; hi[0] += (w.x >> 8) & 0x00FF00FFu;
    lo[1] += w.y & 0x00FF00FFu; hi[1] += (w.y >> 8) & 0x00FF00FFu;
    lo[2] += w.z & 0x00FF00FFu; hi[2] += (w.z >> 8) & 0x00FF00FFu;
    lo[3] += w.w & 0x00FF00FFu; hi[3] += (w.w >> 8) & 0x00FF00FFu;
    if (ink) {
      all.x &= w.x; all.y &= w.y; all.z &= w.z; all.w &= w.w;
      if ((y & 7) == 7 || y == y1) {
        uint8_t *c = pg.ink + (y >> 3) * ncx + (x >> 3);
        c[0] = (uint8_t)((all.x & all.y) == 0xFFFFFFFFu);
        c[1] = (uint8_t)((all.z & all.w) == 0xFFFFFFFFu);
        all = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      }
    }
  }
  unsigned *o = pg.u32 + j.out_off + (x - j.xa);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    atomicAdd(o + 4 * k + 0, lo[k] & 0xFFFFu); atomicAdd(o + 4 * k + 1, hi[k] & 0xFFFFu);
    atomicAdd(o + 4 * k + 2, lo[k] >> 16); atomicAdd(o + 4 * k + 3, hi[k] >> 16);
  }
}

// A lane's share of row [xa, xb] of an aligned GRAY8 row: 16 bytes per load, a row's loads independent of each
// other (a 2480-pixel row is five of them per lane, all in flight); only the two chunks holding xa and xb mask
// bytes off.  MODE 0: sum of the bytes, 1: bytes <= hi, 2: bytes in [lo, hi].
template <int MODE>
__device__ __forceinline__ unsigned row_words(const uint4 *r4, int xa, int xb, int lane, Lt4 loT, Lt4 hiT) {
  const int c0 = xa >> 4, c1 = xb >> 4;
  unsigned acc = 0;
#pragma unroll 4
  for (int c = c0 + lane; c <= c1; c += 32) {
    const uint4 v = __ldg(r4 + c);
    unsigned wv[4] = {v.x, v.y, v.z, v.w}, keep[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (c == c0 || c == c1) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int x0 = 16 * c + 4 * k;
        const int dl = min(max(xa - x0, 0), 4), dh = min(max(x0 + 3 - xb, 0), 4);   // bytes to drop at either end
        keep[k] = (dl + dh >= 4) ? 0u : ((0xFFFFFFFFu << (8 * dl)) & (0xFFFFFFFFu >> (8 * dh)));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (MODE == 0) acc += sum4(wv[k] & keep[k]);
      else acc += (unsigned)__popc(range4_bit7_t<MODE == 1>(wv[k], loT, hiT) & keep[k]);
    }
  }
  return acc;
}

// Row sums: one warp per row.
__global__ void k_linesum_rows(DPage *pages, const DLineJob *jobs, int njobs, int stat, int lo, int hi) {
  int page = blockIdx.z / njobs, job = blockIdx.z % njobs;
  const DLineJob j = jobs[job];
  if (j.axis != 1 || j.xa > j.xb || j.ya > j.yb) return;
  const DImg &im = pages[page].img;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int y = j.ya + blockIdx.y * (blockDim.x >> 5) + warp;
  if (y > j.yb) return;
  unsigned acc = 0;
  if (im.fmt == DF_GRAY8 && (im.pitch & 15) == 0 && ((uintptr_t)im.data & 15) == 0) {
    // aligned rows: 16 bytes per lane and load, a row's loads independent of each other (a 2480-pixel row
    // is five of them per lane, all in flight); only the two chunks holding xa and xb mask bytes off
    const uint4 *r4 = (const uint4 *)(im.data + (size_t)y * im.pitch);
    const Lt4 loT = lt4_make(lo), hiT = lt4_make(hi + 1);
    acc = stat != ST_COUNT_GRAY_RANGE ? row_words<0>(r4, j.xa, j.xb, lane, loT, hiT)
          : lo <= 0 ? row_words<1>(r4, j.xa, j.xb, lane, loT, hiT) : row_words<2>(r4, j.xa, j.xb, lane, loT, hiT);
  } else if (im.fmt == DF_GRAY8) {
    const uint8_t *row = im.data + (size_t)y * im.pitch + j.xa;
    int n = j.xb - j.xa + 1;
    if (stat == ST_COUNT_GRAY_RANGE) {
      unsigned lo4 = (unsigned)lo * 0x01010101u, hi4 = (unsigned)hi * 0x01010101u;
      // bytes outside the run are forced to a value outside [lo,hi] when one exists; otherwise subtracted
      for_words(row, n, lane, 32, [&](unsigned w, unsigned keep) {
        unsigned c = __vcmpgeu4(w, lo4) & __vcmpleu4(w, hi4) & keep;
        acc += (unsigned)__popc(c) >> 3;
      });
    } else {
      for_words(row, n, lane, 32, [&](unsigned w, unsigned keep) { acc += sum4(w & keep); });
    }
  } else {
    for (int x = j.xa + lane; x <= j.xb; x += 32) acc += stat_of(px_load(im, x, y), stat, lo, hi);
  }
  acc = warp_sum_u32(acc);
  if (lane == 0) pages[page].u32[j.out_off + (y - j.ya)] = acc;
}

// Count of pixels with gray in [lo,hi] inside rect k of a static list; pixels
// outside the image read as white (blit.c:148-167 does not clip).  One warp per
// rectangle.  out = pages[p].u32 + out_off + k.
__global__ void __launch_bounds__(256, 8) k_rect_count(DPage *pages, const DRect *rects, int nrects, int lo, int hi, int out_off) {
  int page = blockIdx.y;
  int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (k >= nrects) return;
  const DImg &im = pages[page].img;
  DRect r = rects[k];
  long long total = (r.x1 >= r.x0 && r.y1 >= r.y0) ? (long long)(r.x1 - r.x0 + 1) * (r.y1 - r.y0 + 1) : 0;
  int x0 = max(r.x0, 0), x1 = min(r.x1, im.w - 1), y0 = max(r.y0, 0), y1 = min(r.y1, im.h - 1);
  unsigned cnt = 0;
  long long inside = 0;
  if (x0 <= x1 && y0 <= y1 && total > 0) {
    inside = (long long)(x1 - x0 + 1) * (y1 - y0 + 1);
    int w = x1 - x0 + 1;
    if (im.fmt == DF_GRAY8 && (im.pitch & 3) == 0 && ((uintptr_t)im.data & 3) == 0) {
      // every row of the rectangle has the same word alignment: a lane owns one
      // word column and walks down the rows
      const Lt4 loT = lt4_make(lo), hiT = lt4_make(hi + 1);
      int lead = x0 & 3, nw = (lead + w + 3) >> 2;
      const unsigned *base = (const unsigned *)(im.data + (size_t)y0 * im.pitch + (x0 - lead));
      int wpitch = im.pitch >> 2, rows = y1 - y0 + 1;
      for (int i = lane; i < nw; i += 32) {
        unsigned keep = 0xFFFFFFFFu;
        if (i == 0 && lead) keep &= 0xFFFFFFFFu << (8 * lead);
        int end = lead + w - 4 * i;
        if (end < 4) keep &= 0xFFFFFFFFu >> (8 * (4 - end));
        const unsigned *p = base + i;
        unsigned acc = 0;
        if (lo <= 0)
          for (int r = 0; r < rows; r++, p += wpitch) acc += (unsigned)__popc(range4_bit7_t<true>(*p, loT, hiT) & keep);
        else
          for (int r = 0; r < rows; r++, p += wpitch) acc += (unsigned)__popc(range4_bit7_t<false>(*p, loT, hiT) & keep);
        cnt += acc;
      }
    } else if (im.fmt == DF_GRAY8) {
      unsigned lo4 = (unsigned)lo * 0x01010101u, hi4 = (unsigned)hi * 0x01010101u;
      for (int yy = y0; yy <= y1; yy++)
        for_words(im.data + (size_t)yy * im.pitch + x0, w, lane, 32, [&](unsigned wd, unsigned keep) {
          unsigned c = __vcmpgeu4(wd, lo4) & __vcmpleu4(wd, hi4) & keep;
          cnt += (unsigned)__popc(c) >> 3;
        });
    } else {
      for (int yy = y0; yy <= y1; yy++)
        for (int xx = x0 + lane; xx <= x1; xx += 32) {
          int g = px_gray(px_load(im, xx, yy));
          cnt += (g >= lo && g <= hi) ? 1u : 0u;
        }
    }
  }
  cnt = warp_sum_u32(cnt);
  if (lane == 0) {
    if (255 >= lo && 255 <= hi) cnt += (unsigned)(total - inside);
    pages[page].u32[out_off + k] = cnt;
  }
}

// gx x gy cell statistics over a grid of ncx x ncy cells anchored at (0,0):
//   dark[c]  = #pixels (in image) with gray <= dark_max
//   light[c] = sum of min-channel over the in-image pixels
// One block per cell row; columns accumulate in registers over the g rows and
// merge into shared per-cell counters.  Layout in u32: [dark ncx*ncy][light ncx*ncy].
__global__ void __launch_bounds__(256, 4) k_cellstats(DPage *pages, int gx, int gy, int ncx, int ncy, int dark_max, int out_off) {
  extern __shared__ unsigned sm[];
  unsigned *sd = sm, *sl = sm + ncx;
  int page = blockIdx.y, cy = blockIdx.x;
  const DImg &im = pages[page].img;
  for (int i = threadIdx.x; i < 2 * ncx; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  int y0 = cy * gy, y1 = min(y0 + gy - 1, im.h - 1);
  int xlim = min(ncx * gx, im.w);
  if (im.fmt == DF_GRAY8 && gx >= 4 && (im.pitch & 15) == 0 && ((uintptr_t)im.data & 15) == 0) {
    // sixteen pixels per load: the four words of a chunk are four instances of the word form below, with the
    // loads of a chunk column's rows in flight together
    const Lt4 dmT = lt4_make(min(max(dark_max, -1) + 1, 256));   // v <= dark_max  <=>  v < dark_max + 1
    const uint8_t *const base = im.data;
    const int pitch = im.pitch;
    for (int x = 16 * threadIdx.x; x < xlim; x += 16 * blockDim.x) {
      int c[4];
      unsigned m0[4], m1[4];
      unsigned d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, l0[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int xk = x + 4 * k;
        c[k] = xk / gx;
        const int nb = min((c[k] + 1) * gx - xk, 4);           // bytes of this word in cell c
        const int nv = min(max(xlim - xk, 0), 4);              // bytes of this word inside the grid
        const unsigned mv = nv >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nv)) - 1u);
        m0[k] = (nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u)) & mv;
        m1[k] = mv & ~m0[k];
      }
      const uint8_t *p = base + (size_t)y0 * pitch + x;
#pragma unroll 5
      for (int y = y0; y <= y1; y++, p += pitch) {
        const uint4 q = __ldg((const uint4 *)p);
        const unsigned v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const unsigned dk = lt4(v[k], dmT);                  // bit 7 of every dark byte
          d0[k] += __popc(dk & m0[k]); d1[k] += __popc(dk & m1[k]);
          l0[k] += __vsadu4(v[k] & m0[k], 0u); l1[k] += __vsadu4(v[k] & m1[k], 0u);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(m0[k] | m1[k])) continue;
        if (d0[k]) atomicAdd(&sd[c[k]], d0[k]);
        if (m0[k]) atomicAdd(&sl[c[k]], l0[k]);
        if (m1[k]) { if (d1[k]) atomicAdd(&sd[c[k] + 1], d1[k]); atomicAdd(&sl[c[k] + 1], l1[k]); }
      }
    }
  } else if (im.fmt == DF_GRAY8 && gx >= 4 && (im.pitch & 3) == 0 && ((uintptr_t)im.data & 3) == 0) {
    // four pixels per 32-bit load; a word touches at most two cells (gx >= 4): the first
    // `nb` bytes belong to cell c, the rest to cell c + 1.  Packed-byte compare / SAD do
    // the counting and the sum.
    unsigned dm4 = (unsigned)min(max(dark_max, -1) + 1, 256);   // v <= dark_max  <=>  v < dark_max + 1
    bool all_dark = dm4 > 255u;
    unsigned t4 = (dm4 & 0xFFu) * 0x01010101u;
    for (int x = 4 * threadIdx.x; x < xlim; x += 4 * blockDim.x) {
      int c = x / gx;
      int nb = min((c + 1) * gx - x, 4);               // bytes of this word in cell c
      int nv = min(xlim - x, 4);                       // bytes of this word inside the grid
      unsigned mv = nv >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nv)) - 1u);
      unsigned m0 = (nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u)) & mv, m1 = mv & ~m0;
      unsigned d0 = 0, d1 = 0, l0 = 0, l1 = 0;
#pragma unroll 5
      for (int y = y0; y <= y1; y++) {
        unsigned v = __ldg((const unsigned *)(im.data + (size_t)y * im.pitch + x));
        unsigned dk = all_dark ? 0xFFFFFFFFu : (dm4 == 0u ? 0u : __vcmpltu4(v, t4));
        d0 += __popc(dk & m0); d1 += __popc(dk & m1);
        l0 += __vsadu4(v & m0, 0u); l1 += __vsadu4(v & m1, 0u);
      }
      if (d0) atomicAdd(&sd[c], d0 >> 3);
      atomicAdd(&sl[c], l0);
      if (m1) { if (d1) atomicAdd(&sd[c + 1], d1 >> 3); atomicAdd(&sl[c + 1], l1); }
    }
  } else if (im.fmt == DF_GRAY8 && (gx & 1) == 0 && (im.pitch & 1) == 0 && ((uintptr_t)im.data & 1) == 0) {
    // two pixels per 16-bit load: a pixel pair never straddles a cell when gx is even
    unsigned dm = (unsigned)dark_max;
    for (int x = 2 * threadIdx.x; x < xlim; x += 2 * blockDim.x) {
      unsigned d = 0, l = 0;
      bool two = x + 1 < xlim;
      for (int y = y0; y <= y1; y++) {
        unsigned v = *(const unsigned short *)(im.data + (size_t)y * im.pitch + x);
        unsigned a = v & 0xFFu, b = v >> 8;
        d += (a <= dm); l += a;
        if (two) { d += (b <= dm); l += b; }
      }
      int c = x / gx;
      if (d) atomicAdd(&sd[c], d);
      atomicAdd(&sl[c], l);
    }
  } else
  for (int x = threadIdx.x; x < xlim; x += blockDim.x) {
    unsigned d = 0, l = 0;
    for (int y = y0; y <= y1; y++) {
      Px p = px_load(im, x, y);
      d += (px_gray(p) <= dark_max) ? 1u : 0u;
      l += (unsigned)px_light(p);
    }
    int c = x / gx;
    if (d) atomicAdd(&sd[c], d);
    atomicAdd(&sl[c], l);
  }
  __syncthreads();
  unsigned *od = pages[page].u32 + out_off + (size_t)cy * ncx;
  unsigned *ol = od + (size_t)ncx * ncy;
  for (int i = threadIdx.x; i < ncx; i += blockDim.x) { od[i] = sd[i]; ol[i] = sl[i]; }
}

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

extern "C" {
void b200k_zero_u32(cudaStream_t st, DPage *pages, int npages, int off, int n) {
  if (n <= 0 || npages <= 0) return;
  dim3 g(min(cdiv(n, 256), 256u), npages);
  k_zero_u32<<<g, 256, 0, st>>>(pages, off, n);
}
int b200k_linesums(cudaStream_t st, DPage *pages, int npages, const DLineJob *jobs_dev,
                   const DLineJob *jobs_host, int njobs, int stat, int lo, int hi, int gray8_aligned, int img_w, int img_h,
                   int want_ink) {
  if (njobs <= 0 || npages <= 0) return 0;
  int made_ink = 0;
  int maxc = 0, maxr = 0, rows_len = 0, rows_any = 0, cols_any = 0;
  for (int i = 0; i < njobs; i++) {
    const DLineJob *j = &jobs_host[i];
    if (j->xa > j->xb || j->ya > j->yb) continue;
    if (j->axis == 0) { cols_any = 1; maxc = max(maxc, j->xb - j->xa + 1); maxr = max(maxr, j->yb - j->ya + 1); }
    else { rows_any = 1; rows_len = max(rows_len, j->yb - j->ya + 1); }
  }
  if (cols_any) {
    /* every column job starts and ends on a 4-byte boundary of a GRAY8 image with aligned rows: four columns per thread */
    bool vec = gray8_aligned && stat != ST_COUNT_GRAY_RANGE;
    int full = -1;
    for (int i = 0; i < njobs && vec; i++) {
      const DLineJob *j = &jobs_host[i];
      if (j->axis != 0 || j->xa > j->xb || j->ya > j->yb) continue;
      if ((j->xa & 3) || ((j->xb - j->xa + 1) & 3)) vec = false;
      if (j->xa == 0 && j->ya == 0 && j->xb == img_w - 1 && j->yb == img_h - 1) full = i;
    }
    bool vec16 = vec;
    for (int i = 0; i < njobs && vec16; i++) {
      const DLineJob *j = &jobs_host[i];
      if (j->axis != 0 || j->xa > j->xb || j->ya > j->yb) continue;
      if ((j->xa & 15) || ((j->xb - j->xa + 1) & 15)) vec16 = false;
    }
    if (vec) {
      /* the ink map rides along when ONE job covers the whole image (and nothing else would write it twice) */
      int ink = want_ink && full >= 0 && njobs == 1 && (img_w & 7) == 0 && stat == ST_GRAY;
      if (vec16) {
        dim3 g(cdiv(maxc, 1024), cdiv(maxr, LS_ROWS16), npages * njobs);
        k_linesum_cols16<<<g, 64, 0, st>>>(pages, jobs_dev, njobs, ink);
      } else {
        dim3 g(cdiv(maxc, 512), cdiv(maxr, LS_ROWS), npages * njobs);
        k_linesum_cols4<<<g, 128, 0, st>>>(pages, jobs_dev, njobs, ink);
      }
      made_ink = ink;
    } else {
      dim3 g(cdiv(maxc, 256), cdiv(maxr, LS_ROWS), npages * njobs);
      k_linesum_cols<<<g, 256, 0, st>>>(pages, jobs_dev, njobs, stat, lo, hi);
    }
  }
  if (rows_any) {
    dim3 g(1, cdiv(rows_len, 8), npages * njobs);
    k_linesum_rows<<<g, 256, 0, st>>>(pages, jobs_dev, njobs, stat, lo, hi);
  }
  return made_ink;
}
void b200k_rect_count(cudaStream_t st, DPage *pages, int npages, const DRect *rects_dev, int nrects,
                      int lo, int hi, int out_off) {
  if (nrects <= 0 || npages <= 0) return;
  dim3 g(cdiv(nrects, 8), npages);
  k_rect_count<<<g, 256, 0, st>>>(pages, rects_dev, nrects, lo, hi, out_off);
}
int b200k_cellstats(cudaStream_t st, DPage *pages, int npages, int gx, int gy, int ncx, int ncy,
                    int dark_max, int out_off) {
  if (npages <= 0 || ncx <= 0 || ncy <= 0) return 0;
  size_t sm = (size_t)ncx * 2 * sizeof(unsigned);
  if (sm > 200 * 1024) return -1;
  if (sm > 48 * 1024) cudaFuncSetAttribute(k_cellstats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  dim3 gr(ncy, npages);
  // a thread takes 16 pixels of a cell row at a time (aligned GRAY8 sheets): no more threads than chunks
  unsigned chunks = cdiv((unsigned)ncx * (unsigned)gx, 16u);
  unsigned threads = min(256u, max(64u, cdiv(chunks, 32u) * 32u));
  k_cellstats<<<gr, threads, sm, st>>>(pages, gx, gy, ncx, ncy, dark_max, out_off);
  return 0;
}
}
