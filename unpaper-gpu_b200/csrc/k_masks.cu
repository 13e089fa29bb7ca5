// k_masks.cu — mask / border detection on top of the band line sums, and the
// small per-page "prep" kernels that turn detector results into blit jobs.
// CPU semantics: reference imageprocess/masks.c:54-215 (detect_edge,
// detect_mask, detect_masks), :222-305 (center_mask, align_mask), :349-488
// (border_to_mask, detect_border_edge, detect_border).  Replaces
// backend_cuda.c:452-583 + cuda_kernels_deskew.cu:202-334 (batched scans with a
// 2000-position cap and one D2H per edge).
#include "common.cuh"
#include "launch.h"

// ---- detect_edge (masks.c:54-100) -----------------------------------------
// One warp per (page, point, side).  side 0=left 1=right (column sums),
// 2=top 3=bottom (row sums).  sums for point i, axis a live at
// u32[sum_off + (i*2+a)*sum_stride ...] and cover the whole image extent of
// that axis, for the bar's (clipped) extent on the other axis.
struct EdgeParams {
  int scan_size[2];   // [0]=width (horizontal scan) [1]=height (vertical scan)
  int scan_depth[2];  // horizontal, vertical (-1 = full)
  int scan_step[2];
  float threshold[2];
  int dir_h, dir_v;
  int sum_off, sum_stride;
};

__global__ void k_detect_edges(DPage *pages, EdgeParams ep) {
  DPage &pg = pages[blockIdx.y];
  int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  int pt = warp_global >> 2, side = warp_global & 3;
  if (pt >= pg.point_count) return;
  bool horiz = side < 2;
  if (horiz ? !ep.dir_h : !ep.dir_v) return;
  const DImg &im = pg.img;
  int ax = horiz ? 0 : 1;
  int size = ep.scan_size[ax];
  int depth = ep.scan_depth[ax];
  int step = ep.scan_step[ax] * ((side & 1) ? 1 : -1);
  float thr = ep.threshold[ax];
  int L = horiz ? im.w : im.h;   // extent along the moving axis
  int D = horiz ? im.h : im.w;   // extent along the bar's long axis
  if (depth == -1) depth = D;
  int o_move = horiz ? pg.px[pt] : pg.py[pt];
  int o_long = horiz ? pg.py[pt] : pg.px[pt];
  // rectangle_from_size(origin + (-size/2, -depth/2), (size, depth))
  int m0 = o_move + (-size / 2), l0 = o_long + (-depth / 2), l1 = l0 + depth - 1;
  int l0c = max(l0, 0), l1c = min(l1, D - 1);
  long long lcount = (long long)abs(l0c - l1c) + 1;
  bool l_ok = l0c <= l1c;
  const unsigned *sums = pg.u32 + ep.sum_off + (size_t)(pt * 2 + ax) * ep.sum_stride;

  unsigned total = 0;   // uint32_t total (masks.c:87)
  int result = -1;
  int kmax = (L + abs(size)) / max(abs(step), 1) + 4;  // beyond this the bar has left the image for good
  for (int base = 0; base < kmax && result < 0; base += 32) {
    int k = base + lane;
    int a0 = m0 + k * step, a1 = a0 + size - 1;
    int a0c = max(a0, 0), a1c = min(a1, L - 1);
    unsigned long long sum = 0;
    if (l_ok && a0c <= a1c)
      for (int a = a0c; a <= a1c; a++) sum += sums[a];
    unsigned long long count = (unsigned long long)((long long)abs(a0c - a1c) + 1) * (unsigned long long)lcount;
    unsigned blackness = (unsigned)(uint8_t)(0xFF - (sum / count));   // blit.c:105
    // inclusive scan of blackness -> running total
    unsigned run = blackness;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, run, o);
      if (lane >= o) run += t;
    }
    unsigned tot = total + run;
    float lhs = (float)(int)blackness;
    float rhs = (thr * (float)tot) / (float)(unsigned)(k + 1);
    bool cont = (lhs >= rhs) && blackness != 0;
    unsigned stop = __ballot_sync(0xffffffffu, !cont);
    if (stop) result = base + (__ffs(stop) - 1) + 1;   // count after the failing step
    total += __shfl_sync(0xffffffffu, run, 31);
  }
  if (lane == 0) {
    if (result < 0) { atomicOr(&pg.error, DERR_EDGE_RUNAWAY); result = kmax; }
    pg.edge_count[pt][side] = result;
  }
}

struct MaskAsmParams {
  int scan_size_w, scan_size_h, step_h, step_v, dir_h, dir_v;
  int min_w, max_w, min_h, max_h;
};

// detect_mask + detect_masks bookkeeping (masks.c:107-205)
__global__ void k_assemble_masks2(DPage *pages, int npages, MaskAsmParams mp) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  int count = 0;
  if (!mp.dir_h && !mp.dir_v) { pg.mask_count = 0; return; }
  for (int i = 0; i < pg.point_count; i++) {
    int ox = pg.px[i], oy = pg.py[i];
    DRect m;
    if (mp.dir_h) {
      m.x0 = ox - mp.step_h * pg.edge_count[i][0] - mp.scan_size_w / 2;
      m.x1 = ox + mp.step_h * pg.edge_count[i][1] + mp.scan_size_w / 2;
    } else { m.x0 = 0; m.x1 = pg.img.w - 1; }
    if (mp.dir_v) {
      m.y0 = oy - mp.step_v * pg.edge_count[i][2] - mp.scan_size_h / 2;
      m.y1 = oy + mp.step_v * pg.edge_count[i][3] + mp.scan_size_h / 2;
    } else { m.y0 = 0; m.y1 = pg.img.h - 1; }
    int w = abs(m.x0 - m.x1) + 1, h = abs(m.y0 - m.y1) + 1;
    int ok = 1;
    if ((mp.min_w != -1 && w < mp.min_w) || (mp.max_w != -1 && w > mp.max_w)) {
      m.x0 = ox - mp.max_w / 2; m.x1 = ox + mp.max_w / 2; ok = 0;
    }
    if ((mp.min_h != -1 && h < mp.min_h) || (mp.max_h != -1 && h > mp.max_h)) {
      m.y0 = oy - mp.max_h / 2; m.y1 = oy + mp.max_h / 2; ok = 0;
    }
    pg.masks[i] = m;
    pg.mask_valid[i] = ok;
    if (!(m.x0 == -1 && m.y0 == -1 && m.x1 == -1 && m.y1 == -1)) count++;
  }
  pg.mask_count = count;
}

// ---- detect_border (masks.c:410-488) -------------------------------------
// Dark counts per row over the outside mask's x range live at
// u32[off + (i*2+1)*stride + (y - ya)], per column over its y range at
// u32[off + (i*2+0)*stride + (x - xa)] — both clipped to the image; `oob_dark`
// says whether out-of-image pixels count (only when abs_black_threshold==255).
struct BorderParams {
  int size_w, size_h, step_h, step_v, thr_h, thr_v, dir_h, dir_v;
  int sum_off, sum_stride, oob_dark;
};

__device__ int border_edge(const DPage &pg, const DRect &om, const unsigned *sums, int lo_clip,
                           int hi_clip, int other_span_total, int other_span_in, int a_start,
                           int a_end, int step, int max_step, int threshold, int oob_dark) {
  // bar covers [a_start..a_end] along the moving axis; sums[] indexed from lo_clip
  unsigned result = 0;
  while (result < (unsigned)max_step) {
    unsigned cnt = 0;
    for (int a = a_start; a <= a_end; a++) {
      if (a >= lo_clip && a <= hi_clip) {
        cnt += sums[a - lo_clip];
        if (oob_dark) cnt += (unsigned)(other_span_total - other_span_in);
      } else if (oob_dark) cnt += (unsigned)other_span_total;
    }
    if (cnt >= (unsigned)threshold) return (int)result;
    a_start += step; a_end += step;
    result += (unsigned)abs(step);
  }
  return 0;
}

__global__ void k_detect_border(DPage *pages, int npages, BorderParams bp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int p = t >> 3, rem = t & 7, i = rem >> 2, side = rem & 3;   // side: 0 left 1 right 2 top 3 bottom
  if (p >= npages) return;
  DPage &pg = pages[p];
  if (i >= pg.outside_count) return;
  const DImg &im = pg.img;
  DRect om = pg.outside[i];
  int mw = abs(om.x0 - om.x1) + 1, mh = abs(om.y0 - om.y1) + 1;
  int xa = max(om.x0, 0), xb = min(om.x1, im.w - 1), ya = max(om.y0, 0), yb = min(om.y1, im.h - 1);
  int xspan_total = om.x1 >= om.x0 ? om.x1 - om.x0 + 1 : 0, xspan_in = xb >= xa ? xb - xa + 1 : 0;
  int yspan_total = om.y1 >= om.y0 ? om.y1 - om.y0 + 1 : 0, yspan_in = yb >= ya ? yb - ya + 1 : 0;
  const unsigned *colsum = pg.u32 + bp.sum_off + (size_t)(i * 2 + 0) * bp.sum_stride;
  const unsigned *rowsum = pg.u32 + bp.sum_off + (size_t)(i * 2 + 1) * bp.sum_stride;
  int v = 0;
  if (side < 2) {
    if (bp.dir_h) {
      if (side == 0) v = border_edge(pg, om, colsum, xa, xb, yspan_total, yspan_in, om.x0, om.x0 + bp.size_w, bp.step_h, mw, bp.thr_h, bp.oob_dark);
      else v = border_edge(pg, om, colsum, xa, xb, yspan_total, yspan_in, om.x1 - bp.size_w, om.x1, -bp.step_h, mw, bp.thr_h, bp.oob_dark);
    }
  } else {
    if (bp.dir_v) {
      if (side == 2) v = border_edge(pg, om, rowsum, ya, yb, xspan_total, xspan_in, om.y0, om.y0 + bp.size_h, bp.step_v, mh, bp.thr_v, bp.oob_dark);
      else v = border_edge(pg, om, rowsum, ya, yb, xspan_total, xspan_in, om.y1 - bp.size_h, om.y1, -bp.step_v, mh, bp.thr_v, bp.oob_dark);
    }
  }
  int *b = &pg.border[i].left;
  int base = side == 0 ? om.x0 : side == 2 ? om.y0 : side == 1 ? im.w - om.x1 : im.h - om.y1;
  // Border field order is left, top, right, bottom
  int field = side == 0 ? 0 : side == 2 ? 1 : side == 1 ? 2 : 3;
  b[field] = base + v;
}

// border_to_mask (masks.c:349-364)
__global__ void k_border_to_mask(DPage *pages, int npages) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int p = t >> 1, i = t & 1;
  if (p >= npages) return;
  DPage &pg = pages[p];
  if (i >= pg.outside_count) return;
  DBorder b = pg.border[i];
  pg.border_mask[i] = DRect{b.left, b.top, pg.img.w - b.right - 1, pg.img.h - b.bottom - 1};
}

// ---- move preparation ----------------------------------------------------
struct MoveJobs { DFillJob *fill_aux; DCopyJob *copy_out; DFillJob *wipe; DCopyJob *copy_in; DFillJob *wipe2; };

__device__ void emit_move(DPage &pg, MoveJobs mj, int p, DRect area, int tx, int ty, int enabled) {
  int w = abs(area.x0 - area.x1) + 1, h = abs(area.y0 - area.y1) + 1;
  DImg aux = pg.aux;
  int bpp = bytes_pp(aux.fmt);
  int pitch = bpp ? ((w * bpp + 15) & ~15) : (((w + 7) / 8 + 15) & ~15);
  // capacity check: aux was allocated for aux.pitch * aux.h bytes
  if ((long long)pitch * h > (long long)aux.pitch * aux.h) { enabled = 0; atomicOr(&pg.error, DERR_UNSUPPORTED); }
  aux.w = w; aux.h = h; aux.pitch = pitch;
  // normalised area, and its part inside the image (what copy_rectangle really copies)
  DRect na = DRect{min(area.x0, area.x1), min(area.y0, area.y1), max(area.x0, area.x1), max(area.y0, area.y1)};
  bool area_inside = na.x0 >= 0 && na.y0 >= 0 && na.x1 < pg.img.w && na.y1 < pg.img.h;
  // the temp only needs its background where the copy below leaves holes
  DFillJob f; f.img = aux; f.r = DRect{0, 0, w - 1, h - 1}; f.c[0] = aux.bg[0]; f.c[1] = aux.bg[1]; f.c[2] = aux.bg[2]; f.pad = 0;
  f.enabled = enabled && !area_inside;
  mj.fill_aux[p] = f;
  DCopyJob c; c.src = pg.img; c.dst = aux; c.area = area; c.tx = 0; c.ty = 0; c.enabled = enabled; c.pad = 0;
  mj.copy_out[p] = c;
  // wipe_rectangle(area) followed by the paste of T = [tx,tx+w) x [ty,ty+h): what the
  // paste overwrites need not be wiped first.  area \ T is an L shape (T is the area
  // shifted): a full-width horizontal strip plus a vertical strip beside T.
  DFillJob wj; wj.img = pg.img; wj.c[0] = pg.img.bg[0]; wj.c[1] = pg.img.bg[1]; wj.c[2] = pg.img.bg[2]; wj.pad = 0;
  DFillJob w2 = wj;
  int dxs = tx - na.x0, dys = ty - na.y0;
  bool overlap = abs(dxs) < w && abs(dys) < h;
  if (!overlap) {
    wj.r = na; wj.enabled = enabled; w2.r = na; w2.enabled = 0;
  } else {
    // rows of the area outside T's row range
    if (dys > 0) wj.r = DRect{na.x0, na.y0, na.x1, ty - 1};
    else wj.r = DRect{na.x0, ty + h, na.x1, na.y1};          // empty when dys == 0
    wj.enabled = enabled && dys != 0;
    // remaining rows (those shared with T): columns of the area outside T's column range
    int ry0 = max(na.y0, ty), ry1 = min(na.y1, ty + h - 1);
    if (dxs > 0) w2.r = DRect{na.x0, ry0, tx - 1, ry1};
    else w2.r = DRect{tx + w, ry0, na.x1, ry1};               // empty when dxs == 0
    w2.enabled = enabled && dxs != 0;
  }
  mj.wipe[p] = wj;
  mj.wipe2[p] = w2;
  DCopyJob ci; ci.src = aux; ci.dst = pg.img; ci.area = DRect{0, 0, w - 1, h - 1}; ci.tx = tx; ci.ty = ty; ci.enabled = enabled; ci.pad = 0;
  mj.copy_in[p] = ci;
}

// center_mask (masks.c:222-249) for mask index i
__global__ void k_prep_center(DPage *pages, int npages, int i, MoveJobs mj) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  int enabled = 0;
  DRect area = DRect{0, 0, 0, 0};
  int tx = 0, ty = 0;
  if (i < pg.mask_count) {
    area = pg.masks[i];
    int w = abs(area.x0 - area.x1) + 1, h = abs(area.y0 - area.y1) + 1;
    tx = pg.px[i] + (-w / 2); ty = pg.py[i] + (-h / 2);
    DRect full = DRect{0, 0, pg.img.w - 1, pg.img.h - 1};
    enabled = pt_in_rect(tx, ty, full) && pt_in_rect(tx + w - 1, ty + h - 1, full);
    pg.centered[i] = enabled;
  }
  emit_move(pg, mj, p, area, tx, ty, enabled);
}

struct AlignParams { int left, top, right, bottom, margin_h, margin_v; };

// align_mask (masks.c:265-305) for border mask i inside outside[i]
__global__ void k_prep_align(DPage *pages, int npages, int i, AlignParams ap, MoveJobs mj) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  int enabled = i < pg.outside_count;
  DRect inside = enabled ? pg.border_mask[i] : DRect{0, 0, 0, 0};
  DRect out = enabled ? pg.outside[i] : DRect{0, 0, 0, 0};
  int w = abs(inside.x0 - inside.x1) + 1, h = abs(inside.y0 - inside.y1) + 1;
  int tx, ty;
  if (ap.left) tx = out.x0 + ap.margin_h;
  else if (ap.right) tx = out.x1 - w - ap.margin_h;
  else tx = (out.x0 + out.x1 - w) / 2;
  if (ap.top) ty = out.y0 + ap.margin_v;
  else if (ap.bottom) ty = out.y1 - h - ap.margin_v;
  else ty = (out.y0 + out.y1 - h) / 2;
  emit_move(pg, mj, p, inside, tx, ty, enabled);
}

// row-independent segment boundaries of a move (see k_move_pass in k_blit.cu)
__device__ void move_finish(DPage &pg, DMove mv) {
  const DImg &im = pg.img;
  int W = im.w, H = im.h, bpp = im.fmt == DF_RGB24 ? 3 : 1, rowbytes = W * bpp;
  int nx0 = min(mv.area.x0, mv.area.x1), nx1 = max(mv.area.x0, mv.area.x1);
  int ny0 = min(mv.area.y0, mv.area.y1), ny1 = max(mv.area.y0, mv.area.y1);
  int w = nx1 - nx0 + 1;
  int ax0 = max(nx0, 0), ax1 = min(nx1, W - 1), ay0 = max(ny0, 0), ay1 = min(ny1, H - 1);
  int wc = ax1 - ax0 + 1, hc = ay1 - ay0 + 1;
  bool have_src = wc > 0 && hc > 0;
  int tb0 = mv.tx * bpp, tb1 = (mv.tx + w) * bpp, tsv = tb0 + (have_src ? wc : 0) * bpp, ab0 = ax0 * bpp, ab1 = (ax1 + 1) * bpp;
  int b[16], n = 0;
#define PUSH(v) { int t_ = (v); b[n++] = t_ < 0 ? 0 : t_ > rowbytes ? rowbytes : t_; }
  PUSH(0); PUSH(rowbytes);
  if (mv.enabled) {
    PUSH(tb0); PUSH(tsv); PUSH(tb1);
    if (have_src) { PUSH(ab0); PUSH(ab1); }
  }
  if (mv.use_masks)
    for (int k = 0; k < pg.outside_count && k < D_MAX_BORDERS; k++) {
      DRect r = pg.border_mask[k];
      int m0 = min(r.x0, r.x1) * bpp, m1 = (max(r.x0, r.x1) + 1) * bpp;
      PUSH(m0); PUSH(m1);
      if (mv.enabled) { PUSH(m0 + tb0 - ab0); PUSH(m1 + tb0 - ab0); }
    }
#undef PUSH
  for (int i = 1; i < n; i++) { int v = b[i], j = i - 1; while (j >= 0 && b[j] > v) { b[j + 1] = b[j]; j--; } b[j + 1] = v; }
  mv.nseg = n - 1; mv.pad = 0;
  for (int i = 0; i < 16; i++) mv.bnd[i] = i < n ? b[i] : rowbytes;
  pg.move = mv;
}

// the same decisions as k_prep_center / k_prep_align, for the one-sweep move (k_move_pass)
__global__ void k_prep_center_move(DPage *pages, int npages, int i) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  DMove mv;
  mv.area = DRect{0, 0, 0, 0}; mv.tx = 0; mv.ty = 0; mv.enabled = 0; mv.use_masks = 0;
  if (i < pg.mask_count) {
    mv.area = pg.masks[i];
    int w = abs(mv.area.x0 - mv.area.x1) + 1, h = abs(mv.area.y0 - mv.area.y1) + 1;
    mv.tx = pg.px[i] + (-w / 2); mv.ty = pg.py[i] + (-h / 2);
    DRect full = DRect{0, 0, pg.img.w - 1, pg.img.h - 1};
    mv.enabled = pt_in_rect(mv.tx, mv.ty, full) && pt_in_rect(mv.tx + w - 1, mv.ty + h - 1, full);
    pg.centered[i] = mv.enabled;
  }
  move_finish(pg, mv);
}
__global__ void k_prep_align_move(DPage *pages, int npages, int i, AlignParams ap, int use_masks) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  DMove mv;
  mv.area = DRect{0, 0, 0, 0}; mv.tx = 0; mv.ty = 0; mv.enabled = 0; mv.use_masks = use_masks;
  if (i < pg.outside_count) {
    DRect inside = pg.border_mask[i], out = pg.outside[i];
    int w = abs(inside.x0 - inside.x1) + 1, h = abs(inside.y0 - inside.y1) + 1;
    if (ap.left) mv.tx = out.x0 + ap.margin_h;
    else if (ap.right) mv.tx = out.x1 - w - ap.margin_h;
    else mv.tx = (out.x0 + out.x1 - w) / 2;
    if (ap.top) mv.ty = out.y0 + ap.margin_v;
    else if (ap.bottom) mv.ty = out.y1 - h - ap.margin_v;
    else mv.ty = (out.y0 + out.y1 - h) / 2;
    mv.area = inside; mv.enabled = 1;
  }
  move_finish(pg, mv);
}
__global__ void k_prep_shift_move(DPage *pages, int npages, int dx, int dy) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  DMove mv;
  mv.area = DRect{0, 0, pg.img.w - 1, pg.img.h - 1}; mv.tx = dx; mv.ty = dy; mv.enabled = 1; mv.use_masks = 0;
  move_finish(pg, mv);
}

// apply_masks job over the detected border masks (sheet_stages.c:474-475)
__global__ void k_prep_border_maskjob(DPage *pages, int npages, DMaskJob *jobs, uint8_t r, uint8_t g, uint8_t b) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  DMaskJob j;
  j.img = pg.img; j.rects = pg.border_mask; j.nrects = pg.outside_count;
  j.c[0] = r; j.c[1] = g; j.c[2] = b; j.pad = 0; j.enabled = pg.outside_count > 0; j.pad2 = 0;
  jobs[p] = j;
}

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

extern "C" {
void b200k_detect_masks(cudaStream_t st, DPage *pages, int npages, int max_points,
                        const int scan_size[2], const int scan_depth[2], const int scan_step[2],
                        const float threshold[2], int dir_h, int dir_v, int sum_off, int sum_stride,
                        int min_w, int max_w, int min_h, int max_h) {
  if (npages <= 0) return;
  EdgeParams ep;
  for (int k = 0; k < 2; k++) { ep.scan_size[k] = scan_size[k]; ep.scan_depth[k] = scan_depth[k]; ep.scan_step[k] = scan_step[k]; ep.threshold[k] = threshold[k]; }
  ep.dir_h = dir_h; ep.dir_v = dir_v; ep.sum_off = sum_off; ep.sum_stride = sum_stride;
  if (max_points > 0 && (dir_h || dir_v)) {
    dim3 g(cdiv(max_points * 4, 4), npages);
    k_detect_edges<<<g, 128, 0, st>>>(pages, ep);
  }
  MaskAsmParams mp = {scan_size[0], scan_size[1], scan_step[0], scan_step[1], dir_h, dir_v, min_w, max_w, min_h, max_h};
  k_assemble_masks2<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, mp);
}
void b200k_detect_border(cudaStream_t st, DPage *pages, int npages, int size_w, int size_h, int step_h,
                         int step_v, int thr_h, int thr_v, int dir_h, int dir_v, int sum_off,
                         int sum_stride, int oob_dark) {
  if (npages <= 0) return;
  BorderParams bp = {size_w, size_h, step_h, step_v, thr_h, thr_v, dir_h, dir_v, sum_off, sum_stride, oob_dark};
  k_detect_border<<<cdiv(npages * 8, 64), 64, 0, st>>>(pages, npages, bp);
  k_border_to_mask<<<cdiv(npages * 2, 64), 64, 0, st>>>(pages, npages);
}
void b200k_prep_center(cudaStream_t st, DPage *pages, int npages, int i, DFillJob *fill_aux,
                       DCopyJob *copy_out, DFillJob *wipe, DCopyJob *copy_in, DFillJob *wipe2) {
  MoveJobs mj = {fill_aux, copy_out, wipe, copy_in, wipe2};
  k_prep_center<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, i, mj);
}
void b200k_prep_align(cudaStream_t st, DPage *pages, int npages, int i, int left, int top, int right,
                      int bottom, int margin_h, int margin_v, DFillJob *fill_aux, DCopyJob *copy_out,
                      DFillJob *wipe, DCopyJob *copy_in, DFillJob *wipe2) {
  MoveJobs mj = {fill_aux, copy_out, wipe, copy_in, wipe2};
  AlignParams ap = {left, top, right, bottom, margin_h, margin_v};
  k_prep_align<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, i, ap, mj);
}
void b200k_prep_center_move(cudaStream_t st, DPage *pages, int npages, int i) {
  if (npages > 0) k_prep_center_move<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, i);
}
void b200k_prep_align_move(cudaStream_t st, DPage *pages, int npages, int i, int left, int top, int right,
                           int bottom, int margin_h, int margin_v, int use_masks) {
  AlignParams ap = {left, top, right, bottom, margin_h, margin_v};
  if (npages > 0) k_prep_align_move<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, i, ap, use_masks);
}
void b200k_prep_shift_move(cudaStream_t st, DPage *pages, int npages, int dx, int dy) {
  if (npages > 0) k_prep_shift_move<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, dx, dy);
}
void b200k_prep_border_maskjob(cudaStream_t st, DPage *pages, int npages, DMaskJob *jobs, int r, int g, int b) {
  k_prep_border_maskjob<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, jobs, (uint8_t)r, (uint8_t)g, (uint8_t)b);
}
}
