// k_blit.cu — rectangle fill / copy / mask painting / mirror / quarter turns.
// Replaces reference imageprocess/cuda_kernels_blit.cu and the OpenCV calls of
// imageprocess/opencv_ops.cpp:91-367 with the CPU backend's exact semantics
// (imageprocess/blit.c:20-85, :291-354; masks.c:311-347).
#include "common.cuh"
#include "launch.h"

#define COPY_ROWS 8

// ---- fill ---------------------------------------------------------------
// One job = one rectangle [x0..x1]x[y0..y1] (empty when inverted), clipped to
// the image here.  blockIdx.z = job.
// byte-run helpers: 16-byte stores on an aligned body, bytes on head and tail
__device__ __forceinline__ void fill_run(uint8_t *dst, int n, const uint8_t pat[3], int phase, int tid, int nthreads) {
  // writes dst[i] = pat[(phase + i) % 3] for i in [0, n)
  int head = (int)((16u - ((unsigned)(uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  for (int i = tid; i < head; i += nthreads) dst[i] = pat[(phase + i) % 3];
  int nvec = (n - head) >> 4;
  uint4 *d4 = (uint4 *)(dst + head);
  if (nvec > 0) {
    unsigned w[3][4];   // the three possible 16-byte vectors of a 3-periodic pattern
#pragma unroll
    for (int v = 0; v < 3; v++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int o = phase + head + 16 * v + 4 * k;
        w[v][k] = (unsigned)pat[o % 3] | ((unsigned)pat[(o + 1) % 3] << 8) | ((unsigned)pat[(o + 2) % 3] << 16) | ((unsigned)pat[(o + 3) % 3] << 24);
      }
    for (int i = tid; i < nvec; i += nthreads) {
      int v = i % 3;
      d4[i] = v == 0 ? make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]) : v == 1 ? make_uint4(w[1][0], w[1][1], w[1][2], w[1][3]) : make_uint4(w[2][0], w[2][1], w[2][2], w[2][3]);
    }
  }
  for (int i = head + (nvec << 4) + tid; i < n; i += nthreads) dst[i] = pat[(phase + i) % 3];
}

__device__ __forceinline__ void copy_run(uint8_t *dst, const uint8_t *src, int n, int tid, int nthreads) {
  int head = (int)((16u - ((unsigned)(uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  for (int i = tid; i < head; i += nthreads) dst[i] = src[i];
  int nvec = (n - head) >> 4;
  uint4 *d4 = (uint4 *)(dst + head);
  const uint8_t *sb = src + head;
  unsigned mis = (unsigned)(uintptr_t)sb & 3u;
  if (((uintptr_t)sb & 15u) == 0) {
    const uint4 *s4 = (const uint4 *)sb;
    for (int i = tid; i < nvec; i += nthreads) d4[i] = s4[i];
  } else if (mis == 0) {
    const unsigned *sw = (const unsigned *)sb;
    for (int i = tid; i < nvec; i += nthreads) d4[i] = make_uint4(sw[4 * i], sw[4 * i + 1], sw[4 * i + 2], sw[4 * i + 3]);
  } else {
    const unsigned *sw = (const unsigned *)(sb - mis);   // aligned words straddling the source bytes
    unsigned sh = mis * 8;
    for (int i = tid; i < nvec; i += nthreads) {
      unsigned a = sw[4 * i], b = sw[4 * i + 1], c = sw[4 * i + 2], d = sw[4 * i + 3], e = sw[4 * i + 4];
      d4[i] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, d, sh), __funnelshift_r(d, e, sh));
    }
  }
  for (int i = head + (nvec << 4) + tid; i < n; i += nthreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(256, 8) k_fill_jobs(const DFillJob *jobs) {
  const DFillJob &j = jobs[blockIdx.z];
  if (!j.enabled) return;
  const DImg &im = j.img;
  int x0 = max(j.r.x0, 0), x1 = min(j.r.x1, im.w - 1);
  int y0 = max(j.r.y0, 0), y1 = min(j.r.y1, im.h - 1);
  if (x0 > x1) return;
  int r = j.c[0], g = j.c[1], b = j.c[2];
  int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (y0 + (int)blockIdx.y > y1) return;
  if (im.fmt == DF_GRAY8 || im.fmt == DF_RGB24) {
    uint8_t v = (uint8_t)((r + g + b) / 3);
    uint8_t pat[3] = {im.fmt == DF_GRAY8 ? v : (uint8_t)r, im.fmt == DF_GRAY8 ? v : (uint8_t)g, im.fmt == DF_GRAY8 ? v : (uint8_t)b};
    int bpp = im.fmt == DF_GRAY8 ? 1 : 3;
    for (int y = y0 + blockIdx.y; y <= y1; y += gridDim.y)
      fill_run(im.data + (size_t)y * im.pitch + (size_t)x0 * bpp, (x1 - x0 + 1) * bpp, pat, 0, tid, nth);
    return;
  }
  for (int y = y0 + blockIdx.y; y <= y1; y += gridDim.y)
    for (int x = x0 + tid; x <= x1; x += nth) px_store(im, x, y, r, g, b);
}

// ---- copy (imageprocess/blit.c:30-80) -------------------------------------
__global__ void __launch_bounds__(256, 8) k_copy_jobs(const DCopyJob *jobs) {
  const DCopyJob &j = jobs[blockIdx.z];
  if (!j.enabled) return;
  const DImg &s = j.src, &d = j.dst;
  // clip_rectangle(source, area): normalise, then clip to the source image
  int ax0 = max(min(j.area.x0, j.area.x1), 0), ax1 = min(max(j.area.x0, j.area.x1), s.w - 1);
  int ay0 = max(min(j.area.y0, j.area.y1), 0), ay1 = min(max(j.area.y0, j.area.y1), s.h - 1);
  int width = ax1 - ax0 + 1, height = ay1 - ay0 + 1;
  if (width <= 0 || height <= 0) return;
  int bpp = bytes_pp(s.fmt);
  bool raw = s.fmt == d.fmt && bpp > 0 && j.tx >= 0 && j.ty >= 0 &&
             j.tx + width <= d.w && j.ty + height <= d.h;
  bool expand = (s.fmt == DF_MONOWHITE || s.fmt == DF_MONOBLACK) && d.fmt == DF_GRAY8 && (ax0 & 7) == 0 &&
                j.tx >= 0 && j.ty >= 0 && j.tx + width <= d.w && j.ty + height <= d.h;
  // one block per COPY_ROWS rows; all its threads stride across each row
  for (int r = blockIdx.y * COPY_ROWS; r < min(height, (int)(blockIdx.y + 1) * COPY_ROWS); r++) {
    int sy = ay0 + r, ty = j.ty + r;
    if (raw) {
      const uint8_t *sp = s.data + (size_t)sy * s.pitch + (size_t)ax0 * bpp;
      uint8_t *dp = d.data + (size_t)ty * d.pitch + (size_t)j.tx * bpp;
      copy_run(dp, sp, width * bpp, threadIdx.x, blockDim.x);
    } else if (expand) {
      // 1-bit page -> gray sheet (get_pixel's 0/255, pixel.c:45-62): a source byte becomes 8 bytes
      const uint8_t *sp = s.data + (size_t)sy * s.pitch + (ax0 >> 3);
      uint8_t *dp = d.data + (size_t)ty * d.pitch + j.tx;
      bool al = ((uintptr_t)dp & 7) == 0;
      unsigned inv = s.fmt == DF_MONOWHITE ? 0xFFu : 0u;
      int nb = width >> 3;
      for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        unsigned v = sp[i] ^ inv;   // bit set = white
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          lo |= (0u - ((v >> (7 - k)) & 1u)) & (0xFFu << (8 * k));
          hi |= (0u - ((v >> (3 - k)) & 1u)) & (0xFFu << (8 * k));
        }
        if (al) *(uint2 *)(dp + 8 * (size_t)i) = make_uint2(lo, hi);
        else
          for (int k = 0; k < 8; k++) dp[8 * (size_t)i + k] = (uint8_t)((k < 4 ? lo >> (8 * k) : hi >> (8 * (k - 4))) & 0xFFu);
      }
      for (int i = (nb << 3) + threadIdx.x; i < width; i += blockDim.x) {
        Px p = px_load(s, ax0 + i, sy);
        px_store(d, j.tx + i, ty, p.r, p.g, p.b);
      }
    } else {
      for (int i = threadIdx.x; i < width; i += blockDim.x) {
        Px p = px_load(s, ax0 + i, sy);
        px_set(d, j.tx + i, ty, p.r, p.g, p.b);
      }
    }
  }
}

// ---- apply_masks (masks.c:311-325): paint what no rectangle covers ---------
__global__ void k_apply_masks(const DMaskJob *jobs) {
  const DMaskJob &j = jobs[blockIdx.z];
  if (!j.enabled || j.nrects <= 0) return;
  const DImg &im = j.img;
  bool bytes = im.fmt == DF_GRAY8 || im.fmt == DF_RGB24;
  int bpp = im.fmt == DF_RGB24 ? 3 : 1;
  uint8_t gv = (uint8_t)((j.c[0] + j.c[1] + j.c[2]) / 3);
  uint8_t pat[3] = {im.fmt == DF_GRAY8 ? gv : j.c[0], im.fmt == DF_GRAY8 ? gv : j.c[1], im.fmt == DF_GRAY8 ? gv : j.c[2]};
  const int SPAN = 64;
  int nspan = (im.w + SPAN - 1) / SPAN;
  if (j.nrects == 1 && bytes) {
    // one rectangle (apply_border / single-page border mask): per row, the part
    // left of it, right of it, or the whole row
    DRect r = j.rects[0];
    int ax = max(min(r.x0, r.x1), 0), bx = min(max(r.x0, r.x1), im.w - 1), ay = min(r.y0, r.y1), by = max(r.y0, r.y1);
    for (int y = blockIdx.y; y < im.h; y += gridDim.y) {
      uint8_t *row = im.data + (size_t)y * im.pitch;
      if (y < ay || y > by || ax > bx) { fill_run(row, im.w * bpp, pat, 0, threadIdx.x, blockDim.x); continue; }
      if (ax > 0) fill_run(row, ax * bpp, pat, 0, threadIdx.x, blockDim.x);
      if (bx < im.w - 1) fill_run(row + (size_t)(bx + 1) * bpp, (im.w - 1 - bx) * bpp, pat, 0, threadIdx.x, blockDim.x);
    }
    return;
  }
  for (int y = blockIdx.y; y < im.h; y += gridDim.y) {
    for (int sp = threadIdx.x; sp < nspan; sp += blockDim.x) {
      int xa = sp * SPAN, xb = min(xa + SPAN, im.w) - 1;
      // classify the span against every rectangle (normalised like point_in_rectangle)
      bool inside_one = false, touches = false;
      for (int k = 0; k < j.nrects; k++) {
        DRect r = j.rects[k];
        int ax = min(r.x0, r.x1), bx = max(r.x0, r.x1), ay = min(r.y0, r.y1), by = max(r.y0, r.y1);
        if (y < ay || y > by || xb < ax || xa > bx) continue;
        touches = true;
        if (xa >= ax && xb <= bx) { inside_one = true; break; }
      }
      if (inside_one) continue;
      if (!touches && bytes) {
        fill_run(im.data + (size_t)y * im.pitch + (size_t)xa * bpp, (xb - xa + 1) * bpp, pat, 0, 0, 1);
        continue;
      }
      for (int x = xa; x <= xb; x++) {
        bool inside = false;
        for (int k = 0; k < j.nrects && !inside; k++) inside = pt_in_rect(x, y, j.rects[k]);
        if (!inside) px_store(im, x, y, j.c[0], j.c[1], j.c[2]);
      }
    }
  }
}

// ---- one-sweep rectangle move (sheet engine) --------------------------------------
// center_mask (masks.c:222-249), align_mask (masks.c:265-305) and shift_image
// (blit.c:360-368) are the same three steps on the CPU:
//     temp(w x h, background) <- copy_rectangle(image, area)      [area clipped to the image, lands at temp(0,0)]
//     wipe_rectangle(image, area, background)
//     copy_rectangle(temp, image, full(temp), target)             [what falls outside the image is dropped]
// = five launches and 5 M bytes of traffic as separate blits.  Here the sheet is rendered
// once from the working buffer into the slot's other buffer as a gather:
//     dst(t) = t in T      ? (its temp pixel was copied ? src'(a0' + (t - T0)) : background)
//            : t in area'  ? background
//            :               src'(t)
// T = [tx,tx+w) x [ty,ty+h), area' = area clipped to the image with origin a0', and
// src'(q) = src(q) or — with use_masks, the apply_masks() of the detected border masks that
// precedes align_mask (sheet_stages.c:474-475) fused in — the mask colour when q lies in no mask.
// The x positions where the category of a byte can change (edges of T, of its part that
// has a source, of area', of the masks — unshifted and shifted by the move) do not depend
// on the row: the prep kernel sorts them once per page into DMove.bnd.  A warp owns a row;
// for every segment between two boundaries it classifies the first byte (row-dependent,
// warp-uniform) and streams the segment with 16-byte stores.
#define MOVE_ROWS 8
__device__ __forceinline__ bool px_in_masks(const DPage &pg, int x, int y) {
  for (int k = 0; k < pg.outside_count; k++) if (pt_in_rect(x, y, pg.border_mask[k])) return true;
  return false;
}

// The segment form (a warp per row): layouts whose rows or buffers are not 16-byte aligned.
__device__ __noinline__ void move_rows_segments(const DPage &pg, int y0, int yend, uint8_t c0, uint8_t c1, uint8_t c2) {
  const DImg &im = pg.img;
  uint8_t *dstb = pg.other;
  const int W = im.w, H = im.h;
  const bool rgb = im.fmt == DF_RGB24;
  const int bpp = rgb ? 3 : 1;
  const DRect area = pg.move.area;
  const int nx0 = min(area.x0, area.x1), nx1 = max(area.x0, area.x1);
  const int ny0 = min(area.y0, area.y1), ny1 = max(area.y0, area.y1);
  const int w = nx1 - nx0 + 1, h = ny1 - ny0 + 1;
  const int ax0 = max(nx0, 0), ax1 = min(nx1, W - 1), ay0 = max(ny0, 0), ay1 = min(ny1, H - 1);   // clip_rectangle
  const int wc = ax1 - ax0 + 1, hc = ay1 - ay0 + 1;
  const bool have_src = wc > 0 && hc > 0;
  const bool en = pg.move.enabled != 0;
  const bool um = pg.move.use_masks != 0 && pg.outside_count > 0;   // apply_masks paints nothing without masks (masks.c:313-315)
  const int tx = pg.move.tx, ty = pg.move.ty;
  const int nseg = pg.move.nseg;
  uint8_t bgp[3], mcp[3];
  {
    uint8_t bgg = (uint8_t)((im.bg[0] + im.bg[1] + im.bg[2]) / 3), mcg = (uint8_t)((c0 + c1 + c2) / 3);
    bgp[0] = rgb ? im.bg[0] : bgg; bgp[1] = rgb ? im.bg[1] : bgg; bgp[2] = rgb ? im.bg[2] : bgg;
    mcp[0] = rgb ? c0 : mcg; mcp[1] = rgb ? c1 : mcg; mcp[2] = rgb ? c2 : mcg;
  }
  const int tb0 = tx * bpp, tb1 = (tx + w) * bpp, tsv = tb0 + (have_src ? wc : 0) * bpp, ab0 = ax0 * bpp, ab1 = (ax1 + 1) * bpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int y = y0 + warp; y < yend; y += nwarps) {
    const uint8_t *srow = im.data + (size_t)y * im.pitch;
    uint8_t *drow = dstb + (size_t)y * im.pitch;
    const bool rowT = en && y >= ty && y < ty + h;
    const int v = y - ty;
    const bool srcrow = rowT && have_src && v < hc;
    const int tse = srcrow ? tsv : tb0;             // [tb0, tse): pasted pixels that have a source, [tse, tb1): background
    const bool rowA = en && have_src && y >= ay0 && y <= ay1;
    const uint8_t *trow = im.data + (size_t)(ay0 + v) * im.pitch + ab0 - tb0;   // source of byte b of T: trow[b]
    for (int s = 0; s < nseg; s++) {
      const int p = pg.move.bnd[s], q = pg.move.bnd[s + 1];
      if (q <= p) continue;
      const uint8_t *sp = NULL;
      int sx = 0, sy = 0;
      if (rowT && p >= tb0 && p < tb1) {
        if (p < tse) { sp = trow; sx = (p - tb0) / bpp + ax0; sy = ay0 + v; }
      } else if (!(rowA && p >= ab0 && p < ab1)) { sp = srow; sx = p / bpp; sy = y; }
      if (!sp) fill_run(drow + p, q - p, bgp, p % 3, lane, 32);
      else if (um && !px_in_masks(pg, sx, sy)) fill_run(drow + p, q - p, mcp, p % 3, lane, 32);
      else copy_run(drow + p, sp + p, q - p, lane, 32);
    }
  }
}

// 16 source bytes at any alignment from two aligned 16-byte loads (`mis` = address & 15 is the same for
// every chunk of a page: rows are a multiple of 16 bytes apart)
__device__ __forceinline__ uint4 sel16_shifted(uint4 A, uint4 B, unsigned wsel, unsigned sh) {
  switch (wsel) {
    case 0: return make_uint4(__funnelshift_r(A.x, A.y, sh), __funnelshift_r(A.y, A.z, sh), __funnelshift_r(A.z, A.w, sh), __funnelshift_r(A.w, B.x, sh));
    case 1: return make_uint4(__funnelshift_r(A.y, A.z, sh), __funnelshift_r(A.z, A.w, sh), __funnelshift_r(A.w, B.x, sh), __funnelshift_r(B.x, B.y, sh));
    case 2: return make_uint4(__funnelshift_r(A.z, A.w, sh), __funnelshift_r(A.w, B.x, sh), __funnelshift_r(B.x, B.y, sh), __funnelshift_r(B.y, B.z, sh));
    default: return make_uint4(__funnelshift_r(A.w, B.x, sh), __funnelshift_r(B.x, B.y, sh), __funnelshift_r(B.y, B.z, sh), __funnelshift_r(B.z, B.w, sh));
  }
}
__device__ __forceinline__ uint4 ld16_shifted(const uint8_t *s, unsigned mis) {
  const uint4 *a = (const uint4 *)(s - mis);
  const uint4 A = __ldg(a);
  if (mis == 0) return A;
  return sel16_shifted(A, __ldg(a + 1), mis >> 2, (mis & 3u) * 8u);
}

// The sweep for 16-byte aligned rows (every sheet buffer of the engine): a thread owns one 16-byte chunk
// column of MOVE_VROWS rows, in batches of eight.  What a chunk column can be along x (inside T, inside the part of T that has a
// source, inside area', inside / outside the masks) is decided once, the row-dependent part is a few
// warp-uniform comparisons per row, and every byte moves as one 16-byte load (two aligned loads + funnel
// shifts for the pasted pixels) and one 16-byte store.  The few chunk columns that contain a boundary of
// DMove.bnd (or the ragged end of the row) are rendered afterwards byte by byte, all threads of the block
// sharing them, from the per-byte definition of the gather above.
#define MOVE_VROWS 64
#define MOVE_VTHREADS 160
__global__ void __launch_bounds__(MOVE_VTHREADS, 6) k_move_pass(DPage *pages, uint8_t c0, uint8_t c1, uint8_t c2) {
  const DPage &pg = pages[blockIdx.z];
  const DImg im = pg.img;                 // by value: no reloads of the descriptor behind the stores below
  uint8_t *const dstb = pg.other;
  const int W = im.w, H = im.h, pitch = im.pitch;
  const int y0 = blockIdx.y * MOVE_VROWS, yend = min(H, y0 + MOVE_VROWS);
  if (y0 >= H) return;
  const bool rgb = im.fmt == DF_RGB24;
  if ((pitch & 15) != 0 || (((uintptr_t)im.data | (uintptr_t)dstb) & 15) != 0 || !(rgb || im.fmt == DF_GRAY8)) {
    move_rows_segments(pg, y0, yend, c0, c1, c2);
    return;
  }
  const int bpp = rgb ? 3 : 1, rowbytes = W * bpp;
  const DMove &mv = pg.move;
  const DRect area = mv.area;
  const int nseg = mv.nseg;
  const int nx0 = min(area.x0, area.x1), nx1 = max(area.x0, area.x1);
  const int ny0 = min(area.y0, area.y1), ny1 = max(area.y0, area.y1);
  const int w = nx1 - nx0 + 1, h = ny1 - ny0 + 1;
  const int ax0 = max(nx0, 0), ax1 = min(nx1, W - 1), ay0 = max(ny0, 0), ay1 = min(ny1, H - 1);   // clip_rectangle
  const int wc = ax1 - ax0 + 1, hc = ay1 - ay0 + 1;
  const bool have_src = wc > 0 && hc > 0;
  const bool en = mv.enabled != 0;
  const int nmask = mv.use_masks != 0 ? min(pg.outside_count, D_MAX_BORDERS) : 0;   // apply_masks paints nothing without masks (masks.c:313-315)
  static_assert(D_MAX_BORDERS == 2, "the two border masks live in registers");
  // normalised like point_in_rectangle; an unused slot is empty
  const DRect r0 = pg.border_mask[0], r1 = pg.border_mask[1];
  const int m0xa = nmask > 0 ? min(r0.x0, r0.x1) : 1, m0xb = nmask > 0 ? max(r0.x0, r0.x1) : 0, m0ya = min(r0.y0, r0.y1), m0yb = max(r0.y0, r0.y1);
  const int m1xa = nmask > 1 ? min(r1.x0, r1.x1) : 1, m1xb = nmask > 1 ? max(r1.x0, r1.x1) : 0, m1ya = min(r1.y0, r1.y1), m1yb = max(r1.y0, r1.y1);
#define MV_IN_MASKS(x, y) (((x) >= m0xa && (x) <= m0xb && (y) >= m0ya && (y) <= m0yb) || ((x) >= m1xa && (x) <= m1xb && (y) >= m1ya && (y) <= m1yb))
  const int tx = mv.tx, ty = mv.ty;
  const unsigned bgg = (unsigned)((im.bg[0] + im.bg[1] + im.bg[2]) / 3) & 0xFFu, mcg = (unsigned)((c0 + c1 + c2) / 3) & 0xFFu;
  const unsigned bg0 = rgb ? im.bg[0] : bgg, bg1 = rgb ? im.bg[1] : bgg, bg2 = rgb ? im.bg[2] : bgg;
  const unsigned mc0 = rgb ? c0 : mcg, mc1 = rgb ? c1 : mcg, mc2 = rgb ? c2 : mcg;
  const int tb0 = tx * bpp, tb1 = (tx + w) * bpp, tsv = tb0 + (have_src ? wc : 0) * bpp, ab0 = ax0 * bpp, ab1 = (ax1 + 1) * bpp;
  const unsigned misT = (unsigned)(ab0 - tb0) & 15u;   // alignment of a pasted chunk's source
  // rows at which the kind of a chunk column can change: the eight rows of a batch all behave alike when none of
  // them lies inside the batch (one bit per batch)
  unsigned same_mask = 0;
  for (int bi = 0, ya = y0; ya < yend; bi++, ya += 8) {
    const int yb = min(ya + 8, yend);
    bool sm = true;
#define MV_BRK(yy) sm = sm && ((yy) <= ya || (yy) >= yb)
    if (en) { MV_BRK(ty); MV_BRK(ty + h); if (have_src) { MV_BRK(ty + hc); MV_BRK(ay0); MV_BRK(ay1 + 1); } }
    if (nmask > 0) { MV_BRK(m0ya); MV_BRK(m0yb + 1); if (en) { MV_BRK(m0ya + ty - ay0); MV_BRK(m0yb + 1 + ty - ay0); } }
    if (nmask > 1) { MV_BRK(m1ya); MV_BRK(m1yb + 1); if (en) { MV_BRK(m1ya + ty - ay0); MV_BRK(m1yb + 1 + ty - ay0); } }
#undef MV_BRK
    same_mask |= (sm ? 1u : 0u) << bi;
  }
  static_assert(MOVE_VROWS <= 8 * 32, "one bit per batch of eight rows");
  __shared__ int s_nb, s_bc[20];
  if (threadIdx.x == 0) s_nb = 0;
  __syncthreads();
  const int nch = (rowbytes + 15) >> 4;
  for (int c = threadIdx.x; c < nch; c += MOVE_VTHREADS) {
    const int p = c << 4;
    bool uni = p + 16 <= rowbytes;
    for (int s = 1; s < nseg; s++) { const int bs = mv.bnd[s]; uni = uni && !(bs > p && bs < p + 16); }
    if (!uni) { s_bc[atomicAdd(&s_nb, 1)] = c; continue; }
    const bool inT = en && p >= tb0 && p < tb1, inTs = p < tsv;
    const bool inA = en && have_src && p >= ab0 && p < ab1;
    // the 16 bytes of a 3-periodic pattern that start at byte p (p % 3 == c % 3)
    const int ph = c % 3;
    uint4 bgv, mcv;
    {
      const unsigned b0 = bg0 | (bg1 << 8) | (bg2 << 16) | (bg0 << 24), b1 = bg1 | (bg2 << 8) | (bg0 << 16) | (bg1 << 24), b2 = bg2 | (bg0 << 8) | (bg1 << 16) | (bg2 << 24);
      const unsigned m0 = mc0 | (mc1 << 8) | (mc2 << 16) | (mc0 << 24), m1 = mc1 | (mc2 << 8) | (mc0 << 16) | (mc1 << 24), m2 = mc2 | (mc0 << 8) | (mc1 << 16) | (mc2 << 24);
      bgv = ph == 0 ? make_uint4(b0, b1, b2, b0) : ph == 1 ? make_uint4(b1, b2, b0, b1) : make_uint4(b2, b0, b1, b2);
      mcv = ph == 0 ? make_uint4(m0, m1, m2, m0) : ph == 1 ? make_uint4(m1, m2, m0, m1) : make_uint4(m2, m0, m1, m2);
    }
    const int sxT = (p - tb0) / bpp + ax0, sxO = p / bpp;
    for (int bi = 0, ya = y0; ya < yend; bi++, ya += 8) {
      const int rows = min(8, yend - ya);
      const uint8_t *sown = im.data + (size_t)ya * pitch + p;
      const uint8_t *spst = im.data + ((ptrdiff_t)(ay0 + ya - ty) * pitch + ab0 - tb0 + p);
      uint8_t *d = dstb + (size_t)ya * pitch + p;
      if ((same_mask >> bi) & 1u) {
        // no row of this batch changes what a chunk column is: decide at its first row, then move the rows with
        // all their loads in flight
        const int v = ya - ty;
        int kind;                               // 0 fill, 1 own pixels, 2 pasted pixels
        int sx, sy;
        if (inT && ya >= ty && v < h) {
          if (inTs && have_src && v < hc) { kind = 2; sx = sxT; sy = ay0 + v; } else { kind = 0; sx = sy = 0; }
        } else if (inA && ya >= ay0 && ya <= ay1) { kind = 0; sx = sy = 0; }
        else { kind = 1; sx = sxO; sy = ya; }
        uint4 fv = bgv;
        if (kind != 0 && nmask > 0 && !MV_IN_MASKS(sx, sy)) { kind = 0; fv = mcv; }
        if (kind == 0) {
          for (int r = 0; r < rows; r++, d += pitch) *(uint4 *)d = fv;
        } else if (kind == 1 || misT == 0) {
          const uint8_t *sp = kind == 1 ? sown : spst;
          if (rows == 8) {
            uint4 t[8];
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = __ldg((const uint4 *)(sp + (size_t)k * pitch));
#pragma unroll
            for (int k = 0; k < 8; k++) *(uint4 *)(d + (size_t)k * pitch) = t[k];
          } else
            for (int r = 0; r < rows; r++, sp += pitch, d += pitch) *(uint4 *)d = __ldg((const uint4 *)sp);
        } else {
          const uint8_t *sp = spst - misT;      // the aligned chunk that holds the first source byte
          const unsigned sh = (misT & 3u) * 8u, wsel = misT >> 2;
          int r = 0;
          for (; r + 4 <= rows; r += 4, sp += 4 * (size_t)pitch, d += 4 * (size_t)pitch) {
            uint4 A[4], B[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { A[k] = __ldg((const uint4 *)(sp + (size_t)k * pitch)); B[k] = __ldg((const uint4 *)(sp + (size_t)k * pitch) + 1); }
#pragma unroll
            for (int k = 0; k < 4; k++) *(uint4 *)(d + (size_t)k * pitch) = sel16_shifted(A[k], B[k], wsel, sh);
          }
          for (; r < rows; r++, sp += pitch, d += pitch) *(uint4 *)d = sel16_shifted(__ldg((const uint4 *)sp), __ldg((const uint4 *)sp + 1), wsel, sh);
        }
        continue;
      }
      for (int y = ya; y < ya + rows; y++, sown += pitch, spst += pitch, d += pitch) {
        const int v = y - ty;
        const bool rowT = y >= ty && v < h;
        uint4 val;
        int sx, sy;
        bool fill = false;
        if (inT && rowT) {
          if (inTs && have_src && v < hc) { val = ld16_shifted(spst, misT); sx = sxT; sy = ay0 + v; }
          else fill = true;
        } else if (inA && y >= ay0 && y <= ay1) fill = true;
        else { val = __ldg((const uint4 *)sown); sx = sxO; sy = y; }
        if (fill) val = bgv;
        else if (nmask > 0 && !MV_IN_MASKS(sx, sy)) val = mcv;
        *(uint4 *)d = val;
      }
    }
  }
  __syncthreads();
  // chunk columns with a boundary inside: byte b of row y straight from the definition
  const int nb = s_nb, rows = yend - y0;
  for (int i = threadIdx.x; i < nb * rows * 16; i += MOVE_VTHREADS) {
    const int b = (s_bc[i / (rows * 16)] << 4) + (i & 15), y = y0 + (i >> 4) % rows;
    if (b >= rowbytes) continue;
    const int v = y - ty;
    const uint8_t *sp = NULL;
    int sx = 0, sy = 0;
    if (en && y >= ty && v < h && b >= tb0 && b < tb1) {
      if (have_src && v < hc && b < tsv) { sp = im.data + ((ptrdiff_t)(ay0 + v) * pitch + ab0 - tb0 + b); sx = (b - tb0) / bpp + ax0; sy = ay0 + v; }
    } else if (!(en && have_src && y >= ay0 && y <= ay1 && b >= ab0 && b < ab1)) { sp = im.data + (size_t)y * pitch + b; sx = b / bpp; sy = y; }
    const int ph = b % 3;
    unsigned val;
    if (!sp) val = ph == 0 ? bg0 : ph == 1 ? bg1 : bg2;
    else {
      val = (nmask == 0 || MV_IN_MASKS(sx, sy)) ? *sp : ph == 0 ? mc0 : ph == 1 ? mc1 : mc2;
    }
    dstb[(size_t)y * pitch + b] = (uint8_t)val;
  }
#undef MV_IN_MASKS
}

// ---- mirror (blit.c:320-354): disjoint pixel pairs swapped in place --------
__global__ void k_mirror(DImg im, int dir_h, int dir_v) {
  int y = blockIdx.y;
  int ymax = dir_v ? (im.h - 1) / 2 : im.h - 1;
  if (y > ymax) return;
  int yy = dir_v ? im.h - y - 1 : y;
  int xmax = im.w - 1;
  if (dir_h && (!dir_v || y == yy)) xmax = (im.w - 1) / 2;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x <= xmax; x += gridDim.x * blockDim.x) {
    int xx = dir_h ? im.w - x - 1 : x;
    Px p1 = px_load(im, x, y), p2 = px_load(im, xx, yy);
    px_store(im, x, y, p2.r, p2.g, p2.b);
    px_store(im, xx, yy, p1.r, p1.g, p1.b);
  }
}

// the same for every sheet of a group (grid z = sheet)
__global__ void k_mirror_pages(DPage *pages, int dir_h, int dir_v) {
  const DImg &im = pages[blockIdx.z].img;
  int ymax = dir_v ? (im.h - 1) / 2 : im.h - 1;
  for (int y = blockIdx.y; y <= ymax; y += gridDim.y) {
    int yy = dir_v ? im.h - y - 1 : y;
    int xmax = im.w - 1;
    if (dir_h && (!dir_v || y == yy)) xmax = (im.w - 1) / 2;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x <= xmax; x += gridDim.x * blockDim.x) {
      int xx = dir_h ? im.w - x - 1 : x;
      Px p1 = px_load(im, x, y), p2 = px_load(im, xx, yy);
      px_store(im, x, y, p2.r, p2.g, p2.b);
      px_store(im, xx, yy, p1.r, p1.g, p1.b);
    }
  }
}

// ---- flip_rotate_90 (blit.c:291-314) -------------------------------------
__global__ void k_rotate90(DImg src, DImg dst, int dir) {
  int y = blockIdx.y;
  if (y >= src.h) return;
  int xx = ((dir > 0) ? src.h - 1 : 0) - y * dir;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < src.w; x += gridDim.x * blockDim.x) {
    int yy = ((dir < 0) ? src.w - 1 : 0) + x * dir;
    Px p = px_load(src, x, y);
    px_set(dst, xx, yy, p.r, p.g, p.b);
  }
}

// ---- sheet-engine forms of the size-changing operations (one launch per group) ---------
// flip_rotate_90 of nimages equally sized images stored `*_stride` bytes apart (the input
// pages of a group: options->pre_rotate, sheet_stages.c:134-137)
__global__ void k_rotate90_batch(DImg src, DImg dst, int dir, size_t src_stride, size_t dst_stride) {
  src.data += (size_t)blockIdx.z * src_stride;
  dst.data += (size_t)blockIdx.z * dst_stride;
  for (int y = blockIdx.y; y < src.h; y += gridDim.y) {
    int xx = ((dir > 0) ? src.h - 1 : 0) - y * dir;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < src.w; x += gridDim.x * blockDim.x) {
      int yy = ((dir < 0) ? src.w - 1 : 0) + x * dir;
      Px p = px_load(src, x, y);
      px_set(dst, xx, yy, p.r, p.g, p.b);
    }
  }
}
// flip_rotate_90 of every page's working sheet into its other buffer (options->post_rotate, :511-514)
__global__ void k_rotate90_pages(DPage *pages, int dir, int dw, int dh, int dpitch) {
  const DPage &pg = pages[blockIdx.z];
  DImg src = pg.img, dst = pg.img;
  dst.data = pg.other; dst.w = dw; dst.h = dh; dst.pitch = dpitch;
  for (int y = blockIdx.y; y < src.h; y += gridDim.y) {
    int xx = ((dir > 0) ? src.h - 1 : 0) - y * dir;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < src.w; x += gridDim.x * blockDim.x) {
      int yy = ((dir < 0) ? src.w - 1 : 0) + x * dir;
      Px p = px_load(src, x, y);
      px_set(dst, xx, yy, p.r, p.g, p.b);
    }
  }
}
// resize_and_replace's second half (blit.c:279-281): a background sheet of the new size with the
// (already stretched) working sheet centred on it — center_image, blit.c:175-207
__global__ void k_center_pages(DPage *pages, int dw, int dh, int dpitch) {
  const DPage &pg = pages[blockIdx.z];
  DImg src = pg.img, dst = pg.img;
  dst.data = pg.other; dst.w = dw; dst.h = dh; dst.pitch = dpitch;
  int sx0 = 0, sy0 = 0, sw = src.w, sh = src.h, tx0 = 0, ty0 = 0;
  if (sw <= dw) tx0 = (dw - sw) / 2; else { sx0 = (sw - dw) / 2; sw = dw; }
  if (sh <= dh) ty0 = (dh - sh) / 2; else { sy0 = (sh - dh) / 2; sh = dh; }
  for (int y = blockIdx.y; y < dh; y += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < dw; x += gridDim.x * blockDim.x) {
      int u = x - tx0, v = y - ty0;
      if (u >= 0 && u < sw && v >= 0 && v < sh) { Px p = px_load(src, sx0 + u, sy0 + v); px_store(dst, x, y, p.r, p.g, p.b); }
      else px_store(dst, x, y, dst.bg[0], dst.bg[1], dst.bg[2]);
    }
}

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

extern "C" {
void b200k_rotate90_batch(cudaStream_t st, DImg src, DImg dst, int dir, int nimages, size_t src_stride, size_t dst_stride) {
  if (nimages <= 0 || src.w <= 0 || src.h <= 0) return;
  dim3 g(min(cdiv(src.w, 256), 16u), min((unsigned)src.h, 1024u), nimages);
  k_rotate90_batch<<<g, 256, 0, st>>>(src, dst, dir, src_stride, dst_stride);
}
void b200k_rotate90_pages(cudaStream_t st, DPage *pages, int npages, int sw, int sh, int dir, int dpitch) {
  if (npages <= 0 || sw <= 0 || sh <= 0) return;
  dim3 g(min(cdiv(sw, 256), 16u), min((unsigned)sh, 1024u), npages);
  k_rotate90_pages<<<g, 256, 0, st>>>(pages, dir, sh, sw, dpitch);
}
void b200k_center_pages(cudaStream_t st, DPage *pages, int npages, int dw, int dh, int dpitch) {
  if (npages <= 0 || dw <= 0 || dh <= 0) return;
  dim3 g(min(cdiv(dw, 256), 16u), min((unsigned)dh, 1024u), npages);
  k_center_pages<<<g, 256, 0, st>>>(pages, dw, dh, dpitch);
}
void b200k_fill_jobs(cudaStream_t st, const DFillJob *jobs, int njobs, int maxw, int maxh) {
  if (njobs <= 0 || maxw <= 0 || maxh <= 0) return;
  // one 256-thread block covers a row (16 B per thread per step); ~8 rows per block
  dim3 g(1, min(cdiv((unsigned)maxh, 8u), 1024u), njobs);
  k_fill_jobs<<<g, 256, 0, st>>>(jobs);
}
void b200k_copy_jobs(cudaStream_t st, const DCopyJob *jobs, int njobs, int maxw_bytes, int maxh) {
  if (njobs <= 0 || maxw_bytes <= 0 || maxh <= 0) return;
  (void)maxw_bytes;
  dim3 g(1, cdiv(maxh, COPY_ROWS), njobs);
  k_copy_jobs<<<g, 256, 0, st>>>(jobs);
}
void b200k_move_pass(cudaStream_t st, DPage *pages, int npages, int maxw_bytes, int maxh, int mc_r, int mc_g, int mc_b) {
  if (npages <= 0 || maxw_bytes <= 0 || maxh <= 0) return;
  dim3 g(1, cdiv(maxh, MOVE_VROWS), npages);
  k_move_pass<<<g, MOVE_VTHREADS, 0, st>>>(pages, (uint8_t)mc_r, (uint8_t)mc_g, (uint8_t)mc_b);
}
void b200k_apply_masks(cudaStream_t st, const DMaskJob *jobs, int njobs, int maxw, int maxh) {
  if (njobs <= 0 || maxw <= 0 || maxh <= 0) return;
  dim3 g(1, min(cdiv((unsigned)maxh, 4u), 1024u), njobs);
  k_apply_masks<<<g, 64, 0, st>>>(jobs);
}
void b200k_mirror_pages(cudaStream_t st, DPage *pages, int npages, int maxw, int maxh, int dir_h, int dir_v) {
  if (npages <= 0 || maxw <= 0 || maxh <= 0 || (!dir_h && !dir_v)) return;
  dim3 g(min(cdiv((unsigned)maxw, 256u), 8u), min((unsigned)maxh, 2048u), npages);
  k_mirror_pages<<<g, 256, 0, st>>>(pages, dir_h, dir_v);
}
void b200k_mirror(cudaStream_t st, DImg im, int dir_h, int dir_v) {
  if (im.w <= 0 || im.h <= 0) return;
  dim3 g(min(cdiv(im.w, 256), 64u), im.h, 1);
  k_mirror<<<g, 256, 0, st>>>(im, dir_h, dir_v);
}
void b200k_rotate90(cudaStream_t st, DImg src, DImg dst, int dir) {
  if (src.w <= 0 || src.h <= 0) return;
  dim3 g(min(cdiv(src.w, 256), 64u), src.h, 1);
  k_rotate90<<<g, 256, 0, st>>>(src, dst, dir);
}
}
