// k_filters.cu — blackfilter, noisefilter, blurfilter, grayfilter with the
// reference CPU backend's sequential semantics reproduced exactly
// (reference imageprocess/filters.c:49-402, imageprocess/fill.c:16-107).
// Replaces backend_cuda_filters.c + cuda_kernels_filters.cu + the OpenCV CCL /
// NPP-integral paths of opencv_bridge.cpp (which are NOT CPU-exact).
#include <stdio.h>
#include "common.cuh"
#include "launch.h"

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

/* =========================================================================
 * blackfilter
 *
 * The scan positions (filters.c:60-103) depend only on geometry, so the host
 * enumerates them once, in order, dropping excluded ones.  Flood fills only
 * ever whiten pixels, so an area's darkness can only fall while the filter
 * runs: "hit on the untouched image" is a superset of the true hits.  One CTA
 * per page walks that candidate list in order, re-measures candidates after a
 * fill, and emulates flood_fill()'s recursion (fill.c:81-107) with an explicit
 * frame stack so that the painted set — including fill_line's overruns —
 * matches the CPU's order-dependent result.
 *
 * The emulation is a chain of dependent steps, so it is written for latency:
 *  - every step is BF_ROUND pixels wide (64 chunks of 32; the per-chunk ballots
 *    meet in shared memory and every thread replays the same scalar decision);
 *  - all per-fill state (the four line lengths and counters, the top frame, the
 *    page description) lives in registers: no dynamically indexed local arrays,
 *    no structs behind references that byte stores could alias;
 *  - vertical lines touch one 32-byte sector per pixel and consecutive frames
 *    walk adjacent columns, so the CTA keeps a 32-column strip of the whole
 *    page height in shared memory (A4: 3508 x 32 B = 110 KB), write-back; it
 *    moves when a frame opens outside it and is flushed at the end.
 * ====================================================================== */
#define FF_U 4
#define BF_WARPS 16
#define BF_THREADS (BF_WARPS * 32)
#define BF_CH (BF_WARPS * FF_U)
#define BF_CH_LOG2 6
#define BF_ROUND (BF_CH * 32)
static_assert(BF_CH == 64 && (1 << BF_CH_LOG2) == BF_CH, "two chunks per lane in the replay");

#ifdef BF_STATS
__device__ unsigned long long g_bfs[16];
#define BFS_ADD(i, v) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_bfs[i] += (v); } while (0)
#else
#define BFS_ADD(i, v) do {} while (0)
#endif

struct BfShared {
  unsigned M[BF_CH];
  unsigned I[BF_CH];
  unsigned wcnt[BF_WARPS];
};

// The page as the flood fill sees it; passed by value so that it stays in registers.
struct BfCtx {
  uint8_t *data;       // GRAY8 pixels (gen == NULL)
  uint8_t *sbuf;       // column strip, 32 bytes per row
  const DImg *gen;     // other formats: generic accessors on the device record
  int pitch, w, h;
  int sx0;             // first column of the strip (multiple of 32), -64: nothing loaded
  int son;             // strip usable for this page
};

__device__ __noinline__ int ff_gray_generic(const DImg *im, int x, int y) { return px_gray(px_get(*im, x, y)); }
__device__ __noinline__ void ff_paint_generic(const DImg *im, int x, int y) { px_store(*im, x, y, 255, 255, 255); }

__device__ __forceinline__ bool cx_in(const BfCtx &c, int x, int y) { return (unsigned)x < (unsigned)c.w && (unsigned)y < (unsigned)c.h; }
__device__ __forceinline__ int ff_gray_at(const BfCtx &c, int x, int y) {   // outside the image reads as white
  if (c.gen) return ff_gray_generic(c.gen, x, y);
  if (!cx_in(c, x, y)) return 255;
  unsigned dx = (unsigned)(x - c.sx0);
  if (dx < 32u) return (int)c.sbuf[y * 32 + (int)dx];
  return (int)c.data[(size_t)y * c.pitch + x];
}
__device__ __forceinline__ void ff_paint(const BfCtx &c, int x, int y) {    // (x, y) inside the image
  if (c.gen) { ff_paint_generic(c.gen, x, y); return; }
  unsigned dx = (unsigned)(x - c.sx0);
  if (dx < 32u) c.sbuf[y * 32 + (int)dx] = 255;
  else c.data[(size_t)y * c.pitch + x] = 255;
}
// block-cooperative; callers put barriers around it
__device__ __forceinline__ void strip_copy(const BfCtx &c, bool to_global) {
  int nv = min(32, c.pitch - c.sx0) >> 4;   // 16-byte pieces per row inside the row's pitch
  for (int i = threadIdx.x; i < c.h * nv; i += blockDim.x) {
    int row = i / nv, part = i - row * nv;
    uint4 *g = (uint4 *)(c.data + (size_t)row * c.pitch + c.sx0 + part * 16);
    uint4 *l = (uint4 *)(c.sbuf + row * 32 + part * 16);
    if (to_global) *g = *l; else *l = *g;
  }
}

// candidate `idx` of a frame (fill.c:45-79: the two neighbours of every painted
// pixel of the left, top, right, bottom line, nearest first)
__device__ __forceinline__ void ff_cand(int cx, int cy, int L, int T, int R, unsigned idx, int &x, int &y) {
  unsigned d = (idx >> 1) + 1u, sub = idx & 1u;
  unsigned nL = 2u * L, nT = 2u * T, nR = 2u * R;
  if (idx < nL) { x = cx - (int)d; y = cy + (sub ? -1 : 1); return; }
  idx -= nL; d = (idx >> 1) + 1u;
  if (idx < nT) { x = cx + (sub ? -1 : 1); y = cy - (int)d; return; }
  idx -= nT; d = (idx >> 1) + 1u;
  if (idx < nR) { x = cx + (int)d; y = cy + (sub ? -1 : 1); return; }
  idx -= nR; d = (idx >> 1) + 1u;
  x = cx + (sub ? -1 : 1); y = cy + (int)d;
}

#define SEL4(i, a, b, c, d) ((i) == 0 ? (a) : (i) == 1 ? (b) : (i) == 2 ? (c) : (d))

// flood_fill(x, y) (fill.c:81-107) for a pixel that is known to match, run to
// completion.  Returns the strip position (it may have moved).
__device__ __noinline__ int ff_flood(BfCtx c, unsigned long long *stack, int stack_cap, unsigned *err, BfShared &sh,
                                     int x, int y, int lo, int hi, unsigned long long intensity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int sp = 0;
  int t_cx = 0, t_cy = 0, t_L = 0, t_T = 0, t_R = 0, t_B = 0;   // the top frame
  unsigned t_cur = 0;
  int fx = x, fy = y;
  bool pending = true;
  for (;;) {
    if (pending) {
      // ---- open a frame: paint the centre and its cross (fill.c:85-95) ----
      if (sp >= stack_cap) { if (threadIdx.x == 0) atomicOr(err, DERR_STACK_OVERFLOW); break; }
      if (sp > 0 && threadIdx.x == 0) {   // spill the caller's frame
        unsigned long long *s = stack + (size_t)(sp - 1) * 4;
        s[0] = (unsigned)t_cx | ((unsigned long long)(unsigned)t_cy << 32);
        s[1] = (unsigned long long)(unsigned)t_L | ((unsigned long long)(unsigned)t_T << 32);
        s[2] = (unsigned long long)(unsigned)t_R | ((unsigned long long)(unsigned)t_B << 32);
        s[3] = t_cur;
      }
      if (c.son && (unsigned)(fx - c.sx0) >= 32u) {   // bring the strip to the column this frame walks
        BFS_ADD(14, 1);
        __syncthreads();
        if (c.sx0 >= 0) strip_copy(c, true);
        __syncthreads();
        c.sx0 = fx & ~31;
        strip_copy(c, false);
      }
      if (threadIdx.x == 0) ff_paint(c, fx, fy);
      __syncthreads();
      BFS_ADD(1, 1);
      // fill_line (fill.c:16-43) x 4: the four lines of the cross touch disjoint pixels,
      // so they advance together and share the rounds.  d: 0 left, 1 up, 2 right, 3 down
      int d0 = 0, d1 = 0, d2 = 0, d3 = 0;
      unsigned long long c0 = 1, c1 = 1, c2 = 1, c3 = 1;
      unsigned active = intensity == 0 ? 0u : 0xFu;   // intensity 0: every line stops on its first pixel
      while (active) {
        BFS_ADD(0, 1);
        const int na = __popc(active);
        const int shift = BF_CH_LOG2 - (na == 1 ? 0 : na == 2 ? 1 : 2);   // log2(chunks per line)
        const int per = 1 << shift;
        unsigned packed = 0;   // direction of the li-th active line, 2 bits each
        {
          int q = 0;
#pragma unroll
          for (int d = 0; d < 4; d++) if ((active >> d) & 1u) { packed |= (unsigned)d << (2 * q); q++; }
        }
        bool full = true;
#pragma unroll
        for (int u = 0; u < FF_U; u++) {
          int ch = warp * FF_U + u, li = ch >> shift;
          unsigned M = 0, I = 0;
          if (li < na) {
            int d = (packed >> (2 * li)) & 3;
            int dist = SEL4(d, d0, d1, d2, d3);
            int s = dist + ((ch & (per - 1)) << 5) + lane + 1;
            int dx = (d & 1) ? 0 : d - 1, dy = (d & 1) ? d - 2 : 0;
            int qx = fx + s * dx, qy = fy + s * dy;
            int g = ff_gray_at(c, qx, qy);
            M = __ballot_sync(0xffffffffu, g >= lo && g <= hi);
            I = __ballot_sync(0xffffffffu, cx_in(c, qx, qy));
            if ((M & I) != 0xffffffffu) full = false;
          }
          if (lane == 0) { sh.M[ch] = M; sh.I[ch] = I; }
        }
        // barrier + "did every evaluated chunk match in full?" in one step
        bool all_full = __syncthreads_and(full) != 0;
        int t0 = 0, t1 = 0, t2 = 0, t3 = 0;   // pixels painted in this round, per active line
        unsigned still = active;
        if (all_full) {
          // nothing stops: every active line advances by all its chunks
          t0 = t1 = t2 = t3 = per * 32;
          c0 = c1 = c2 = c3 = intensity;
        } else {
          // which chunks did not match in full (two chunks per lane)
          unsigned Ma = sh.M[lane], Ia = sh.I[lane], Mb = sh.M[32 + lane], Ib = sh.I[32 + lane];
          unsigned long long nf = (unsigned long long)__ballot_sync(0xffffffffu, (Ma & Ia) != 0xffffffffu) |
                                  ((unsigned long long)__ballot_sync(0xffffffffu, (Mb & Ib) != 0xffffffffu) << 32);
#pragma unroll
          for (int li = 0; li < 4; li++) {
            if (li >= na) break;
            int d = (packed >> (2 * li)) & 3;
            unsigned long long cnt = SEL4(d, c0, c1, c2, c3);
            int base = li << shift;
            unsigned long long lm = nf >> base;
            if (per < 64) lm &= (1ull << per) - 1ull;
            int k = 0, tot = 0;
            bool stop = false;
            while (k < per && !stop) {
              unsigned long long rest = lm >> k;
              if (rest == 0) { tot += (per - k) * 32; cnt = intensity; break; }
              int skip = __ffsll((long long)rest) - 1;     // full chunks: every pixel matches
              if (skip > 0) { tot += skip * 32; cnt = intensity; k += skip; }
              unsigned M = sh.M[base + k], I = sh.I[base + k];
              int painted = 32;
              for (int i = 0; i < 32; i++) {               // fill.c:27-39, pixel by pixel
                if ((M >> i) & 1u) cnt = intensity; else cnt--;
                if (cnt == 0 || !((I >> i) & 1u)) { painted = i; break; }
              }
              tot += painted;
              if (painted < 32) stop = true;
              k++;
            }
            if (stop) still &= ~(1u << d);
            if (d == 0) c0 = cnt; else if (d == 1) c1 = cnt; else if (d == 2) c2 = cnt; else c3 = cnt;
            if (li == 0) t0 = tot; else if (li == 1) t1 = tot; else if (li == 2) t2 = tot; else t3 = tot;
          }
        }
#pragma unroll
        for (int u = 0; u < FF_U; u++) {
          int ch = warp * FF_U + u, li = ch >> shift;
          if (li < na) {
            int d = (packed >> (2 * li)) & 3;
            int idx = ((ch & (per - 1)) << 5) + lane;
            if (idx < SEL4(li, t0, t1, t2, t3)) {
              int s = SEL4(d, d0, d1, d2, d3) + idx + 1;
              int dx = (d & 1) ? 0 : d - 1, dy = (d & 1) ? d - 2 : 0;
              ff_paint(c, fx + s * dx, fy + s * dy);
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int li = 0; li < 4; li++) {
          if (li >= na) break;
          int d = (packed >> (2 * li)) & 3, tot = SEL4(li, t0, t1, t2, t3);
          if (d == 0) d0 += tot; else if (d == 1) d1 += tot; else if (d == 2) d2 += tot; else d3 += tot;
        }
        active = still;
      }
      t_cx = fx; t_cy = fy; t_L = d0; t_T = d1; t_R = d2; t_B = d3; t_cur = 0;
      sp++;
      pending = false;
    }
    // ---- flood_fill_around_line x 4 (fill.c:45-79, :97-104): next still-matching neighbour ----
    unsigned total = 2u * ((unsigned)t_L + t_T + t_R + t_B);
    while (t_cur < total) {
      BFS_ADD(2, 1);
      unsigned anym = 0;
#pragma unroll
      for (int u = 0; u < FF_U; u++) {
        int ch = warp * FF_U + u;
        unsigned idx = t_cur + ch * 32 + lane;
        bool m = false;
        if (idx < total) {
          int qx, qy;
          ff_cand(t_cx, t_cy, t_L, t_T, t_R, idx, qx, qy);
          if (cx_in(c, qx, qy)) { int g = ff_gray_at(c, qx, qy); m = g >= lo && g <= hi; }
        }
        unsigned M = __ballot_sync(0xffffffffu, m);
        if (lane == 0) sh.M[ch] = M;
        anym |= M;
      }
      if (__syncthreads_or(anym != 0)) {   // barrier; the common round has no match at all
        unsigned Ma = sh.M[lane], Mb = sh.M[32 + lane];
        unsigned b_lo = __ballot_sync(0xffffffffu, Ma != 0), b_hi = __ballot_sync(0xffffffffu, Mb != 0);
        int hit = b_lo ? __ffs(b_lo) - 1 : 32 + __ffs(b_hi) - 1;
        unsigned Mh = sh.M[hit];
        __syncthreads();   // sh.M is rewritten by the next round
        unsigned idx = t_cur + hit * 32 + (__ffs(Mh) - 1);
        ff_cand(t_cx, t_cy, t_L, t_T, t_R, idx, fx, fy);
        t_cur = idx + 1;
        pending = true;
        break;
      }
      t_cur += BF_ROUND;
    }
    if (pending) continue;
    sp--;   // frame exhausted: return to the caller's frame
    if (sp == 0) break;
    const unsigned long long *s = stack + (size_t)(sp - 1) * 4;
    unsigned long long a = s[0], b = s[1], e = s[2], f = s[3];
    t_cx = (int)(unsigned)a; t_cy = (int)(unsigned)(a >> 32);
    t_L = (int)(unsigned)b; t_T = (int)(unsigned)(b >> 32);
    t_R = (int)(unsigned)e; t_B = (int)(unsigned)(e >> 32);
    t_cur = (unsigned)f;
  }
  return c.sx0;
}

__global__ void __launch_bounds__(BF_THREADS) k_bf_scan(DPage *pages, const DBfPos *pos, int npos, int abs_threshold,
                          unsigned long long intensity, int mask_lo, int mask_hi, int flag_off, int strip_rows) {
  __shared__ BfShared sh;
  extern __shared__ uint4 bf_strip[];
  DPage &pg = pages[blockIdx.x];
  const DImg im = pg.img;
  BfCtx c;
  c.data = im.data; c.sbuf = (uint8_t *)bf_strip; c.gen = im.fmt == DF_GRAY8 ? (const DImg *)0 : &pg.img;
  c.pitch = im.pitch; c.w = im.w; c.h = im.h; c.sx0 = -64;
  c.son = im.fmt == DF_GRAY8 && im.h <= strip_rows && (im.pitch & 15) == 0 && ((uintptr_t)im.data & 15) == 0;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *cand = (uint8_t *)(pg.u32 + flag_off);
#ifdef BF_STATS
  long long T0 = clock64();
#endif
  // phase 1: darkness of every position on the untouched image (blit.c:131-146)
  for (int k = threadIdx.x; k < npos; k += blockDim.x) {
    DBfPos q = pos[k];
    int x0 = max(q.r.x0, 0), x1 = min(q.r.x1, im.w - 1), y0 = max(q.r.y0, 0), y1 = min(q.r.y1, im.h - 1);
    unsigned long long cnt = (unsigned long long)(abs(x0 - x1) + 1) * (unsigned long long)(abs(y0 - y1) + 1);
    unsigned long long s = 0;
    if (x0 <= x1 && y0 <= y1) {
      const unsigned *b = pg.u32 + q.sum_off;
      if (q.axis == 0) for (int x = x0; x <= x1; x++) s += b[x];
      else for (int y = y0; y <= y1; y++) s += b[y];
    }
    int darkness = (int)(uint8_t)(0xFF - (s / cnt));
    cand[k] = darkness >= abs_threshold ? 1 : 0;
  }
  __syncthreads();
  // ordered list of the candidates: every warp compacts its own segment of positions
  unsigned *clist = pg.u32 + flag_off + (npos + 3) / 4 + 1;
  int nc;
  {
    int seg = (npos + BF_WARPS - 1) / BF_WARPS, b0 = warp * seg, b1 = min(b0 + seg, npos);
    unsigned cntw = 0;
    for (int base = b0; base < b1; base += 32) {
      int k = base + lane;
      cntw += __popc(__ballot_sync(0xffffffffu, k < b1 && cand[k]));
    }
    if (lane == 0) sh.wcnt[warp] = cntw;
    __syncthreads();
    unsigned off = 0, total = 0;
    for (int w = 0; w < BF_WARPS; w++) { if (w < warp) off += sh.wcnt[w]; total += sh.wcnt[w]; }
    for (int base = b0; base < b1; base += 32) {
      int k = base + lane;
      bool cf = k < b1 && cand[k];
      unsigned m = __ballot_sync(0xffffffffu, cf);
      if (cf) clist[off + __popc(m & ((1u << lane) - 1u))] = (unsigned)k;
      off += __popc(m);
    }
    nc = (int)total;
    __syncthreads();
  }
  // phase 2: candidates in scan order.  Before the first fill a candidate is a true hit;
  // afterwards a candidate has to be measured again, and as long as no further fill
  // happens those measurements are independent: one warp each, BF_WARPS per round
  bool dirty = false;
  unsigned fills = 0;
  for (int ci = 0; ci < nc;) {
    if (dirty) {
      BFS_ADD(3, 1);
      int mine = ci + warp;
      bool dark = false;
      if (mine < nc) {
        DBfPos q = pos[clist[mine]];
        int x0 = max(q.r.x0, 0), x1 = min(q.r.x1, im.w - 1), y0 = max(q.r.y0, 0), y1 = min(q.r.y1, im.h - 1);
        unsigned long long cnt = (unsigned long long)(abs(x0 - x1) + 1) * (unsigned long long)(abs(y0 - y1) + 1);
        unsigned long long s = 0;
        if (x0 <= x1 && y0 <= y1) {
          int w = x1 - x0 + 1, n = w * (y1 - y0 + 1);
          for (int i = lane; i < n; i += 32)
            s += c.gen ? (unsigned)px_darkinv(px_load(*c.gen, x0 + i % w, y0 + i / w))
                       : (unsigned)ff_gray_at(c, x0 + i % w, y0 + i / w);
        }
        s = warp_sum_u64(s);
        int darkness = (int)(uint8_t)(0xFF - (s / cnt));
        dark = darkness >= abs_threshold;
      }
      if (lane == 0) sh.M[warp] = dark ? 1u : 0u;
      __syncthreads();
      int first = -1;
      for (int w = 0; w < BF_WARPS; w++) if (sh.M[w]) { first = w; break; }
      __syncthreads();
      if (first < 0) { ci += BF_WARPS; continue; }
      ci += first;
    }
    DBfPos q = pos[clist[ci]];
    ci++;
    int x0 = max(q.r.x0, 0), x1 = min(q.r.x1, im.w - 1), y0 = max(q.r.y0, 0), y1 = min(q.r.y1, im.h - 1);
    dirty = true;
    fills++;
    // flood_fill from every pixel of the area in raster order (filters.c:86-89);
    // pixels outside the image never match
    int w = x1 - x0 + 1, n = (x0 <= x1 && y0 <= y1) ? w * (y1 - y0 + 1) : 0;
    for (int rb = 0; rb < n;) {
      unsigned anym = 0;
#pragma unroll
      for (int u = 0; u < FF_U; u++) {
        int ch = warp * FF_U + u;
        int i = rb + ch * 32 + lane;
        bool m = false;
        if (i < n) { int g = ff_gray_at(c, x0 + i % w, y0 + i / w); m = g >= mask_lo && g <= mask_hi; }
        unsigned M = __ballot_sync(0xffffffffu, m);
        if (lane == 0) sh.M[ch] = M;
        anym |= M;
      }
      int hit = -1;
      unsigned Mh = 0;
      if (__syncthreads_or(anym != 0)) {
        for (int ch = 0; ch < BF_CH; ch++) { unsigned M = sh.M[ch]; if (M) { hit = ch; Mh = M; break; } }
        __syncthreads();
      }
      if (hit < 0) { rb += BF_ROUND; continue; }
      int i = rb + hit * 32 + (__ffs(Mh) - 1);
      rb = i + 1;
      c.sx0 = ff_flood(c, (unsigned long long *)pg.stack, pg.stack_cap, (unsigned *)&pg.error, sh,
                       x0 + i % w, y0 + i / w, mask_lo, mask_hi, intensity);
    }
  }
  if (c.sx0 >= 0) { __syncthreads(); strip_copy(c, true); }
  if (threadIdx.x == 0) pg.bf_fills = fills;
#ifdef BF_STATS
  BFS_ADD(7, (unsigned long long)(clock64() - T0));
  BFS_ADD(8, (unsigned long long)nc);
  BFS_ADD(9, fills);
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("BFSTATS cross_steps %llu frames %llu cand_rounds %llu remeasure_rounds %llu total_cycles %llu ncand %llu fills %llu strip_moves %llu\n",
           g_bfs[0], g_bfs[1], g_bfs[2], g_bfs[3], g_bfs[7], g_bfs[8], g_bfs[9], g_bfs[14]);
#endif
}

/* =========================================================================
 * noisefilter (filters.c:243-348)
 *
 * NOT a component-size filter: pixels are visited in raster order, and a
 * dark pixel whose ring-neighbourhood count is <= intensity clears itself and
 * the rings up to the first empty one, which changes what later pixels see.
 * Exact parallel form:
 *   (1) classify: a pixel of an 8-connected set of >intensity "ring-dark"
 *       pixels (min channel < white) that lies outside the left/top band where
 *       the reference's unsigned ring walk is truncated can never be cleared
 *       -> PERMANENT.  Found with a bounded union walk inside a shared-memory
 *       tile (component size only matters up to intensity+1).
 *   (2) everything else that is trigger-dark (max channel < white) is MUTABLE
 *       and decided in raster order; two mutable pixels farther apart than
 *       R = 2*intensity cannot see each other's effects, so every round
 *       decides, in parallel, each undecided pixel that has no undecided
 *       raster-earlier pixel within R.
 * ====================================================================== */

#define NF_LIVE 1u
#define NF_MUT 2u
#define NF_UNDEC 4u
#define NF_TRIG 8u
#define NF_TW 64
#define NF_TH 16
#define NF_MAXI 15

// number of ring-dark, out-of-band pixels in the 3x3 block centred on (tx, ty)
__device__ __forceinline__ int nf_count9(const uint8_t *tile, int tw, int tx, int ty) {
  const uint8_t *t0 = tile + (ty - 1) * tw + (tx - 1);
  int c = 0;
#pragma unroll
  for (int dy = 0; dy < 3; dy++)
#pragma unroll
    for (int dx = 0; dx < 3; dx++) c += ((t0[dy * tw + dx] & 3) == 1);
  return c;
}
// Cheap sufficient test for "component has >= need pixels": the 3x3 block around
// the pixel, or around one of its dark neighbours, already holds that many (all
// of them touch that centre, which touches the pixel).  Needs a 2-pixel halo.
__device__ __forceinline__ bool nf_obviously_big(const uint8_t *tile, int tw, int tx, int ty, int need) {
  if (need > 9) return false;
  if (nf_count9(tile, tw, tx, ty) >= need) return true;
#pragma unroll 1
  for (int k = 0; k < 9; k++) {
    if (k == 4) continue;
    int nx = tx + k % 3 - 1, ny = ty + k / 3 - 1;
    if ((tile[ny * tw + nx] & 3) == 1 && nf_count9(tile, tw, nx, ny) >= need) return true;
  }
  return false;
}

// Bounded walk over 8-connected ring-dark, out-of-band pixels ((code & 3) == 1) of
// a shared-memory tile: true when fewer than `need` pixels are reachable from
// (tx, ty).  Kept out of line: inlined and unrolled it dwarfs the kernels.
__device__ __noinline__ bool nf_small_component(const uint8_t *tile, int tw, int th, int tx, int ty, int need) {
  short vx[NF_MAXI + 1], vy[NF_MAXI + 1];
  int n = 1, head = 0;
  vx[0] = (short)tx; vy[0] = (short)ty;
  while (head < n && n < need) {
    int cx = vx[head], cy = vy[head]; head++;
#pragma unroll 1
    for (int dy = -1; dy <= 1 && n < need; dy++)
#pragma unroll 1
      for (int dx = -1; dx <= 1 && n < need; dx++) {
        if (!dx && !dy) continue;
        int nx = cx + dx, ny = cy + dy;
        if (nx < 0 || ny < 0 || nx >= tw || ny >= th) continue;
        if ((tile[ny * tw + nx] & 3) != 1) continue;
        bool seen = false;
#pragma unroll 1
        for (int k = 0; k < n; k++) seen |= (vx[k] == nx && vy[k] == ny);
        if (!seen) { vx[n] = (short)nx; vy[n] = (short)ny; n++; }
      }
  }
  return n < need;
}

__global__ void k_nf_classify(DPage *pages, int intensity, int white, int all_mutable) {
  extern __shared__ uint8_t tile[];
  DPage &pg = pages[blockIdx.z];
  const DImg &im = pg.img;
  int halo = all_mutable ? 0 : intensity + 1;
  int tw = NF_TW + 2 * halo, th = NF_TH + 2 * halo;
  int bx = blockIdx.x * NF_TW, by = blockIdx.y * NF_TH;
  if (bx >= im.w || by >= im.h) return;
  int band = 2 * intensity;
  for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
    int x = bx - halo + i % tw, y = by - halo + i / tw;
    uint8_t v = 0;
    if (in_img(im, x, y)) {
      Px p = px_load(im, x, y);
      if (px_light(p) < white) v = 1 | ((x < band || y < band) ? 2 : 0) | (px_darkinv(p) < white ? 4 : 0);
    }
    tile[i] = v;
  }
  __syncthreads();
  int need = intensity + 1;
  for (int i = threadIdx.x; i < NF_TW * NF_TH; i += blockDim.x) {
    int lx = i % NF_TW, ly = i / NF_TW;
    int x = bx + lx, y = by + ly;
    if (x >= im.w || y >= im.h) continue;
    uint8_t v = tile[(ly + halo) * tw + lx + halo];
    uint8_t c = 0;
    if (v & 1) {
      bool mut = all_mutable || (v & 2);
      if (!mut && halo >= 2 && nf_obviously_big(tile, tw, lx + halo, ly + halo, need)) goto classified;
      if (!mut) mut = nf_small_component(tile, tw, th, lx + halo, ly + halo, need);
    classified:
      c = NF_LIVE | ((v & 4) ? NF_TRIG : 0);
      if (mut) {
        c |= NF_MUT;
        if (v & 4) {
          c |= NF_UNDEC;
          unsigned idx = atomicAdd(&pg.list_n, 1u);
          if (idx < (unsigned)pg.list_cap) pg.list[idx] = ((unsigned)y << 16) | (unsigned)x;
          else atomicOr(&pg.error, DERR_LIST_OVERFLOW);
        }
      }
    }
    pg.cls[(size_t)y * im.w + x] = c;
  }
}

// GRAY8 specialisation (intensity <= 7): 4 pixels per 32-bit load, tile x-halo
// fixed at 8 so that every tile word is 4-byte aligned, 4 classes per store.
#define NF_G8_HX 8
#define NF_G8_TWB (NF_TW + 2 * NF_G8_HX)
#define NF_G8_TH 32   /* interior rows per block: two 16-row halves per thread */
__global__ void k_nf_classify_g8(DPage *pages, int intensity, int white) {
  extern __shared__ uint8_t tile[];
  DPage &pg = pages[blockIdx.z];
  const DImg &im = pg.img;
  int halo = intensity + 1;
  int th = NF_G8_TH + 2 * halo;
  int bx = blockIdx.x * NF_TW, by = blockIdx.y * NF_G8_TH;
  if (bx >= im.w || by >= im.h) return;
  int band = 2 * intensity;
  bool aligned = ((im.pitch & 3) == 0) && (((uintptr_t)im.data & 3) == 0);
  const int WPR = NF_G8_TWB / 4;
  unsigned white4 = (unsigned)min(max(white, 0), 255) * 0x01010101u;
  if (white > 255) white4 = 0xFFFFFFFFu;
  for (int i = threadIdx.x; i < WPR * th; i += blockDim.x) {
    int r = i / WPR, c = i % WPR;
    int y = by - halo + r, x = bx - NF_G8_HX + 4 * c;
    unsigned codes = 0;
    if (y >= 0 && y < im.h && x + 3 >= 0 && x < im.w) {
      const uint8_t *rp = im.data + (size_t)y * im.pitch;
      unsigned v, inimg = 0xFFFFFFFFu;
      if (aligned && x >= 0 && x + 3 < im.w) v = *(const unsigned *)(rp + x);
      else {
        v = 0xFFFFFFFFu; inimg = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) if (x + k >= 0 && x + k < im.w) { v = (v & ~(0xFFu << (8 * k))) | ((unsigned)rp[x + k] << (8 * k)); inimg |= 0xFFu << (8 * k); }
      }
      // 0xFF in every byte lane whose pixel is ring-dark (value < white)
      unsigned dark = __vcmpltu4(v, white4) & inimg;
      if (dark) {
        unsigned inband;
        if (y < band || x + 3 < band) inband = 0xFFFFFFFFu;
        else if (x >= band) inband = 0u;
        else { inband = 0; for (int k = 0; k < 4; k++) if (x + k < band) inband |= 0xFFu << (8 * k); }
        codes = (dark & 0x05050505u) | (dark & inband & 0x02020202u);
      }
    }
    ((unsigned *)tile)[i] = codes;
  }
  __syncthreads();
  int need = intensity + 1;
  int t = threadIdx.x;
  for (int half = 0; half < NF_G8_TH / 16; half++) {
  int ly = t / (NF_TW / 4) + 16 * half, lx0 = 4 * (t % (NF_TW / 4));
  int y = by + ly;
  if (y >= im.h) return;
  unsigned out = 0;
  const int tw = NF_G8_TWB;
  // four white pixels (the common case): one shared-memory word, nothing to classify
  bool any = *(const unsigned *)(tile + (ly + halo) * tw + lx0 + NF_G8_HX) != 0u;
#pragma unroll 1
  for (int q = 0; any && q < 4; q++) {
    int lx = lx0 + q, x = bx + lx;
    if (x >= im.w) break;
    int tx = lx + NF_G8_HX, ty = ly + halo;
    uint8_t v = tile[ty * tw + tx];
    uint8_t c = 0;
    if (v & 1) {
      bool mut = (v & 2) != 0;
      if (!mut && !nf_obviously_big(tile, tw, tx, ty, need)) mut = nf_small_component(tile, tw, th, tx, ty, need);
      c = NF_LIVE | NF_TRIG;
      if (mut) {
        c |= NF_MUT | NF_UNDEC;
        unsigned idx = atomicAdd(&pg.list_n, 1u);
        if (idx < (unsigned)pg.list_cap) pg.list[idx] = ((unsigned)y << 16) | (unsigned)x;
        else atomicOr(&pg.error, DERR_LIST_OVERFLOW);
      }
    }
    out |= (unsigned)c << (8 * q);
  }
  size_t o = (size_t)y * im.w + bx + lx0;
  if ((im.w & 3) == 0 && bx + lx0 + 3 < im.w && ((uintptr_t)pg.cls & 3) == 0) *(unsigned *)(pg.cls + o) = out;
  else for (int q = 0; q < 4 && bx + lx0 + q < im.w; q++) pg.cls[o + q] = (uint8_t)(out >> (8 * q));
  }
}

/* GRAY8 bit-plane specialisation (intensity <= 7, 16-byte aligned rows).
 *
 * The tile (64x32 interior, 32 columns / intensity+1 rows of halo) is held as
 * one bit per pixel, four 32-bit words per row:
 *   dark  = value < white and inside the image        (ring-dark == trigger-dark for gray)
 *   nb    = dark outside the left/top band            (pixels that may form permanent components)
 * "Has >= need nb-pixels in its 3x3 block" is then a bit-sliced population count
 * over 9 shifted planes (core), and every nb pixel that is a core pixel or
 * touches one belongs to a component of >= need pixels (all members of a core
 * pixel's block touch it) — settled for 32 pixels per instruction.  Only the
 * few dark pixels that this test leaves open run the exact bounded walk. */
#define NFB_NW 12                     /* words per plane row: one halo word each side */
#define NFB_TW (32 * (NFB_NW - 2))
#define NFB_TH 64
#define NFB_ROWS (NFB_TH + 2 * 8)

__device__ __forceinline__ bool nfb_bit(const unsigned (*nb)[NFB_NW], int x, int y) { return (nb[y][x >> 5] >> (x & 31)) & 1u; }

__device__ __noinline__ bool nfb_small_component(const unsigned (*nb)[NFB_NW], int rows, int tx, int ty, int need) {
  short vx[8], vy[8];
  int n = 1, head = 0;
  vx[0] = (short)tx; vy[0] = (short)ty;
  while (head < n && n < need) {
    int cx = vx[head], cy = vy[head]; head++;
#pragma unroll 1
    for (int k = 0; k < 9 && n < need; k++) {
      if (k == 4) continue;
      int nx = cx + k % 3 - 1, ny = cy + k / 3 - 1;
      if (nx < 0 || ny < 0 || nx >= 32 * NFB_NW || ny >= rows) continue;
      if (!nfb_bit(nb, nx, ny)) continue;
      bool seen = false;
#pragma unroll 1
      for (int q = 0; q < n; q++) seen |= (vx[q] == nx && vy[q] == ny);
      if (!seen) { vx[n] = (short)nx; vy[n] = (short)ny; n++; }
    }
  }
  return n < need;
}

__global__ void __launch_bounds__(256, 8) k_nf_classify_bits(DPage *pages, int intensity, int white) {
  // six bit planes, each with one guard word before and after: the word left of word 0 of a row is the last word
  // of the row above (and the other way round), whose bit next to the row boundary is always clear — a halo word
  // only carries the eight pixels next to the interior, a word that leaves the image none beyond it — so the
  // neighbour exchanges below need no edge cases
  __shared__ unsigned s_planes[6][NFB_ROWS * NFB_NW + 2];
  unsigned (*const s_dark)[NFB_NW] = (unsigned (*)[NFB_NW])(&s_planes[0][1]);
  unsigned (*const s_nb)[NFB_NW] = (unsigned (*)[NFB_NW])(&s_planes[1][1]);
  unsigned *const f_nb = &s_planes[1][1], *const f_h0 = &s_planes[2][1], *const f_h1 = &s_planes[3][1];
  unsigned *const f_core = &s_planes[4][1], *const f_big = &s_planes[5][1];
  unsigned (*const s_big)[NFB_NW] = (unsigned (*)[NFB_NW])f_big;
  DPage &pg = pages[blockIdx.z];
  const DImg &im = pg.img;
  uint8_t *const cls = pg.cls;
  int halo = intensity + 1, rows = NFB_TH + 2 * halo, need = intensity + 1, band = 2 * intensity;
  int bx = blockIdx.x * NFB_TW, by = blockIdx.y * NFB_TH;
  if (bx >= im.w || by >= im.h) return;
  if (threadIdx.x == 0) { f_nb[-1] = 0u; f_nb[rows * NFB_NW] = 0u; f_core[-1] = 0u; f_core[rows * NFB_NW] = 0u; }   // the words next to the planes in use
  int x_org = bx - 32, y_org = by - halo;
  unsigned white4 = white > 255 ? 0xFFFFFFFFu : (unsigned)max(white, 0) * 0x01010101u;
  // phase 1: bit planes
  for (int item = threadIdx.x; item < rows * NFB_NW; item += blockDim.x) {
    int r = item / NFB_NW, k = item % NFB_NW;
    int y = y_org + r, x0 = x_org + 32 * k;
    unsigned dark = 0;
    if (y >= 0 && y < im.h && x0 + 31 >= 0 && x0 < im.w) {
      const uint8_t *rp = im.data + (size_t)y * im.pitch;
      if (x0 >= 0 && x0 + 31 < im.w && (k == 0 || k == NFB_NW - 1)) {
        // halo word: only the 8 pixels next to the interior are ever looked at (halo <= 8)
        int off = k == 0 ? 24 : 0;
        uint2 a = *(const uint2 *)(rp + x0 + off);
        unsigned m0 = __vcmpltu4(a.x, white4), m1 = __vcmpltu4(a.y, white4);
        unsigned d8 = ((((m0 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) | (((((m1 & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 4);
        dark = d8 << off;
      } else if (x0 >= 0 && x0 + 31 < im.w) {
        const uint4 *q = (const uint4 *)(rp + x0);
        uint4 a = q[0], b = q[1];
        unsigned wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
          unsigned m = __vcmpltu4(wv[j], white4);
          dark |= ((((m & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << (4 * j);
        }
      } else {
        for (int i = 0; i < 32; i++) { int x = x0 + i; if (x >= 0 && x < im.w && (int)rp[x] < white) dark |= 1u << i; }
      }
    }
    unsigned inb;
    if (y < band || x0 + 31 < band) inb = 0xFFFFFFFFu;
    else if (x0 >= band) inb = 0u;
    else inb = (band - x0 >= 32) ? 0xFFFFFFFFu : ((1u << (band - x0)) - 1u);
    s_dark[r][k] = dark;
    s_nb[r][k] = dark & ~inb;
  }
  __syncthreads();
  // phase 2a: per row, how many of (left, self, right) are set: h1:h0
  const int nitems = rows * NFB_NW;
  for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
    const unsigned C = f_nb[i], P = f_nb[i - 1], N = f_nb[i + 1];
    const unsigned L = __funnelshift_l(P, C, 1), R = __funnelshift_r(C, N, 1);   // (C << 1) | (P >> 31), (C >> 1) | (N << 31)
    f_h0[i] = L ^ C ^ R;
    f_h1[i] = (L & C) | (C & R) | (L & R);
  }
  __syncthreads();
  // phase 2b: 3x3 population (bit-sliced) >= need, centred on an nb pixel
  {
    unsigned kb[4];
#pragma unroll
    for (int b = 0; b < 4; b++) kb[b] = ((need >> b) & 1) ? 0xFFFFFFFFu : 0u;
    const unsigned live = need > 9 ? 0u : 0xFFFFFFFFu;
    for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
      unsigned core = 0;
      if (i >= NFB_NW && i < nitems - NFB_NW) {        // rows 1 .. rows - 2
        const unsigned a0 = f_h0[i - NFB_NW], a1 = f_h1[i - NFB_NW], b0 = f_h0[i], b1 = f_h1[i], c0 = f_h0[i + NFB_NW], c1 = f_h1[i + NFB_NW];
        const unsigned s0 = a0 ^ b0 ^ c0, k0 = (a0 & b0) | (b0 & c0) | (a0 & c0);      // weight 1, carry (weight 2)
        const unsigned t1 = a1 ^ b1 ^ c1, k1 = (a1 & b1) | (b1 & c1) | (a1 & c1);      // weight 2, carry (weight 4)
        const unsigned u1 = t1 ^ k0, c2 = t1 & k0;                                      // weight 2, carry (weight 4)
        const unsigned v2 = k1 ^ c2, v3 = k1 & c2;                                      // weight 4, weight 8
        const unsigned bits[4] = {s0, u1, v2, v3};
        unsigned gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 3; b >= 0; b--) {
          gt |= eq & bits[b] & ~kb[b];
          eq &= ~(bits[b] ^ kb[b]);
        }
        core = (gt | eq) & f_nb[i] & live;
      }
      f_core[i] = core;
    }
  }
  __syncthreads();
  // phase 2c: nb pixels that are a core pixel or touch one
  for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
    unsigned D = 0;
    if (i >= NFB_NW && i < nitems - NFB_NW) {
#pragma unroll
      for (int rr = -1; rr <= 1; rr++) {
        const int j = i + rr * NFB_NW;
        const unsigned C = f_core[j], P = f_core[j - 1], N = f_core[j + 1];
        D |= C | __funnelshift_l(P, C, 1) | __funnelshift_r(C, N, 1);
      }
    }
    f_big[i] = f_nb[i] & D;
  }
  __syncthreads();
  // phase 3: class bytes.  An item is one 32-pixel word of the bit planes (two 16-byte stores) when the
  // class map allows it, else eight pixels.
  const bool wide = (im.w & 15) == 0 && ((uintptr_t)cls & 15) == 0;
  const int ppi = wide ? 32 : 8;                       // pixels per item
  const int ipr = NFB_TW / ppi;                        // items per tile row
  for (int it = threadIdx.x; it < ipr * NFB_TH; it += blockDim.x) {
    const int ly = wide ? it / (NFB_TW / 32) : it / (NFB_TW / 8), lx0 = (it - ly * ipr) * ppi;   // divisions by constants
    const int y = by + ly, r = ly + halo;
    if (y >= im.h || bx + lx0 >= im.w) continue;
    const int k = 1 + (lx0 >> 5), sh = lx0 & 31;
    const unsigned pm = wide ? 0xFFFFFFFFu : 0xFFu;
    const unsigned darkw = (s_dark[r][k] >> sh) & pm, nbw = (s_nb[r][k] >> sh) & pm, bigw = (s_big[r][k] >> sh) & pm;
    // mutable = dark inside the left/top band, or dark outside it, not settled by the bit planes and
    // found small by the exact walk (rare); no per-pixel loop for the rest
    unsigned mutw = darkw & ~nbw;
    unsigned openw = nbw & ~bigw;
    while (openw) {
      const int i = __ffs(openw) - 1;
      openw &= openw - 1;
      if (nfb_small_component(s_nb, rows, 32 + lx0 + i, r, need)) mutw |= 1u << i;
    }
    if (mutw) {
      unsigned idx = atomicAdd(&pg.list_n, (unsigned)__popc(mutw));
      for (unsigned m = mutw; m; m &= m - 1, idx++) {
        if (idx < (unsigned)pg.list_cap) pg.list[idx] = ((unsigned)y << 16) | (unsigned)(bx + lx0 + __ffs(m) - 1);
        else atomicOr(&pg.error, DERR_LIST_OVERFLOW);
      }
    }
    // four bits -> four bytes (bit i -> bit 0 of byte i), then the class codes
    const size_t o = (size_t)y * im.w + bx + lx0;
    if (wide) {
      unsigned cw[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const unsigned d4 = (((darkw >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u, m4 = (((mutw >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u;
        cw[q] = d4 * (NF_LIVE | NF_TRIG) + m4 * (NF_MUT | NF_UNDEC);
      }
      uint4 *o4 = (uint4 *)(cls + o);
      if (bx + lx0 + 31 < im.w) { o4[0] = make_uint4(cw[0], cw[1], cw[2], cw[3]); o4[1] = make_uint4(cw[4], cw[5], cw[6], cw[7]); }
      else {   // the image ends inside this word (its width is a multiple of 16)
        if (bx + lx0 + 15 < im.w) o4[0] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      }
    } else {
      const unsigned dl = ((darkw & 0xFu) * 0x00204081u) & 0x01010101u, dh = ((darkw >> 4) * 0x00204081u) & 0x01010101u;
      const unsigned ml = ((mutw & 0xFu) * 0x00204081u) & 0x01010101u, mh = ((mutw >> 4) * 0x00204081u) & 0x01010101u;
      const unsigned lo = dl * (NF_LIVE | NF_TRIG) + ml * (NF_MUT | NF_UNDEC), hi = dh * (NF_LIVE | NF_TRIG) + mh * (NF_MUT | NF_UNDEC);
      if ((im.w & 7) == 0 && ((uintptr_t)cls & 7) == 0) *(uint2 *)(cls + o) = make_uint2(lo, hi);
      else for (int i = 0; i < 8 && bx + lx0 + i < im.w; i++) cls[o + i] = (uint8_t)((i < 4 ? lo >> (8 * i) : hi >> (8 * (i - 4))) & 0xFFu);
    }
  }
}

__device__ __forceinline__ bool nf_live(const uint8_t *cls, int w, int h, int x, int y) {
  return (unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h && (cls[(size_t)y * w + x] & NF_LIVE);
}
// noisefilter_compare_and_clear (filters.c:243-254) on the class map
__device__ __forceinline__ unsigned nf_cc(DPage &pg, uint8_t *cls, int x, int y, bool clear) {
  const DImg &im = pg.img;
  if (!nf_live(cls, im.w, im.h, x, y)) return 0;
  if (clear) {
    uint8_t c = cls[(size_t)y * im.w + x];
    if (!(c & NF_MUT)) atomicOr(&pg.error, DERR_UNSUPPORTED);  // permanence proof violated
    cls[(size_t)y * im.w + x] = 0;
    px_store(im, x, y, 255, 255, 255);
  }
  return 1;
}
// noisefilter_count_pixel_neighbors_level (filters.c:256-285) incl. the
// unsigned-compare truncation near the left/top edge
__device__ unsigned nf_ring(DPage &pg, uint8_t *cls, int px, int py, int level, bool clear) {
  unsigned count = 0;
  if (px >= level)
    for (int xx = px - level; xx <= px + level; xx++) {
      count += nf_cc(pg, cls, xx, py - level, clear);
      count += nf_cc(pg, cls, xx, py + level, clear);
    }
  if (py >= level - 1)
    for (int yy = py - (level - 1); yy <= py + (level - 1); yy++) {
      count += nf_cc(pg, cls, px - level, yy, clear);
      count += nf_cc(pg, cls, px + level, yy, clear);
    }
  return count;
}

__global__ void __launch_bounds__(1024) k_nf_resolve(DPage *pages, int intensity) {
  DPage &pg = pages[blockIdx.x];
  const DImg &im = pg.img;
  uint8_t *cls = pg.cls;
  // a truncated list would leave undecided pixels that nobody owns: refuse (the
  // caller sees DERR_LIST_OVERFLOW) instead of spinning
  if (pg.list_n > (unsigned)pg.list_cap) return;
  int n = (int)pg.list_n;
  int R = 2 * intensity;
  __shared__ unsigned s_clusters;
  if (threadIdx.x == 0) s_clusters = 0;
  __syncthreads();
  // every round decides at least the raster-first undecided pixel
  for (int round = 0; round <= n; round++) {
    // phase A: readiness against the state at the start of the round
    int pending = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      unsigned v = pg.list[e];
      int x = v & 0xFFFF, y = (v >> 16) & 0x7FFF;
      if (!(cls[(size_t)y * im.w + x] & NF_UNDEC)) continue;
      pending = 1;
      bool ready = true;
      int ya = max(y - R, 0), xa = max(x - R, 0), xb = min(x + R, im.w - 1);
      for (int yy = ya; yy <= y && ready; yy++) {
        const uint8_t *row = cls + (size_t)yy * im.w;
        int xe = (yy == y) ? x - 1 : xb;
        if (xe < xa) continue;
        // the window row as aligned words (independent loads; the class map has 64 bytes
        // of slack behind its last row), bytes outside [xa, xe] masked off
        uintptr_t p0 = (uintptr_t)(row + xa), p1 = (uintptr_t)(row + xe);
        const unsigned *wp = (const unsigned *)(p0 & ~(uintptr_t)3);
        int nw = (int)(((p1 & ~(uintptr_t)3) - (p0 & ~(uintptr_t)3)) >> 2) + 1;
        unsigned acc = 0;
        for (int k = 0; k < nw; k++) {
          unsigned v = wp[k];
          if (k == 0) v &= 0xFFFFFFFFu << (8 * (unsigned)(p0 & 3));
          if (k == nw - 1) v &= 0xFFFFFFFFu >> (8 * (3 - (unsigned)(p1 & 3)));
          acc |= v;
        }
        if (acc & (NF_UNDEC * 0x01010101u)) ready = false;
      }
      if (ready) pg.list[e] = v | 0x80000000u;
    }
    if (!__syncthreads_or(pending)) break;
    // phase B: ready pixels are pairwise farther apart than R -> independent
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      unsigned v = pg.list[e];
      if (!(v & 0x80000000u)) continue;
      pg.list[e] = v & 0x7FFFFFFFu;
      int x = v & 0xFFFF, y = (v >> 16) & 0x7FFF;
      size_t o = (size_t)y * im.w + x;
      uint8_t c = cls[o];
      if (c & NF_LIVE) {
        unsigned long long count = 1;   // filters.c:287-302
        unsigned l;
        int level = 1;
        do { l = nf_ring(pg, cls, x, y, level, false); count += l; level++; } while (l != 0 && level <= intensity);
        if (count <= (unsigned long long)intensity) {   // filters.c:304-317
          cls[o] = 0;
          px_store(im, x, y, 255, 255, 255);
          level = 1;
          do { l = nf_ring(pg, cls, x, y, level, true); level++; } while (l != 0);
          atomicAdd(&s_clusters, 1u);
          c = 0;
        }
      }
      if (c) cls[o] = c & ~NF_UNDEC;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) pg.nf_clusters = s_clusters;
}

/* =========================================================================
 * blurfilter (filters.c:149-232)
 *
 * Every count the CPU loop reads is taken strictly below / right of what it
 * has wiped so far, i.e. on the untouched image: count them all in parallel
 * (k_rect_count), replay the scalar state machine — INCLUDING the reference's
 * aliased prev/cur/next pointers into one zero-initialised buffer
 * (filters.c:160-167) — and wipe the flagged blocks.
 * layout in u32: [c0: n][h: nrows*(n+1)][state: 3*(n+2)][flags: nrows*n]
 * ====================================================================== */
// the scalar state machine of filters.c:170-230 on arrays `c0`, `h`, `flat`, `flag`
// (shared or global memory)
__device__ __forceinline__ void blur_state_machine(const unsigned *c0, const unsigned *h, unsigned *flat, unsigned *flag,
                                                   int n, int nrows, unsigned long long T, float intensity) {
  unsigned Tu = (unsigned)T;
  for (int i = 0; i < 3 * (n + 2); i++) flat[i] = 0;
  int prev = 0, cur = 1, next = 2;
  flat[cur + 0] = Tu; flat[cur + n] = Tu; flat[next + 0] = Tu; flat[next + n] = Tu;
  for (int b = 0; b < n; b++) flat[cur + 1 + b] = c0[b];
  for (int r = 0; r < nrows; r++) {
    flat[next + 0] = h[(size_t)r * (n + 1) + 0];
    for (int b = 0, block = 1; b < n; b++, block++) {
      flat[next + block + 1] = h[(size_t)r * (n + 1) + block];
      unsigned m1 = max(flat[prev + block - 1], max(flat[prev + block + 1], flat[cur + block]));
      unsigned m = max(flat[next + block - 1], max(flat[next + block + 1], m1));
      bool wipe = ((float)(unsigned long long)m / (float)T) <= intensity;
      flag[(size_t)r * n + b] = wipe ? 1u : 0u;
      if (wipe) flat[cur + block] = Tu;
    }
    int tmp = prev; prev = cur; cur = next; next = tmp;
  }
}

// one warp per page: the counts are staged in shared memory so the serial walk
// pays shared-memory latency per step, not an L2 round trip
__global__ void __launch_bounds__(32) k_blur_decide_sm(DPage *pages, int n, int nrows, unsigned long long T,
                                                      float intensity, int cnt_off, int state_off, int flag_off) {
  extern __shared__ unsigned bsm[];   // [c0: n][h: nrows*(n+1)][flat: 3*(n+2)][flag: nrows*n]
  DPage &pg = pages[blockIdx.x];
  int ncnt = n + nrows * (n + 1);
  unsigned *flat = bsm + ncnt, *flag = flat + 3 * (n + 2);
  const unsigned *src = pg.u32 + cnt_off;
  for (int i = threadIdx.x; i < ncnt; i += 32) bsm[i] = src[i];
  __syncwarp();
  if (threadIdx.x == 0) blur_state_machine(bsm, bsm + n, flat, flag, n, nrows, T, intensity);
  __syncwarp();
  unsigned *gflat = pg.u32 + state_off, *gflag = pg.u32 + flag_off;
  for (int i = threadIdx.x; i < 3 * (n + 2); i += 32) gflat[i] = flat[i];
  for (int i = threadIdx.x; i < nrows * n; i += 32) gflag[i] = flag[i];
}

__global__ void k_blur_decide(DPage *pages, int npages, int n, int nrows, unsigned long long T,
                              float intensity, int cnt_off, int state_off, int flag_off) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  const unsigned *c0 = pg.u32 + cnt_off;
  blur_state_machine(c0, c0 + n, pg.u32 + state_off, pg.u32 + flag_off, n, nrows, T, intensity);
}

__global__ void k_blur_wipe(DPage *pages, int n, int bw, int bh, int flag_off) {
  DPage &pg = pages[blockIdx.z];
  int b = blockIdx.x, r = blockIdx.y;
  if (!pg.u32[flag_off + (size_t)r * n + b]) return;
  const DImg &im = pg.img;
  int x0 = b * bw, y0 = r * bh;
  if (im.fmt == DF_GRAY8 && (bw & 3) == 0 && (im.pitch & 3) == 0 && ((uintptr_t)im.data & 3) == 0) {
    // a warp per row, 32-bit stores (x0 is a multiple of 4); what lies outside the image is dropped like set_pixel does
    const int xe = min(x0 + bw, im.w), ye = min(y0 + bh, im.h);
    const int nw = max(xe - x0, 0) >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int y = y0 + warp; y < ye; y += nwarps) {
      uint8_t *row = im.data + (size_t)y * im.pitch + x0;
      for (int i = lane; i < nw; i += 32) ((unsigned *)row)[i] = 0xFFFFFFFFu;
      for (int x = x0 + 4 * nw + lane; x < xe; x += 32) row[x - x0] = 255;
    }
    return;
  }
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
    int x = x0 + i % bw, y = y0 + i / bw;
    px_set(im, x, y, 255, 255, 255);
  }
}

/* =========================================================================
 * grayfilter (filters.c:370-402)
 *
 * Windows are visited in raster order and a wiped window raises the lightness
 * of every later window it overlaps, so decisions cascade.  A wiped window
 * had zero dark pixels, hence dark counts never change: keep per-cell
 * (gx x gy, g = gcd(size, step)) dark counts and lightness sums, and run the
 * cascade as a skewed wavefront (window (i,j) only depends on windows with a
 * smaller j + skew*i) inside one CTA per page.
 * layout in u32: [dark: nc][light: nc][wiped: nc]   nc = ncx*ncy
 * ====================================================================== */
struct GrayParams {
  int gx, gy, ncx, ncy;       // cell size / grid
  int wcx, wcy, scx, scy;     // window size / step in cells
  int nwx, nwy, skew;
  int size_w, size_h, step_h, step_v;
  int abs_threshold, oob_dark;
  int off;
};

// Per window, before the cascade: a window with a dark pixel is never wiped, and
// wiping a window whose cells are all pure white changes nothing; only the rest
// ("live" windows) can alter the sheet.  Marks them and the wavefronts that hold one.
__global__ void k_gray_windows(DPage *pages, GrayParams gp, int white_off, int wflag_off, int wave_off) {
  DPage &pg = pages[blockIdx.y];
  const DImg &im = pg.img;
  int nc = gp.ncx * gp.ncy;
  const unsigned *dark = pg.u32 + gp.off;
  const unsigned *white = pg.u32 + white_off;
  unsigned *wflag = pg.u32 + wflag_off;
  unsigned *wave = pg.u32 + wave_off;
  int nw = gp.nwx * gp.nwy;
  for (int wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nw; wi += gridDim.x * blockDim.x) {
    int i = wi / gp.nwx, j = wi % gp.nwx;
    int x0 = j * gp.step_h, y0 = i * gp.step_v;
    int x1 = x0 + gp.size_w - 1, y1 = y0 + gp.size_h - 1;
    int x0c = max(x0, 0), x1c = min(x1, im.w - 1), y0c = max(y0, 0), y1c = min(y1, im.h - 1);
    bool inside = x0c <= x1c && y0c <= y1c;
    unsigned long long d = 0;
    bool allwhite = true;
    int cx0 = j * gp.scx, cy0 = i * gp.scy;
    for (int cy = cy0; cy < cy0 + gp.wcy; cy++)
      for (int cx = cx0; cx < cx0 + gp.wcx; cx++) {
        d += dark[cy * gp.ncx + cx];
        allwhite = allwhite && white[cy * gp.ncx + cx];
      }
    long long area_in = inside ? (long long)(x1c - x0c + 1) * (y1c - y0c + 1) : 0;
    if (gp.oob_dark) d += (unsigned long long)((long long)gp.size_w * gp.size_h - area_in);
    // Y400A: a wipe also rewrites alpha, so an all-white window is not a no-op there
    if (im.fmt == DF_Y400A) allwhite = false;
    unsigned live = (d == 0 && !allwhite && inside) ? 1u : 0u;
    wflag[wi] = live;
    if (live) wave[j + gp.skew * i] = 1u;
  }
  (void)nc;
}

__global__ void k_gray_cascade(DPage *pages, GrayParams gp, int wflag_off, int wave_off) {
  DPage &pg = pages[blockIdx.x];
  const DImg &im = pg.img;
  int nc = gp.ncx * gp.ncy;
  unsigned *light = pg.u32 + gp.off + nc;
  unsigned *wiped = pg.u32 + gp.off + 2 * nc;
  const unsigned *wflag = pg.u32 + wflag_off;
  const unsigned *wave = pg.u32 + wave_off;
  int nwave = gp.nwx + gp.skew * (gp.nwy - 1);
  for (int t = 0; t < nwave; t++) {
    if (!wave[t]) continue;   // block-uniform: no live window on this wavefront
    for (int i = threadIdx.x; i < gp.nwy; i += blockDim.x) {
      int j = t - gp.skew * i;
      if (j < 0 || j >= gp.nwx || !wflag[i * gp.nwx + j]) continue;
      int x0 = j * gp.step_h, y0 = i * gp.step_v;
      int x1 = x0 + gp.size_w - 1, y1 = y0 + gp.size_h - 1;
      int x0c = max(x0, 0), x1c = min(x1, im.w - 1), y0c = max(y0, 0), y1c = min(y1, im.h - 1);
      unsigned long long l = 0;
      int cx0 = j * gp.scx, cy0 = i * gp.scy;
      for (int cy = cy0; cy < cy0 + gp.wcy; cy++)
        for (int cx = cx0; cx < cx0 + gp.wcx; cx++) l += light[cy * gp.ncx + cx];
      // inverse_lightness_rect (blit.c:111-126) on the clipped window (live windows intersect the image)
      unsigned long long cnt = (unsigned long long)(abs(x0c - x1c) + 1) * (unsigned long long)(abs(y0c - y1c) + 1);
      int lightness = (int)(uint8_t)(0xFF - (l / cnt));
      if (lightness < gp.abs_threshold) {
        for (int cy = cy0; cy < cy0 + gp.wcy; cy++)
          for (int cx = cx0; cx < cx0 + gp.wcx; cx++) {
            int ax0 = cx * gp.gx, ay0 = cy * gp.gy;
            int w = max(0, min(ax0 + gp.gx, im.w) - ax0), h = max(0, min(ay0 + gp.gy, im.h) - ay0);
            light[cy * gp.ncx + cx] = 255u * (unsigned)(w * h);
            wiped[cy * gp.ncx + cx] = 1;
          }
      }
    }
    __syncthreads();
  }
}

__global__ void k_zero_range(DPage *pages, int off_a, int na, int off_b, int nb) {
  unsigned *u = pages[blockIdx.y].u32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += gridDim.x * blockDim.x) {
    if (i < na) u[off_a + i] = 0u; else u[off_b + (i - na)] = 0u;
  }
}

// Snapshot which cells were already pure white before the cascade
// (light == 255*npix) so that the wipe pass can skip them.
__global__ void k_gray_prewhite(DPage *pages, GrayParams gp, int white_off) {
  DPage &pg = pages[blockIdx.y];
  const DImg &im = pg.img;
  int nc = gp.ncx * gp.ncy;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    int cx = c % gp.ncx, cy = c / gp.ncx;
    int ax0 = cx * gp.gx, ay0 = cy * gp.gy;
    int w = max(0, min(ax0 + gp.gx, im.w) - ax0), h = max(0, min(ay0 + gp.gy, im.h) - ay0);
    pg.u32[white_off + c] = (pg.u32[gp.off + nc + c] == 255u * (unsigned)(w * h)) ? 1u : 0u;
  }
}

// A thread per cell: wiped cells that were not pure white before are rare, so the pass is a scan of the
// cell flags; the lanes of a warp then paint each flagged cell of their 32 together.
__global__ void __launch_bounds__(256) k_gray_wipe(DPage *pages, GrayParams gp, int white_off) {
  DPage &pg = pages[blockIdx.y];
  const DImg &im = pg.img;
  const int nc = gp.ncx * gp.ncy;
  const unsigned *wiped = pg.u32 + gp.off + 2 * nc;
  const unsigned *white = pg.u32 + white_off;
  const bool skip_white = im.fmt != DF_Y400A;
  const int lane = threadIdx.x & 31;
  for (int c0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; c0 < nc; c0 += gridDim.x * blockDim.x) {   // warp-uniform trip count
    const int c = c0 + lane;
    const bool f = c < nc && wiped[c] && !(skip_white && white[c]);
    unsigned m = __ballot_sync(0xffffffffu, f);
    while (m) {
      const int cc = c0 + __ffs(m) - 1;
      m &= m - 1;
      const int cx = cc % gp.ncx, cy = cc / gp.ncx;
      const int x0 = cx * gp.gx, y0 = cy * gp.gy;
      const int w = min(x0 + gp.gx, im.w) - x0, h = min(y0 + gp.gy, im.h) - y0;
      if (w <= 0 || h <= 0) continue;
      for (int i = lane; i < w * h; i += 32) px_store(im, x0 + i % w, y0 + i / w, 255, 255, 255);
    }
  }
}

extern "C" {

void b200k_bf_scan(cudaStream_t st, DPage *pages, int npages, const DBfPos *pos_dev, int npos,
                   int abs_threshold, long long intensity, int mask_lo, int mask_hi, int flag_off, int maxh) {
  if (npages <= 0 || npos <= 0) return;
  // the shared-memory column strip: 32 bytes per image row, if the page height allows
  int strip_rows = (maxh > 0 && (size_t)maxh * 32 <= 200 * 1024) ? maxh : 0;
  size_t sm = (size_t)strip_rows * 32;
  // per device and cheap, so set on every launch (an engine per GPU may live in one process)
  if (sm > 40 * 1024) cudaFuncSetAttribute(k_bf_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_bf_scan<<<npages, BF_THREADS, sm, st>>>(pages, pos_dev, npos, abs_threshold, (unsigned long long)intensity,
                                  mask_lo, mask_hi, flag_off, strip_rows);
}

int b200k_noisefilter(cudaStream_t st, DPage *pages, int npages, int maxw, int maxh, int fmt,
                      unsigned long long intensity, int white, int flags) {
  if (npages <= 0 || intensity == 0) return 0;
  if (intensity > 4000) return -1;
  int I = (int)intensity;
  int all_mut = I > NF_MAXI;
  int halo = all_mut ? 0 : I + 1;
  size_t sm = (size_t)(NF_TW + 2 * halo) * (NF_TH + 2 * halo);
  dim3 g(cdiv(maxw, NF_TW), cdiv(maxh, NF_TH), npages);
  if (fmt == DF_GRAY8 && I <= 7 && (flags & 1)) {
    // rows 16-byte aligned (checked by the caller): one bit per pixel
    dim3 gb(cdiv(maxw, NFB_TW), cdiv(maxh, NFB_TH), npages);
    k_nf_classify_bits<<<gb, 256, 0, st>>>(pages, I, white);
  } else if (fmt == DF_GRAY8 && I <= 7) {
    size_t smg = (size_t)NF_G8_TWB * (NF_G8_TH + 2 * (I + 1));
    dim3 g8(cdiv(maxw, NF_TW), cdiv(maxh, NF_G8_TH), npages);
    k_nf_classify_g8<<<g8, 256, smg, st>>>(pages, I, white);
  } else
    k_nf_classify<<<g, 256, sm, st>>>(pages, I, white, all_mut);
  k_nf_resolve<<<npages, 1024, 0, st>>>(pages, I);   // one CTA per page: latency-bound rounds, so as wide as a CTA gets
  return 0;
}

void b200k_blur_decide(cudaStream_t st, DPage *pages, int npages, int n, int nrows,
                       unsigned long long T, float intensity, int cnt_off, int state_off, int flag_off) {
  if (npages <= 0) return;
  size_t sm = ((size_t)n + (size_t)nrows * (n + 1) + 3 * ((size_t)n + 2) + (size_t)nrows * n) * sizeof(unsigned);
  if (sm <= 40 * 1024)
    k_blur_decide_sm<<<npages, 32, sm, st>>>(pages, n, nrows, T, intensity, cnt_off, state_off, flag_off);
  else
    k_blur_decide<<<cdiv(npages, 32), 32, 0, st>>>(pages, npages, n, nrows, T, intensity, cnt_off, state_off, flag_off);
}
void b200k_blur_wipe(cudaStream_t st, DPage *pages, int npages, int n, int nrows, int bw, int bh, int flag_off) {
  if (npages <= 0 || n <= 0 || nrows <= 0) return;
  dim3 g(n, nrows, npages);
  k_blur_wipe<<<g, 256, 0, st>>>(pages, n, bw, bh, flag_off);
}

void b200k_gray_cascade(cudaStream_t st, DPage *pages, int npages, const int *gpi, int white_off) {
  if (npages <= 0) return;
  GrayParams gp;
  gp.gx = gpi[0]; gp.gy = gpi[1]; gp.ncx = gpi[2]; gp.ncy = gpi[3]; gp.wcx = gpi[4]; gp.wcy = gpi[5];
  gp.scx = gpi[6]; gp.scy = gpi[7]; gp.nwx = gpi[8]; gp.nwy = gpi[9]; gp.skew = gpi[10];
  gp.size_w = gpi[11]; gp.size_h = gpi[12]; gp.step_h = gpi[13]; gp.step_v = gpi[14];
  gp.abs_threshold = gpi[15]; gp.oob_dark = gpi[16]; gp.off = gpi[17];
  int nc = gp.ncx * gp.ncy;
  int nw = gp.nwx * gp.nwy, nwave = gp.nwx + gp.skew * (gp.nwy - 1);
  int wflag_off = white_off + nc, wave_off = wflag_off + nw;
  dim3 g1(min(cdiv(nc, 256), 128u), npages);
  k_gray_prewhite<<<g1, 256, 0, st>>>(pages, gp, white_off);
  // clear wiped[] and the wavefront marks, then classify windows
  k_zero_range<<<dim3(min(cdiv(nc + nwave, 256), 128u), npages), 256, 0, st>>>(pages, gp.off + 2 * nc, nc, wave_off, nwave);
  k_gray_windows<<<dim3(min(cdiv(nw, 128), 256u), npages), 128, 0, st>>>(pages, gp, white_off, wflag_off, wave_off);
  k_gray_cascade<<<npages, 256, 0, st>>>(pages, gp, wflag_off, wave_off);
  dim3 g2(min(cdiv(nc, 256), 256u), npages);
  k_gray_wipe<<<g2, 256, 0, st>>>(pages, gp, white_off);
}
}
