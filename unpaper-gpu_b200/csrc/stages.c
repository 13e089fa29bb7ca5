/* stages.c — plans (static device tables derived from geometry + parameters)
 * and stage functions (the kernel sequence of one pipeline step for a group
 * of pages).  Shared by the vtable (group of one) and the sheet engine. */
#define _GNU_SOURCE
#include <math.h>
#include <string.h>

#include "host.h"

#include <libavutil/pixfmt.h>

int b200_fmt_to_dev(int f) {
  switch (f) {
  case AV_PIX_FMT_GRAY8: return DF_GRAY8;
  case AV_PIX_FMT_Y400A: return DF_Y400A;
  case AV_PIX_FMT_RGB24: return DF_RGB24;
  case AV_PIX_FMT_MONOWHITE: return DF_MONOWHITE;
  case AV_PIX_FMT_MONOBLACK: return DF_MONOBLACK;
  default: return -1;
  }
}

int b200_fmt_row_bytes(int f, int width) {
  switch (f) {
  case AV_PIX_FMT_GRAY8: return width;
  case AV_PIX_FMT_Y400A: return 2 * width;
  case AV_PIX_FMT_RGB24: return 3 * width;
  case AV_PIX_FMT_MONOWHITE:
  case AV_PIX_FMT_MONOBLACK: return (width + 7) / 8;
  default: return -1;
  }
}

void *blob_upload(const void *host, size_t bytes) {
  if (bytes == 0 || !host) return b200_dev_alloc(4);   /* an empty table: a valid pointer nobody reads */
  void *d = b200_dev_alloc(bytes);
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return d;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static int igcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

void scratch_need_all(ScratchNeed *n, int w, int h, int fmt) {
  int bpp = fmt == DF_GRAY8 ? 1 : fmt == DF_Y400A ? 2 : fmt == DF_RGB24 ? 3 : 0;
  int aw = w + 64, ah = h + 64;
  n->aux_pitch = bpp ? ((aw * bpp + 15) & ~15) : (((aw + 7) / 8 + 15) & ~15);
  n->aux_h = ah;
  n->aux_bytes = (size_t)n->aux_pitch * ah + 64;
  n->cls_bytes = (size_t)w * h + 64;
  n->list_cap = imax(4096, (w * h) / 4);   /* noisefilter: see nf_list_cap() */
  n->u32_cap = 0;
  n->stack_cap = 1 << 16;
}

/* Entries of the noisefilter's list of pixels that are decided in raster order.  Up to
 * intensity 15 only pixels of small clusters and of the left/top band land there (a quarter
 * of the image is far more than a scan holds); above that every dark pixel does
 * (k_filters.cu: all_mutable), so the list must be able to hold the whole image. */
int nf_list_cap(int w, int h, uint64_t intensity) {
  long long px = (long long)w * h;
  if (intensity > 15) return (int)(px < 0x7fffffffLL ? px : 0x7fffffffLL);
  return imax(4096, (int)(px / 4));
}

/* ---- blackfilter plan: reference filters.c:49-127 loop structure ---------- */

static bool pt_in(int x, int y, int w, int h) { return x >= 0 && x < w && y >= 0 && y < h; }
static bool pt_in_rect_n(int x, int y, Rectangle r) {
  int ax = imin(r.vertex[0].x, r.vertex[1].x), bx = imax(r.vertex[0].x, r.vertex[1].x);
  int ay = imin(r.vertex[0].y, r.vertex[1].y), by = imax(r.vertex[0].y, r.vertex[1].y);
  return x >= ax && x <= bx && y >= ay && y <= by;
}
/* rectangles_overlap (primitives.c:117-123): only the first rectangle's corners are tested */
static bool excluded(DRect a, const BlackfilterParameters *p) {
  for (size_t i = 0; i < p->exclusions_count; i++)
    if (pt_in_rect_n(a.x0, a.y0, p->exclusions[i]) || pt_in_rect_n(a.x1, a.y1, p->exclusions[i])) return true;
  return false;
}

typedef struct { DBfPos *v; int n, cap; } PosVec;
static void pos_push(PosVec *pv, DBfPos q) {
  if (pv->n == pv->cap) { pv->cap = pv->cap ? pv->cap * 2 : 1024; pv->v = (DBfPos *)realloc(pv->v, (size_t)pv->cap * sizeof(DBfPos)); }
  pv->v[pv->n++] = q;
}

typedef struct { int a, b, axis; } Band;

static int bf_scan_emulate(PosVec *pv, Band *bands, int *nbands, int maxbands, int w, int h,
                           const BlackfilterParameters *p, int step_x, int step_y, int sw, int sh,
                           int shift_x, int shift_y, int axis) {
  if ((step_x == 0 && step_y == 0) || sw <= 0 || sh <= 0) return -1;
  if ((step_x < 0) || (step_y < 0)) return -1;
  DRect a = {0, 0, sw - 1, sh - 1};
  long guard = 0;
  while (pt_in(a.x0, a.y0, w, h)) {
    if (!pt_in(a.x1, a.y1, w, h)) {
      int dx = (w - 1) - a.x1, dy = (h - 1) - a.y1;
      a.x0 += dx; a.x1 += dx; a.y0 += dy; a.y1 += dy;
    }
    do {
      if (!excluded(a, p)) {
        int ka = axis == 0 ? a.y0 : a.x0, kb = axis == 0 ? a.y1 : a.x1;
        int bi = -1;
        for (int i = 0; i < *nbands; i++) if (bands[i].a == ka && bands[i].b == kb && bands[i].axis == axis) { bi = i; break; }
        if (bi < 0) { if (*nbands >= maxbands) return -2; bi = (*nbands)++; bands[bi].a = ka; bands[bi].b = kb; bands[bi].axis = axis; }
        DBfPos q; q.r = a; q.sum_off = bi; q.axis = axis;   /* sum_off patched to an offset below */
        pos_push(pv, q);
      }
      a.x0 += step_x; a.x1 += step_x; a.y0 += step_y; a.y1 += step_y;
      if (++guard > 50000000L) return -3;
    } while (pt_in(a.x0, a.y0, w, h));
    a.x0 += shift_x; a.x1 += shift_x; a.y0 += shift_y; a.y1 += shift_y;
  }
  return 0;
}

int bf_plan_build(BfPlan *pl, int w, int h, const BlackfilterParameters *p, int abt) {
  memset(pl, 0, sizeof(*pl));
  PosVec pv = {0};
  enum { MAXB = 256 };
  Band bands[MAXB];
  int nb = 0, rc = 0;
  if (p->scan_direction.horizontal)
    rc = bf_scan_emulate(&pv, bands, &nb, MAXB, w, h, p, p->scan_step.horizontal, 0, p->scan_size.width,
                         (int)p->scan_depth.vertical, 0, (int)p->scan_depth.vertical, 0);
  if (rc == 0 && p->scan_direction.vertical)
    rc = bf_scan_emulate(&pv, bands, &nb, MAXB, w, h, p, 0, p->scan_step.vertical, (int)p->scan_depth.horizontal,
                         p->scan_size.height, (int)p->scan_depth.horizontal, 0, 1);
  if (rc != 0) { free(pv.v); b200_set_error("blackfilter: unsupported scan parameters (%d)", rc); return -1; }
  int stride = imax(w, h);
  pl->njobs = nb;
  pl->jobs_host = (DLineJob *)calloc((size_t)imax(nb, 1), sizeof(DLineJob));
  for (int i = 0; i < nb; i++) {
    DLineJob *j = &pl->jobs_host[i];
    j->axis = bands[i].axis;
    j->out_off = i * stride;
    if (bands[i].axis == 0) { j->xa = 0; j->xb = w - 1; j->ya = imax(bands[i].a, 0); j->yb = imin(bands[i].b, h - 1); }
    else { j->ya = 0; j->yb = h - 1; j->xa = imax(bands[i].a, 0); j->xb = imin(bands[i].b, w - 1); }
  }
  for (int k = 0; k < pv.n; k++) pv.v[k].sum_off = pv.v[k].sum_off * stride;
  pl->npos = pv.n;
  pl->sums_len = nb * stride;
  pl->flag_off = pl->sums_len;
  /* [flags: n bytes][ordered candidate list: n u32] */
  pl->u32_need = pl->flag_off + (pv.n + 3) / 4 + 1 + pv.n + 1;
  pl->pos_dev = (DBfPos *)blob_upload(pv.v, (size_t)pv.n * sizeof(DBfPos));
  pl->jobs_dev = (DLineJob *)blob_upload(pl->jobs_host, (size_t)nb * sizeof(DLineJob));
  free(pv.v);
  pl->abs_threshold = p->abs_threshold;
  pl->intensity = p->intensity;   /* int32 -> uint64 at the flood_fill call (filters.c:87-88) */
  pl->mask_hi = abt;
  return 0;
}
void bf_plan_free(BfPlan *pl) {
  if (pl->pos_dev) b200_dev_free(pl->pos_dev);
  if (pl->jobs_dev) b200_dev_free(pl->jobs_dev);
  free(pl->jobs_host);
  memset(pl, 0, sizeof(*pl));
}

/* ---- blurfilter plan (filters.c:149-232) --------------------------------- */
int blur_plan_build(BlurPlan *pl, int w, int h, const BlurfilterParameters *p, int abs_white) {
  memset(pl, 0, sizeof(*pl));
  int bw = p->scan_size.width, bh = p->scan_size.height;
  if (bw <= 0 || bh <= 0) { b200_set_error("blurfilter: bad scan size"); return -1; }
  int n = w / bw;
  int nrows = (h >= bh) ? (h - bh) / bh + 1 : 0;
  pl->n = n; pl->nrows = nrows; pl->bw = bw; pl->bh = bh;
  pl->nrects = n + nrows * (n + 1);
  DRect *r = (DRect *)calloc((size_t)imax(pl->nrects, 1), sizeof(DRect));
  int k = 0;
  for (int j = 0; j < n; j++) r[k++] = (DRect){j * bw, 0, j * bw + bw - 1, bh - 1};
  for (int row = 0; row < nrows; row++)
    for (int j = 0; j <= n; j++) {
      int y = row * bh + p->scan_step.vertical;
      r[k++] = (DRect){j * bw, y, j * bw + bw - 1, y + bh - 1};
    }
  pl->rects_dev = (DRect *)blob_upload(r, (size_t)pl->nrects * sizeof(DRect));
  free(r);
  pl->cnt_off = 0;
  pl->state_off = pl->nrects;
  pl->flag_off = pl->state_off + 3 * (n + 2);
  pl->u32_need = pl->flag_off + nrows * n + 1;
  pl->T = (unsigned long long)(bw * bh);
  pl->intensity = p->intensity;
  pl->white = abs_white;
  return 0;
}
void blur_plan_free(BlurPlan *pl) {
  if (pl->rects_dev) b200_dev_free(pl->rects_dev);
  memset(pl, 0, sizeof(*pl));
}

/* ---- grayfilter plan (filters.c:370-402) ----------------------------------- */
int gray_plan_build(GrayPlan *pl, int w, int h, const GrayfilterParameters *p, int abt) {
  memset(pl, 0, sizeof(*pl));
  int sw = p->scan_size.width, sh = p->scan_size.height, st_h = p->scan_step.horizontal, st_v = p->scan_step.vertical;
  if (sw <= 0 || sh <= 0 || st_h <= 0 || st_v <= 0) { b200_set_error("grayfilter: bad scan size/step"); return -1; }
  int gx = igcd(sw, st_h), gy = igcd(sh, st_v);
  int wcx = sw / gx, wcy = sh / gy, scx = st_h / gx, scy = st_v / gy;
  int nwx = (w + st_h - 1) / st_h + 1;   /* x = 0, step, ... up to the first value >= width */
  int nwy = h / st_v + 1;                /* y = 0, step, ... while y <= height */
  int ncx = (nwx - 1) * scx + wcx, ncy = (nwy - 1) * scy + wcy;
  int skew = (sw + st_h - 1) / st_h;
  int v[18] = {gx, gy, ncx, ncy, wcx, wcy, scx, scy, nwx, nwy, skew, sw, sh, st_h, st_v,
               p->abs_threshold, abt == 255, 0};
  memcpy(pl->gp, v, sizeof(v));
  long long nc = (long long)ncx * ncy;
  if (nc * 4 > 0x7fffffffLL / 2) { b200_set_error("grayfilter: cell grid too large"); return -1; }
  pl->white_off = (int)(3 * nc);
  /* [dark][light][wiped][white][window flags][wavefront marks] */
  pl->u32_need = (int)(4 * nc) + nwx * nwy + nwx + skew * (nwy - 1) + 4;
  pl->dark_max = abt;
  pl->ok = 1;
  /* k_cellstats keeps two counters per cell of a cell row in shared memory */
  if ((size_t)ncx * 2 * sizeof(unsigned) > 200 * 1024) { b200_set_error("grayfilter: cell row too wide for this scan size/step"); return -1; }
  return 0;
}

/* ---- detect_masks plan (masks.c:54-100) ----------------------------------- */
int mask_plan_build(MaskPlan *pl, int w, int h, const MaskDetectionParameters *p, const Point *pts, int npts) {
  memset(pl, 0, sizeof(*pl));
  pl->p = *p;
  pl->stride = imax(w, h);
  pl->max_points = npts;
  pl->jobs_host = (DLineJob *)calloc((size_t)imax(2 * npts, 1), sizeof(DLineJob));
  int nj = 0;
  for (int i = 0; i < npts; i++) {
    if (p->scan_direction.horizontal) {
      int depth = p->scan_depth.horizontal == -1 ? h : p->scan_depth.horizontal;
      int l0 = pts[i].y + (-depth / 2), l1 = l0 + depth - 1;
      pl->jobs_host[nj++] = (DLineJob){0, w - 1, imax(l0, 0), imin(l1, h - 1), 0, (i * 2 + 0) * pl->stride};
    }
    if (p->scan_direction.vertical) {
      int depth = p->scan_depth.vertical == -1 ? w : p->scan_depth.vertical;
      int l0 = pts[i].x + (-depth / 2), l1 = l0 + depth - 1;
      pl->jobs_host[nj++] = (DLineJob){imax(l0, 0), imin(l1, w - 1), 0, h - 1, 1, (i * 2 + 1) * pl->stride};
    }
  }
  pl->njobs = nj;
  pl->jobs_dev = (DLineJob *)blob_upload(pl->jobs_host, (size_t)nj * sizeof(DLineJob));
  pl->u32_need = npts * 2 * pl->stride + 1;
  return 0;
}
void mask_plan_free(MaskPlan *pl) {
  if (pl->jobs_dev) b200_dev_free(pl->jobs_dev);
  free(pl->jobs_host);
  memset(pl, 0, sizeof(*pl));
}

/* ---- detect_border plan (masks.c:391-448) --------------------------------- */
int border_plan_build(BorderPlan *pl, int w, int h, const BorderScanParameters *p,
                      const Rectangle *outside, int n, int abt) {
  memset(pl, 0, sizeof(*pl));
  pl->p = *p;
  pl->stride = imax(w, h);
  pl->abt = abt;
  pl->oob_dark = abt == 255;
  pl->jobs_host = (DLineJob *)calloc((size_t)imax(2 * n, 1), sizeof(DLineJob));
  int nj = 0;
  for (int i = 0; i < n; i++) {
    int xa = imax(outside[i].vertex[0].x, 0), xb = imin(outside[i].vertex[1].x, w - 1);
    int ya = imax(outside[i].vertex[0].y, 0), yb = imin(outside[i].vertex[1].y, h - 1);
    if (p->scan_direction.horizontal) pl->jobs_host[nj++] = (DLineJob){xa, xb, ya, yb, 0, (i * 2 + 0) * pl->stride};
    if (p->scan_direction.vertical) pl->jobs_host[nj++] = (DLineJob){xa, xb, ya, yb, 1, (i * 2 + 1) * pl->stride};
  }
  pl->njobs = nj;
  pl->jobs_dev = (DLineJob *)blob_upload(pl->jobs_host, (size_t)nj * sizeof(DLineJob));
  pl->u32_need = n * 2 * pl->stride + 1;
  return 0;
}
void border_plan_free(BorderPlan *pl) {
  if (pl->jobs_dev) b200_dev_free(pl->jobs_dev);
  free(pl->jobs_host);
  memset(pl, 0, sizeof(*pl));
}

/* ---- detect_rotation plan (deskew.c:148-245) ------------------------------ */
static float rot_pair_value(float a, float b, int count, float max_dev) {
  float rotation[2] = {a, b};
  float total = 0.0;
  for (int i = 0; i < count; i++) total += rotation[i];
  float average = total / count;
  total = 0.0;
  for (int i = 0; i < count; i++) total += powf(rotation[i] - average, 2);
  float deviation = sqrtf(total);
  return deviation <= max_dev ? average : 0.0f;
}

int rot_plan_build(RotPlan *pl, int w, int h, const DeskewParameters *p, int max_masks, bool with_pair) {
  memset(pl, 0, sizeof(*pl));
  pl->p = *p;
  if (!(p->deskewScanStepRad > 0.0f)) { b200_set_error("deskew: scan step must be > 0"); return -1; }
  int cap = 8192, n = 0;
  pl->rot_host = (float *)malloc(sizeof(float) * cap);
  pl->tan_host = (float *)malloc(sizeof(float) * cap);
  /* deskew.c:156-160, evaluated in float exactly as written there */
  for (float rotation = 0.0; rotation <= p->deskewScanRangeRad;
       rotation = (rotation >= 0.0) ? -(rotation + p->deskewScanStepRad) : -rotation) {
    if (n >= cap) { b200_set_error("deskew: too many scan angles"); free(pl->rot_host); free(pl->tan_host); return -1; }
    pl->rot_host[n] = rotation;
    pl->tan_host[n] = tanf(rotation);
    n++;
  }
  pl->nangles = n;
  pl->rot_dev = (float *)blob_upload(pl->rot_host, sizeof(float) * n);
  pl->tan_dev = (float *)blob_upload(pl->tan_host, sizeof(float) * n);
  pl->edges[0] = p->scan_edges.left; pl->edges[1] = p->scan_edges.top;
  pl->edges[2] = p->scan_edges.right; pl->edges[3] = p->scan_edges.bottom;
  pl->host_tail = (pl->edges[0] != 0) + (pl->edges[1] != 0) + (pl->edges[2] != 0) + (pl->edges[3] != 0) > 2 || !(with_pair && n <= 512);
  pl->peak_off = 0;
  pl->u32_need = max_masks * 4 * n + 1;
  int sc = p->deskewScanSize == -1 ? imax(w, h) : p->deskewScanSize;
  pl->scan_cap = imax(1, imin(sc, 10000));
  /* rows of the column-prefix table: the scan length is at most the mask height (<= h + 1) */
  pl->pre_need = (long long)(imin(pl->scan_cap, h + 1) + 2) * w;
  /* a scan line changes column at most every 1/|tan| rows */
  float maxtan = 0.0f;
  for (int i = 0; i < n; i++) if (fabsf(pl->tan_host[i]) > maxtan) maxtan = fabsf(pl->tan_host[i]);
  long long runs = (long long)(maxtan * imin(pl->scan_cap, imax(w, h) + 1)) + 4;
  pl->run_cap = (runs <= 700) ? (int)runs : 0;   /* 2 x 8 x run_cap ints of shared memory */
  if (with_pair && n <= 512) {
    int m = 2 * n;
    float *t = (float *)malloc(sizeof(float) * 4 * (size_t)m * m);
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) {
        float a = i < n ? pl->rot_host[i] : -pl->rot_host[i - n];
        float b = j < n ? pl->rot_host[j] : -pl->rot_host[j - n];
        float r = (i == j) ? rot_pair_value(a, b, 1, p->deskewScanDeviationRad)
                           : rot_pair_value(a, b, 2, p->deskewScanDeviationRad);
        float *e = t + ((size_t)i * m + j) * 4;
        e[0] = r; e[1] = sinf(-r); e[2] = cosf(-r); e[3] = 0.0f;
      }
    /* the diagonal doubles as the single-edge table; make sure a genuine pair
     * (i,i) gives the same value (it does: (a+a)/2 == a exactly) */
    pl->pair_dev = (float *)blob_upload(t, sizeof(float) * 4 * (size_t)m * m);
    free(t);
  }
  return 0;
}
void rot_plan_free(RotPlan *pl) {
  if (pl->rot_dev) b200_dev_free(pl->rot_dev);
  if (pl->tan_dev) b200_dev_free(pl->tan_dev);
  if (pl->pair_dev) b200_dev_free(pl->pair_dev);
  free(pl->rot_host); free(pl->tan_host);
  memset(pl, 0, sizeof(*pl));
}

float rot_finalize_host(const RotPlan *pl, const int angle_idx[4]) {
  float rotation[4];
  int count = 0;
  for (int e = 0; e < 4; e++) {
    if (!pl->edges[e]) continue;
    float r = angle_idx[e] >= 0 ? pl->rot_host[angle_idx[e]] : 0.0f;
    rotation[count++] = (e == 1 || e == 3) ? -r : r;
  }
  /* deskew.c:218-240 */
  float total = 0.0;
  for (int i = 0; i < count; i++) total += rotation[i];
  float average = total / count;
  total = 0.0;
  for (int i = 0; i < count; i++) total += powf(rotation[i] - average, 2);
  float deviation = sqrtf(total);
  if (deviation <= pl->p.deskewScanDeviationRad) return average;
  return 0.0;
}

/* ---- stages ---------------------------------------------------------------- */

static int bppf(int fmt) { return fmt == DF_GRAY8 ? 1 : fmt == DF_Y400A ? 2 : fmt == DF_RGB24 ? 3 : 1; }

void stage_blackfilter(StageCtx *c, const BfPlan *pl) {
  if (pl->npos <= 0) return;
  b200k_zero_u32(c->st, c->pages, c->npages, 0, pl->sums_len);
  b200k_linesums(c->st, c->pages, c->npages, pl->jobs_dev, pl->jobs_host, pl->njobs, ST_MAXCH, 0, 0,
                 c->fmt == DF_GRAY8 && c->rows_aligned16, c->w, c->h, 0);
  b200k_bf_scan(c->st, c->pages, c->npages, pl->pos_dev, pl->npos, pl->abs_threshold, pl->intensity, 0,
                pl->mask_hi, pl->flag_off, c->h);
  c->launches += 4;
}

int stage_noisefilter(StageCtx *c, uint64_t intensity, int white) {
  if (c->h >= 32768 || c->w >= 65536) { b200_set_error("noisefilter: image too large"); return -1; }
  int rc = b200k_noisefilter(c->st, c->pages, c->npages, c->w, c->h, c->fmt, intensity, white, c->rows_aligned16 ? 1 : 0);
  if (rc) b200_set_error("noisefilter: unsupported intensity");
  c->launches += 2;
  return rc;
}

void stage_blurfilter(StageCtx *c, const BlurPlan *pl) {
  if (pl->n <= 0 || pl->nrows <= 0) return;
  b200k_rect_count(c->st, c->pages, c->npages, pl->rects_dev, pl->nrects, 0, pl->white, pl->cnt_off);
  b200k_blur_decide(c->st, c->pages, c->npages, pl->n, pl->nrows, pl->T, pl->intensity, pl->cnt_off,
                    pl->state_off, pl->flag_off);
  b200k_blur_wipe(c->st, c->pages, c->npages, pl->n, pl->nrows, pl->bw, pl->bh, pl->flag_off);
  c->launches += 3;
}

int stage_grayfilter(StageCtx *c, const GrayPlan *pl) {
  const int *g = pl->gp;
  int rc = b200k_cellstats(c->st, c->pages, c->npages, g[0], g[1], g[2], g[3], pl->dark_max, 0);
  if (rc) { b200_set_error("grayfilter: cell row too wide"); return rc; }
  b200k_gray_cascade(c->st, c->pages, c->npages, pl->gp, pl->white_off);
  c->launches += 6;
  return 0;
}

void stage_detect_masks(StageCtx *c, const MaskPlan *pl) {
  const MaskDetectionParameters *p = &pl->p;
  if (pl->njobs > 0) {
    b200k_zero_u32(c->st, c->pages, c->npages, 0, pl->u32_need);
    /* the pass over the sheet also leaves the ink map the next rotation wants (engine only) */
    c->ink_fresh = b200k_linesums(c->st, c->pages, c->npages, pl->jobs_dev, pl->jobs_host, pl->njobs, ST_GRAY, 0, 0,
                                  c->fmt == DF_GRAY8 && c->rows_aligned16, c->w, c->h, c->want_ink);
    c->launches += 3;
  }
  int size[2] = {p->scan_size.width, p->scan_size.height};
  int depth[2] = {p->scan_depth.horizontal, p->scan_depth.vertical};
  int step[2] = {p->scan_step.horizontal, p->scan_step.vertical};
  float thr[2] = {p->scan_threshold.horizontal, p->scan_threshold.vertical};
  b200k_detect_masks(c->st, c->pages, c->npages, pl->max_points, size, depth, step, thr,
                     p->scan_direction.horizontal, p->scan_direction.vertical, 0, pl->stride,
                     p->minimum_width, p->maximum_width, p->minimum_height, p->maximum_height);
  c->launches += 2;
}

/* detect_rotation_cpu's float tail (deskew.c:218-240) needs the host's libm when no
 * pair table covers the case (3-4 scan edges, or more than 512 angles): the angle
 * indices come back through pinned memory, a stream-ordered host function evaluates
 * rot_finalize_host + sinf/cosf, the values go back up.  No host thread blocks. */
static void CUDART_CB rot_host_tail(void *arg) {
  RotHostJob *j = (RotHostJob *)arg;
  for (int p = 0; p < j->n; p++) {
    float r = rot_finalize_host(j->pl, j->pulled[p].rot_angle_idx[j->mi]);
    float *t = j->tab + 4 * p;
    t[0] = r; t[1] = sinf(-r); t[2] = cosf(-r); t[3] = 0.0f;   /* deskew.c:260-261 */
  }
}

/* detect_rotation() of mask `mi` on every page of the group */
int stage_detect_rotation_mask(StageCtx *c, const RotPlan *pl, int mi) {
  int rc = b200k_rot_peaks(c->st, c->pages, c->npages, mi, 1, pl->tan_dev, pl->nangles,
                           pl->p.deskewScanSize, pl->p.deskewScanDepth, pl->edges, pl->peak_off, pl->scan_cap,
                           c->w, 1, pl->run_cap);
  if (rc) { b200_set_error("deskew: scan size too large"); return rc; }
  b200k_rot_finalize(c->st, c->pages, c->npages, pl->rot_dev, pl->pair_dev, pl->nangles, pl->edges,
                     pl->peak_off, pl->p.deskewScanDeviationRad, mi, 1);
  c->launches += ((pl->edges[0] || pl->edges[2]) ? 2 : 0) + ((pl->edges[1] || pl->edges[3]) ? 1 : 0) + 1;
  if (pl->host_tail && c->rot_jobs) {
    RotHostJob *j = &c->rot_jobs[mi];
    j->pl = pl; j->pulled = c->rot_pull; j->tab = c->rot_tab_host; j->n = c->npages; j->mi = mi;
    CUDA_OK(cudaMemcpyAsync(c->rot_pull, c->pages, sizeof(DPage) * c->npages, cudaMemcpyDeviceToHost, c->st));
    CUDA_OK(cudaLaunchHostFunc(c->st, rot_host_tail, j));
    CUDA_OK(cudaMemcpyAsync(c->rot_tab_dev, c->rot_tab_host, sizeof(float) * 4 * c->npages, cudaMemcpyHostToDevice, c->st));
    b200k_rot_set_sincos(c->st, c->pages, c->npages, mi, c->rot_tab_dev);
    c->launches += 1;
  }
  return 0;
}

int stage_detect_rotation(StageCtx *c, const RotPlan *pl, int max_masks) {
  for (int mi = 0; mi < max_masks; mi++) {
    int rc = stage_detect_rotation_mask(c, pl, mi);
    if (rc) return rc;
  }
  return 0;
}

void stage_deskew_mask(StageCtx *c, int interp, int mi) {
  int aw = c->w + 64, ah = c->h + 64;
  b200k_rotate(c->st, c->pages, c->npages, mi, interp, aw, ah, c->copyA);
  b200k_copy_jobs(c->st, c->copyA, c->npages, aw * bppf(c->fmt), ah);
  c->launches += 2 + (interp == 2);
}

void stage_deskew(StageCtx *c, int interp, int max_masks) {
  for (int mi = 0; mi < max_masks; mi++) stage_deskew_mask(c, interp, mi);
}

/* ---- sheet-engine forms: one sweep img -> other, then the buffers change roles ---- */

void stage_deskew_mask_pass(StageCtx *c, int interp, int mi) {
  b200k_rotate_sheet(c->st, c->pages, c->npages, mi, interp, c->fmt, c->w, c->h, c->ink_fresh);
  b200k_swap_sheets(c->st, c->pages, c->npages);
  c->parity ^= 1;
  c->launches += 2 + (interp == 2 && !c->ink_fresh);
  c->ink_fresh = 0;            /* the sheet has changed */
}

static void move_pass(StageCtx *c, Pixel mask_color) {
  b200k_move_pass(c->st, c->pages, c->npages, c->w * bppf(c->fmt), c->h, mask_color.r, mask_color.g, mask_color.b);
  b200k_swap_sheets(c->st, c->pages, c->npages);
  c->parity ^= 1;
  c->launches += 2;
}

void stage_center_masks_pass(StageCtx *c, int max_masks) {
  Pixel none = {0, 0, 0};
  for (int i = 0; i < max_masks; i++) {
    b200k_prep_center_move(c->st, c->pages, c->npages, i);
    c->launches += 1;
    move_pass(c, none);
  }
}

/* apply_masks(border masks) + align_mask() per outside area (sheet_stages.c:474-483); the
 * mask painting rides on the first move */
/* final_dst != NULL: the last sweep renders sheet p into final_dst + p * final_stride (tight rows equal to
 * the sheet's pitch) — the output stage then has nothing left to copy */
void stage_align_masks_pass(StageCtx *c, const MaskAlignmentParameters *p, int n_outside, Pixel mask_color,
                            uint8_t *final_dst, size_t final_stride) {
  for (int i = 0; i < n_outside; i++) {
    b200k_prep_align_move(c->st, c->pages, c->npages, i, p->alignment.left, p->alignment.top, p->alignment.right,
                          p->alignment.bottom, p->margin.horizontal, p->margin.vertical, i == 0);
    c->launches += 1;
    if (final_dst && i == n_outside - 1) { b200k_set_other(c->st, c->pages, c->npages, final_dst, final_stride); c->launches += 1; }
    move_pass(c, mask_color);
  }
}

void stage_shift_pass(StageCtx *c, Delta d) {
  Pixel none = {0, 0, 0};
  b200k_prep_shift_move(c->st, c->pages, c->npages, d.horizontal, d.vertical);
  c->launches += 1;
  move_pass(c, none);
}

static void run_move(StageCtx *c) {
  int aw = c->w + 64, ah = c->h + 64;
  b200k_fill_jobs(c->st, c->fillA, c->npages, aw, ah);
  b200k_copy_jobs(c->st, c->copyA, c->npages, aw * bppf(c->fmt), ah);
  b200k_fill_jobs(c->st, c->fillB, c->npages, aw, ah);
  b200k_fill_jobs(c->st, c->fillC, c->npages, aw, ah);
  b200k_copy_jobs(c->st, c->copyB, c->npages, aw * bppf(c->fmt), ah);
  c->launches += 5;
}

void stage_center_masks(StageCtx *c, int max_masks) {
  for (int i = 0; i < max_masks; i++) {
    b200k_prep_center(c->st, c->pages, c->npages, i, c->fillA, c->copyA, c->fillB, c->copyB, c->fillC);
    c->launches += 1;
    run_move(c);
  }
}

void stage_detect_border(StageCtx *c, const BorderPlan *pl) {
  const BorderScanParameters *p = &pl->p;
  if (pl->njobs > 0) {
    b200k_zero_u32(c->st, c->pages, c->npages, 0, pl->u32_need);
    b200k_linesums(c->st, c->pages, c->npages, pl->jobs_dev, pl->jobs_host, pl->njobs, ST_COUNT_GRAY_RANGE, 0, pl->abt,
                   c->fmt == DF_GRAY8 && c->rows_aligned16, c->w, c->h, 0);
    c->launches += 3;
  }
  b200k_detect_border(c->st, c->pages, c->npages, p->scan_size.width, p->scan_size.height,
                      p->scan_step.horizontal, p->scan_step.vertical, p->scan_threshold.horizontal,
                      p->scan_threshold.vertical, p->scan_direction.horizontal, p->scan_direction.vertical,
                      0, pl->stride, pl->oob_dark);
  c->launches += 2;
}

void stage_apply_border_masks(StageCtx *c, Pixel color) {
  b200k_prep_border_maskjob(c->st, c->pages, c->npages, c->maskJ, color.r, color.g, color.b);
  b200k_apply_masks(c->st, c->maskJ, c->npages, c->w, c->h);
  c->launches += 2;
}

void stage_align_masks(StageCtx *c, const MaskAlignmentParameters *p, int n_outside) {
  for (int i = 0; i < n_outside; i++) {
    b200k_prep_align(c->st, c->pages, c->npages, i, p->alignment.left, p->alignment.top, p->alignment.right,
                     p->alignment.bottom, p->margin.horizontal, p->margin.vertical, c->fillA, c->copyA,
                     c->fillB, c->copyB, c->fillC);
    c->launches += 1;
    run_move(c);
  }
}
