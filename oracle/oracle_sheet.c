/* oracle_sheet.c — TEST INFRASTRUCTURE: host-buffer entry points of the CPU
 * restatement (`orc_host_*`, same signatures as `unpaper_b200_host_*`) and the
 * restated process_sheet() stage order (reference src/core/sheet_stages.c:44-534)
 * as orc_process_sheets().  See oracle_ops.c for the pinning. */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static OImg wrap(const B200HostImage *h) {
  OImg im = {h->data, h->width, h->height, h->linesize, h->format, {h->background.r, h->background.g, h->background.b}, h->abs_black_threshold};
  return im;
}
static ORect R(const Rectangle *r) { ORect o = {r->vertex[0].x, r->vertex[0].y, r->vertex[1].x, r->vertex[1].y}; return o; }
static Rectangle U_(ORect r) { Rectangle o = {{{r.x0, r.y0}, {r.x1, r.y1}}}; return o; }
static OPx P(Pixel c) { OPx p = {c.r, c.g, c.b}; return p; }
static int row_bytes(int fmt, int w) { return fmt == B200_FMT_GRAY8 ? w : fmt == B200_FMT_Y400A ? 2 * w : fmt == B200_FMT_RGB24 ? 3 * w : (w + 7) / 8; }
static int export_img(OImg *r, B200HostImage *out) {
  int rc = 0;
  if (r->w != out->width || r->h != out->height) rc = -2;
  else for (int y = 0; y < r->h; y++) memcpy(out->data + (size_t)y * out->linesize, r->d + (size_t)y * r->ls, (size_t)row_bytes(r->fmt, r->w));
  o_free(r);
  return rc;
}

int orc_host_wipe_rectangle(B200HostImage *img, const Rectangle *a, Pixel c) { OImg im = wrap(img); o_wipe(&im, R(a), P(c)); return 0; }
int orc_host_copy_rectangle(const B200HostImage *s, B200HostImage *d, const Rectangle *a, Point t) { OImg si = wrap(s), di = wrap(d); o_copy(&si, &di, R(a), t.x, t.y); return 0; }
int orc_host_center_image(const B200HostImage *s, B200HostImage *d, Point o, RectangleSize z) { OImg si = wrap(s), di = wrap(d); o_center_image(&si, &di, o.x, o.y, z.width, z.height); return 0; }
int orc_host_stretch(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  OImg s = wrap(img);
  if (s.w == out->width && s.h == out->height) { OImg c = o_new(s.w, s.h, s.fmt, &s); for (int y = 0; y < s.h; y++) memcpy(c.d + (size_t)y * c.ls, s.d + (size_t)y * s.ls, (size_t)row_bytes(s.fmt, s.w)); return export_img(&c, out); }
  OImg r = o_stretch(&s, out->width, out->height, interp); return export_img(&r, out);
}
int orc_host_resize(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  OImg s = wrap(img);
  OImg tight = o_new(s.w, s.h, s.fmt, &s);
  for (int y = 0; y < s.h; y++) memcpy(tight.d + (size_t)y * tight.ls, s.d + (size_t)y * s.ls, (size_t)row_bytes(s.fmt, s.w));
  OImg r;
  if (s.w == out->width && s.h == out->height) r = tight; else { r = o_resize(&tight, out->width, out->height, interp); o_free(&tight); }
  return export_img(&r, out);
}
int orc_host_flip_rotate_90(const B200HostImage *img, B200HostImage *out, int32_t dir) { OImg s = wrap(img); OImg r = o_rotate90(&s, dir); return export_img(&r, out); }
int orc_host_mirror(B200HostImage *img, Direction d) { OImg im = wrap(img); o_mirror(&im, d.horizontal, d.vertical); return 0; }
int orc_host_shift(const B200HostImage *img, B200HostImage *out, Delta d) { OImg s = wrap(img); OImg r = o_shift(&s, d.horizontal, d.vertical); return export_img(&r, out); }
int orc_host_apply_masks(B200HostImage *img, const Rectangle *m, size_t n, Pixel c) {
  OImg im = wrap(img); ORect *r = (ORect *)malloc(sizeof(ORect) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) r[i] = R(&m[i]);
  o_apply_masks(&im, r, n, P(c)); free(r); return 0;
}
int orc_host_apply_wipes(B200HostImage *img, const Wipes *w, Pixel c) {
  OImg im = wrap(img); ORect r[MAX_WIPES];
  for (size_t i = 0; i < w->count; i++) r[i] = R(&w->areas[i]);
  o_apply_wipes(&im, r, w->count, P(c)); return 0;
}
int orc_host_apply_border(B200HostImage *img, const Border *b, Pixel c) { OImg im = wrap(img); o_apply_border(&im, *b, P(c)); return 0; }
int orc_host_detect_masks(const B200HostImage *img, const MaskDetectionParameters *p, const Point *pts, size_t n, Rectangle *out) {
  OImg im = wrap(img); ORect *r = (ORect *)malloc(sizeof(ORect) * (n ? n : 1));
  size_t c = o_detect_masks(&im, p, pts, n, r);
  if (p->scan_direction.horizontal || p->scan_direction.vertical) for (size_t i = 0; i < n; i++) out[i] = U_(r[i]);
  free(r); return (int)c;
}
int orc_host_center_mask(B200HostImage *img, Point c, const Rectangle *a) { OImg im = wrap(img); o_center_mask(&im, c.x, c.y, R(a)); return 0; }
int orc_host_align_mask(B200HostImage *img, const Rectangle *in, const Rectangle *out, const MaskAlignmentParameters *p) { OImg im = wrap(img); o_align_mask(&im, R(in), R(out), p); return 0; }
int orc_host_detect_border(const B200HostImage *img, const BorderScanParameters *p, const Rectangle *o, Border *out) { OImg im = wrap(img); *out = o_detect_border(&im, p, R(o)); return 0; }
int orc_host_blackfilter(B200HostImage *img, const BlackfilterParameters *p) { OImg im = wrap(img); o_blackfilter(&im, p); return 0; }
int orc_host_blurfilter(B200HostImage *img, const BlurfilterParameters *p, uint8_t w) { OImg im = wrap(img); o_blurfilter(&im, p, w); return 0; }
int orc_host_noisefilter(B200HostImage *img, uint64_t i, uint8_t w) { OImg im = wrap(img); o_noisefilter(&im, i, w); return 0; }
int orc_host_grayfilter(B200HostImage *img, const GrayfilterParameters *p) { OImg im = wrap(img); o_grayfilter(&im, p); return 0; }
int orc_host_detect_rotation(const B200HostImage *img, const Rectangle *m, const DeskewParameters *p, float *out) { OImg im = wrap(img); *out = o_detect_rotation(&im, R(m), p); return 0; }
int orc_host_deskew(B200HostImage *img, const Rectangle *m, float rad, int32_t interp) { OImg im = wrap(img); o_deskew(&im, R(m), rad, interp); return 0; }

/* saveImage()'s pixel-format conversion (file.c:197-260).  file.c itself needs
 * libavcodec/libavformat and is not part of oracle/_ref, so for this one function
 * the restatement is pinned only through the A1 golden record (gray < 170 -> black
 * against the reference's goldenA1.pbm, tests/golden/make_golden.py) and, for the
 * generic branch, through copy_rectangle which IS checked against oracle/_ref. */
int orc_host_convert_format(const B200HostImage *in, B200HostImage *out) {
  OImg s = wrap(in), d = wrap(out);
  int w = s.w, h = s.h;
  if (d.w != w || d.h != h) return -1;
  if (s.fmt == d.fmt) {                                 /* file.c:210: no conversion */
    int row = s.fmt == B200_FMT_RGB24 ? 3 * w : s.fmt == B200_FMT_Y400A ? 2 * w : s.fmt == B200_FMT_GRAY8 ? w : (w + 7) / 8;
    for (int y = 0; y < h; y++) memcpy(d.d + (size_t)y * d.ls, s.d + (size_t)y * s.ls, (size_t)row);
    return 0;
  }
  if ((s.fmt == B200_FMT_RGB24 || s.fmt == B200_FMT_GRAY8) && d.fmt == B200_FMT_MONOWHITE) {   /* :215-243 */
    for (int y = 0; y < h; y++) {
      const uint8_t *src = s.d + (size_t)y * s.ls;
      uint8_t *dst = d.d + (size_t)y * d.ls;
      for (int x = 0; x < w; x++) {
        int gray = s.fmt == B200_FMT_RGB24 ? (src[x * 3] + src[x * 3 + 1] + src[x * 3 + 2]) / 3 : src[x];
        int bit = x % 8;
        if (bit == 0) dst[x / 8] = 0;
        if (gray < s.abt) dst[x / 8] |= (0x80 >> bit);
      }
    }
  } else if (s.fmt == B200_FMT_MONOBLACK && d.fmt == B200_FMT_MONOWHITE) {                       /* :244-255 */
    for (int y = 0; y < h; y++)
      for (int x = 0; x < (w + 7) / 8; x++) d.d[(size_t)y * d.ls + x] = s.d[(size_t)y * s.ls + x] ^ 0xFF;
  } else {                                                                                          /* :257-259 */
    ORect full = {0, 0, w - 1, h - 1};
    o_copy(&s, &d, full, 0, 0);
  }
  return 0;
}

/* ---- process_sheet(): decode, pre, filters, masks, deskew, post (sheet_stages.c) ---- */

static int one_sheet(const B200SheetConfig *c, const uint8_t *pages, int pw, int ph, int fmt, uint8_t *out, B200SheetResult *res) {
  int W = pw * c->input_count, H = ph;
  OImg like = {NULL, 0, 0, 0, 0, {c->sheet_background.r, c->sheet_background.g, c->sheet_background.b}, c->abs_black_threshold};
  /* the working sheet is always RGB24 (sheet_stages.c:153-155) */
  OImg sheet = o_new(W, H, B200_FMT_RGB24, &like);
  ORect full = {0, 0, W - 1, H - 1};
  o_wipe(&sheet, full, P(c->sheet_background));
  for (int j = 0; j < c->input_count; j++) {                              /* :158-165 */
    OImg page = {(uint8_t *)pages + (size_t)row_bytes(fmt, pw) * ph * j, pw, ph, row_bytes(fmt, pw), fmt, {255, 255, 255}, c->abs_black_threshold};
    memcpy(page.bg, like.bg, 3);
    o_center_image(&page, &sheet, W * j / c->input_count, 0, W / c->input_count, H);
  }
  memset(res, 0, sizeof(*res));
  res->sheet_width = W; res->sheet_height = H;
  OPx mc = P(c->mask_color);
  /* pre stage (:187-325) */
  Point pts[8]; int npts = c->point_count; memcpy(pts, c->points, sizeof(Point) * 8);
  ORect outside[2]; int nout = 0;
  MaskDetectionParameters mp = c->mask_detection;
  if (c->pre_mirror.horizontal || c->pre_mirror.vertical) o_mirror(&sheet, c->pre_mirror.horizontal, c->pre_mirror.vertical);   /* :200-203 */
  if (c->pre_shift.horizontal != 0 || c->pre_shift.vertical != 0) {                                                              /* :205-208 */
    OImg n = o_shift(&sheet, c->pre_shift.horizontal, c->pre_shift.vertical); o_free(&sheet); sheet = n;
  }
  if (c->pre_mask_count > 0) { ORect r[8]; for (int i = 0; i < c->pre_mask_count; i++) r[i] = R(&c->pre_masks[i]); o_apply_masks(&sheet, r, (size_t)c->pre_mask_count, mc); }
  if (c->layout == LAYOUT_SINGLE) {
    if (npts == 0) pts[npts++] = (Point){W / 2, H / 2};
    if (mp.maximum_width == -1) mp.maximum_width = W;
    if (mp.maximum_height == -1) mp.maximum_height = H;
    outside[nout++] = full;
  } else if (c->layout == LAYOUT_DOUBLE) {
    if (npts == 0) { pts[npts++] = (Point){W / 4, H / 2}; pts[npts++] = (Point){W - W / 4, H / 2}; }
    if (mp.maximum_width == -1) mp.maximum_width = W / 2;
    if (mp.maximum_height == -1) mp.maximum_height = H;
    outside[nout++] = (ORect){0, 0, W / 2, H - 1};
    outside[nout++] = (ORect){W / 2, 0, W - 1, H - 1};
  }
  if (mp.maximum_width == -1) mp.maximum_width = W;
  if (mp.maximum_height == -1) mp.maximum_height = H;
  if (!c->no_wipe) { ORect r[8]; for (int i = 0; i < c->pre_wipe_count; i++) r[i] = R(&c->pre_wipes[i]); o_apply_wipes(&sheet, r, (size_t)c->pre_wipe_count, mc); }
  if (!c->no_border) o_apply_border(&sheet, c->pre_border, mc);
  BlackfilterParameters bf = c->blackfilter;
  Rectangle excl[4]; size_t nex = 0;
  for (size_t i = 0; i < c->blackfilter.exclusions_count && i < 2; i++) excl[nex++] = c->blackfilter.exclusions[i];
  if (nex == 0 && c->layout != LAYOUT_NONE) {                               /* :298-322 */
    if (c->layout == LAYOUT_SINGLE) excl[nex++] = (Rectangle){{{W / 4, H / 4}, {W / 4 + W / 2 - 1, H / 4 + H / 2 - 1}}};
    else {
      int fw = W / 4, fh = H / 2, ox = W / 8, oy = H / 4;
      excl[nex++] = (Rectangle){{{ox, oy}, {ox + fw - 1, oy + fh - 1}}};
      excl[nex++] = (Rectangle){{{ox + W / 2, oy}, {ox + W / 2 + fw - 1, oy + fh - 1}}};
    }
  }
  bf.exclusions = excl; bf.exclusions_count = nex;
  /* filters stage (:327-357) */
  if (!c->no_blackfilter) res->blackfilter_fills = o_blackfilter(&sheet, &bf);
  if (!c->no_noisefilter) res->noise_clusters = (int32_t)o_noisefilter(&sheet, c->noisefilter_intensity, c->abs_white_threshold);
  if (!c->no_blurfilter) o_blurfilter(&sheet, &c->blurfilter, c->abs_white_threshold);
  /* masks stage (:359-388): the first detection's result is discarded */
  ORect masks[8]; size_t nmask = 0;
  if (!c->no_grayfilter) o_grayfilter(&sheet, &c->grayfilter);
  /* deskew stage (:390-418) */
  if (!c->no_deskew) {
    if (!c->no_mask_scan) {
      nmask = o_detect_masks(&sheet, &mp, pts, (size_t)npts, masks);
      res->deskew_mask_count = (int32_t)nmask;
      for (size_t i = 0; i < nmask && i < B200_TRACE_MAX_MASKS; i++) res->deskew_masks[i] = U_(masks[i]);
    }
    for (size_t i = 0; i < nmask; i++) {
      float rot = o_detect_rotation(&sheet, masks[i], &c->deskew);
      if (i < B200_TRACE_MAX_MASKS) res->rotation[i] = rot;
      if (rot != 0.0) o_deskew(&sheet, masks[i], rot, c->interpolate_type);
    }
  }
  /* post stage (:420-534) */
  if (!c->no_mask_center) {
    if (!c->no_mask_scan) {
      nmask = o_detect_masks(&sheet, &mp, pts, (size_t)npts, masks);
      res->center_mask_count = (int32_t)nmask;
      for (size_t i = 0; i < nmask && i < B200_TRACE_MAX_MASKS; i++) res->center_masks[i] = U_(masks[i]);
    }
    for (size_t i = 0; i < nmask; i++) { int did = o_center_mask(&sheet, pts[i].x, pts[i].y, masks[i]); if (i < B200_TRACE_MAX_MASKS) res->centered[i] = did; }
  }
  if (!c->no_wipe) {
    ORect r[9]; int n = 0;
    for (int i = 0; i < c->wipe_count; i++) r[n++] = R(&c->wipes[i]);
    if (c->layout == LAYOUT_DOUBLE && (c->middle_wipe[0] > 0 || c->middle_wipe[1] > 0)) r[n++] = (ORect){W / 2 - c->middle_wipe[0], 0, W / 2 + c->middle_wipe[1], H - 1};
    o_apply_wipes(&sheet, r, (size_t)n, mc);
  }
  if (!c->no_border) o_apply_border(&sheet, c->border, mc);
  if (!c->no_border_scan) {
    ORect bm[2];
    for (int i = 0; i < nout; i++) {
      Border b = o_detect_border(&sheet, &c->border_scan, outside[i]);
      bm[i] = o_border_to_mask(&sheet, b);
      res->borders[i] = b; res->border_masks[i] = U_(bm[i]);
    }
    res->border_count = nout;
    o_apply_masks(&sheet, bm, (size_t)nout, mc);
    if (!c->no_border_align) for (int i = 0; i < nout; i++) o_align_mask(&sheet, bm[i], outside[i], &c->mask_alignment);
  }
  if (!c->no_wipe) { ORect r[8]; for (int i = 0; i < c->post_wipe_count; i++) r[i] = R(&c->post_wipes[i]); o_apply_wipes(&sheet, r, (size_t)c->post_wipe_count, mc); }
  if (!c->no_border) o_apply_border(&sheet, c->post_border, mc);
  if (c->post_mirror.horizontal || c->post_mirror.vertical) o_mirror(&sheet, c->post_mirror.horizontal, c->post_mirror.vertical); /* :499-502 */
  if (c->post_shift.horizontal != 0 || c->post_shift.vertical != 0) {                                                             /* :504-508 */
    OImg n = o_shift(&sheet, c->post_shift.horizontal, c->post_shift.vertical); o_free(&sheet); sheet = n;
  }
  /* output in the page's format (saveImage(), file.c:211-262) */
  if (out) {
    OImg o = {out, W, H, row_bytes(fmt, W), fmt, {255, 255, 255}, c->abs_black_threshold};
    o_copy(&sheet, &o, full, 0, 0);
  }
  o_free(&sheet);
  return 0;
}

typedef struct { const B200SheetConfig *cfg; const uint8_t *pages; uint8_t *out; B200SheetResult *res; int pw, ph, fmt, n; size_t in_b, out_b; atomic_int next; } Job;
static void *worker(void *a) {
  Job *j = (Job *)a;
  for (;;) {
    int i = atomic_fetch_add(&j->next, 1);
    if (i >= j->n) break;
    B200SheetResult tmp;
    one_sheet(j->cfg, j->pages + j->in_b * (size_t)i, j->pw, j->ph, j->fmt, j->out ? j->out + j->out_b * (size_t)i : NULL, j->res ? &j->res[i] : &tmp);
  }
  return NULL;
}
int orc_process_sheets(const B200SheetConfig *cfg, const uint8_t *pages, int pw, int ph, int fmt, int n, uint8_t *out,
                       B200SheetResult *results, int threads, int *sw, int *sh) {
  if (fmt != B200_FMT_GRAY8 && fmt != B200_FMT_RGB24 && fmt != B200_FMT_MONOWHITE && fmt != B200_FMT_MONOBLACK) return -1;
  int W = pw * cfg->input_count;
  if (sw) *sw = W;
  if (sh) *sh = ph;
  Job j = {cfg, pages, out, results, pw, ph, fmt, n, (size_t)row_bytes(fmt, pw) * ph * cfg->input_count, (size_t)row_bytes(fmt, W) * ph};
  atomic_init(&j.next, 0);
  if (threads <= 1) worker(&j);
  else {
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(*th));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  return 0;
}
