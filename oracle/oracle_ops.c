/* oracle_ops.c — TEST INFRASTRUCTURE: CPU restatement of the reference's
 * per-sheet image operations (the `*_cpu` backend of ErrorTzy/unpaper-gpu).
 *
 * Plain sequential C, written from the reference's behaviour, each function
 * citing the reference file:line it restates.  It is the checker that travels
 * without /root/reference; it is PINNED in tests/test_oracle.py against
 *   (1) the unmodified reference compiled into oracle/_ref (bit-exact on every
 *       op, on random images and on synthetic pages), and
 *   (2) the reference's own golden image goldenA1.pbm for the default pipeline
 *       (through the vectors committed under tests/golden/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Entry points: `orc_host_<op>` with the signatures of `unpaper_b200_host_<op>`
 * (include/unpaper_b200.h), and orc_process_sheets() in oracle_sheet.c.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- pixel access (reference imageprocess/pixel.c:20-173) ------------------ */

int o_inside(const OImg *im, int x, int y) { return x >= 0 && y >= 0 && x < im->w && y < im->h; }

OPx o_get(const OImg *im, int x, int y) {
  OPx p = {255, 255, 255};                     /* outside = white, pixel.c:23-25 */
  if (!o_inside(im, x, y)) return p;
  const uint8_t *row = im->d + (size_t)y * im->ls;
  switch (im->fmt) {
  case B200_FMT_GRAY8: p.r = p.g = p.b = row[x]; break;
  case B200_FMT_Y400A: p.r = p.g = p.b = row[2 * x]; break;
  case B200_FMT_RGB24: p.r = row[3 * x]; p.g = row[3 * x + 1]; p.b = row[3 * x + 2]; break;
  case B200_FMT_MONOWHITE: p.r = p.g = p.b = (row[x / 8] & (128 >> (x % 8))) ? 0 : 255; break;
  case B200_FMT_MONOBLACK: p.r = p.g = p.b = (row[x / 8] & (128 >> (x % 8))) ? 255 : 0; break;
  }
  return p;
}
int o_gray(OPx p) { return (p.r + p.g + p.b) / 3; }                         /* pixel.c:16-18 */
int o_min(OPx p) { int m = p.r < p.g ? p.r : p.g; return m < p.b ? m : p.b; } /* lightness */
int o_max(OPx p) { int m = p.r > p.g ? p.r : p.g; return m > p.b ? m : p.b; } /* darkness_inverse */

void o_set(OImg *im, int x, int y, OPx p) {                                  /* pixel.c:136-173 */
  if (!o_inside(im, x, y)) return;
  uint8_t *row = im->d + (size_t)y * im->ls;
  int g = o_gray(p);
  switch (im->fmt) {
  case B200_FMT_GRAY8: row[x] = (uint8_t)g; break;
  case B200_FMT_Y400A: row[2 * x] = (uint8_t)g; row[2 * x + 1] = 0xFF; break;
  case B200_FMT_RGB24: row[3 * x] = (uint8_t)p.r; row[3 * x + 1] = (uint8_t)p.g; row[3 * x + 2] = (uint8_t)p.b; break;
  default: {
    int black = g < im->abt;
    if (im->fmt == B200_FMT_MONOWHITE) black = !black;
    if (!black) row[x / 8] |= (uint8_t)(128 >> (x % 8)); else row[x / 8] &= (uint8_t)~(128 >> (x % 8));
  }
  }
}

static const OPx WHITE_PX = {255, 255, 255};
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

ORect o_norm(ORect r) {                                                      /* primitives.c:46-61 */
  ORect n = {imin(r.x0, r.x1), imin(r.y0, r.y1), imax(r.x0, r.x1), imax(r.y0, r.y1)};
  return n;
}
ORect o_clip(const OImg *im, ORect r) {                                       /* image.c:72-88 */
  ORect n = o_norm(r);
  ORect c = {imax(n.x0, 0), imax(n.y0, 0), imin(n.x1, im->w - 1), imin(n.y1, im->h - 1)};
  return c;
}
int o_pt_in(int x, int y, ORect r) { ORect n = o_norm(r); return x >= n.x0 && x <= n.x1 && y >= n.y0 && y <= n.y1; }
/* count_pixels(size_of_rectangle()) with abs() (primitives.c:39-44,90-94) */
uint64_t o_count(ORect r) { return (uint64_t)((abs(r.x0 - r.x1) + 1) * (abs(r.y0 - r.y1) + 1)); }

OImg o_new(int w, int h, int fmt, const OImg *like) {
  OImg im; memset(&im, 0, sizeof(im));
  int row = fmt == B200_FMT_GRAY8 ? w : fmt == B200_FMT_Y400A ? 2 * w : fmt == B200_FMT_RGB24 ? 3 * w : (w + 7) / 8;
  im.w = w; im.h = h; im.fmt = fmt; im.ls = (row + 7) / 8 * 8;
  im.d = (uint8_t *)calloc((size_t)im.ls * (size_t)(h > 0 ? h : 1) + 16, 1);
  if (like) { memcpy(im.bg, like->bg, 3); im.abt = like->abt; } else { im.bg[0] = im.bg[1] = im.bg[2] = 255; }
  return im;
}
void o_free(OImg *im) { free(im->d); im->d = NULL; }

/* ---- blit (imageprocess/blit.c) -------------------------------------------- */

void o_wipe(OImg *im, ORect area, OPx c) {                                   /* blit.c:20-24 */
  ORect a = o_clip(im, area);
  for (int y = a.y0; y <= a.y1; y++) for (int x = a.x0; x <= a.x1; x++) o_set(im, x, y, c);
}

void o_copy(const OImg *s, OImg *t, ORect area, int tx, int ty) {            /* blit.c:30-80 */
  ORect a = o_clip(s, area);
  int w = a.x1 - a.x0 + 1, h = a.y1 - a.y0 + 1;
  int bpp = s->fmt == B200_FMT_GRAY8 ? 1 : s->fmt == B200_FMT_Y400A ? 2 : s->fmt == B200_FMT_RGB24 ? 3 : 0;
  if (s->fmt == t->fmt && bpp && w > 0 && h > 0 && tx >= 0 && ty >= 0 && tx + w <= t->w && ty + h <= t->h) {
    for (int y = 0; y < h; y++)     /* same-format fast path copies bytes verbatim (alpha included) */
      memcpy(t->d + (size_t)(ty + y) * t->ls + (size_t)tx * bpp, s->d + (size_t)(a.y0 + y) * s->ls + (size_t)a.x0 * bpp, (size_t)w * bpp);
    return;
  }
  for (int y = a.y0; y <= a.y1; y++) for (int x = a.x0; x <= a.x1; x++) o_set(t, tx + (x - a.x0), ty + (y - a.y0), o_get(s, x, y));
}

void o_center_image(const OImg *s, OImg *t, int ox, int oy, int tw, int th) { /* blit.c:175-202 */
  int sx = 0, sy = 0, sw = s->w, sh = s->h;
  if (sw < tw || sh < th) { ORect r = {ox, oy, ox + tw - 1, oy + th - 1}; OPx bg = {t->bg[0], t->bg[1], t->bg[2]}; o_wipe(t, r, bg); }
  if (sw <= tw) ox += (tw - sw) / 2; else { sx += (sw - tw) / 2; sw = tw; }
  if (sh <= th) oy += (th - sh) / 2; else { sy += (sh - th) / 2; sh = th; }
  ORect a = {sx, sy, sx + sw - 1, sy + sh - 1};
  o_copy(s, t, a, ox, oy);
}

/* ---- interpolation (imageprocess/interpolate.c:13-129) ---------------------- */

static uint8_t clip8(int a) { return a < 0 ? 0 : a > 255 ? 255 : (uint8_t)a; }
static uint8_t cubic1(float f, uint8_t a, uint8_t b, uint8_t c, uint8_t d) {  /* interpolate.c:24-32 */
  int result = b + 0.5f * f * (c - a + f * (2.0f * a - 5.0f * b + 4.0f * c - d + f * (3.0f * (b - c) + d - a)));
  return clip8(result);
}
static OPx cubic_px(float f, OPx q[4]) {
  OPx o = {cubic1(f, q[0].r, q[1].r, q[2].r, q[3].r), cubic1(f, q[0].g, q[1].g, q[2].g, q[3].g), cubic1(f, q[0].b, q[1].b, q[2].b, q[3].b)};
  return o;
}
static uint8_t lin1(float x, uint8_t a, uint8_t b) { return (1.0f - x) * a + x * b; }
static OPx lin_px(float f, OPx a, OPx b) { OPx o = {lin1(f, a.r, b.r), lin1(f, a.g, b.g), lin1(f, a.b, b.b)}; return o; }

OPx o_interp(const OImg *im, float fx, float fy, int type) {
  if (type == INTERP_NN) return o_get(im, (int)roundf(fx), (int)roundf(fy));
  if (type == INTERP_LINEAR) {                                                /* interpolate.c:76-117, quirks kept */
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy), x2 = (int)ceil(fx), y2 = (int)ceilf(fy);
    if (!o_inside(im, x2, y2)) return o_get(im, x1, y1);
    if (x1 == x2 && y1 == y2) return o_get(im, x1, y1);
    if (x1 == x2) return lin_px(fx - x1, o_get(im, x1, y1), o_get(im, x2, y2));
    if (y1 == y2) return lin_px(fy - y1, o_get(im, x1, y1), o_get(im, x2, y2));
    OPx h1 = lin_px(fx - x1, o_get(im, x1, y1), o_get(im, x2, y1));
    OPx h2 = lin_px(fx - x1, o_get(im, x1, y2), o_get(im, x2, y2));
    return lin_px(fy - y1, h1, h2);
  }
  int px = (int)fx, py = (int)fy;                                             /* truncation toward zero, interpolate.c:45 */
  OPx rows[4];
  for (int i = -1; i < 3; i++) {
    OPx q[4] = {o_get(im, px - 1, py + i), o_get(im, px, py + i), o_get(im, px + 1, py + i), o_get(im, px + 2, py + i)};
    rows[i + 1] = cubic_px(fx - px, q);
  }
  return cubic_px(fy - py, rows);
}

OImg o_stretch(const OImg *s, int w, int h, int type) {                        /* blit.c:209-239 */
  OImg t = o_new(w, h, s->fmt, s);
  const float hr = (float)s->w / (float)w, vr = (float)s->h / (float)h;
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) o_set(&t, x, y, o_interp(s, x * hr, y * vr, type));
  return t;
}

OImg o_resize(const OImg *s, int w, int h, int type) {                         /* blit.c:246-284 */
  const float hr = (float)w / (float)s->w, vr = (float)h / (float)s->h;
  int sw, sh;
  if (hr < vr) { sw = w; sh = s->h * hr; } else if (vr < hr) { sw = s->w * vr; sh = h; } else { sw = w; sh = h; }
  OImg st;
  if (sw == s->w && sh == s->h) { st = o_new(s->w, s->h, s->fmt, s); memcpy(st.d, s->d, (size_t)s->ls * s->h); }
  else st = o_stretch(s, sw, sh, type);
  if (w == sw && h == sh) return st;
  OImg r = o_new(w, h, s->fmt, s);
  ORect full = {0, 0, w - 1, h - 1}; OPx bg = {s->bg[0], s->bg[1], s->bg[2]};
  o_wipe(&r, full, bg);
  o_center_image(&st, &r, 0, 0, w, h);
  o_free(&st);
  return r;
}

OImg o_rotate90(const OImg *s, int dir) {                                      /* blit.c:291-314 */
  OImg t = o_new(s->h, s->w, s->fmt, s);
  for (int y = 0; y < s->h; y++) {
    int xx = ((dir > 0) ? s->h - 1 : 0) - y * dir;
    for (int x = 0; x < s->w; x++) { int yy = ((dir < 0) ? s->w - 1 : 0) + x * dir; o_set(&t, xx, yy, o_get(s, x, y)); }
  }
  return t;
}

void o_mirror(OImg *im, int dh, int dv) {                                       /* blit.c:320-354 */
  int xmax = im->w - 1, ymax = im->h - 1;
  if (dh && !dv) xmax = (im->w - 1) / 2;
  if (dv) ymax = (im->h - 1) / 2;
  for (int y = 0; y <= ymax; y++) {
    int yy = dv ? im->h - y - 1 : y;
    if (dv && dh && y == yy) xmax = (im->w - 1) / 2;
    for (int x = 0; x <= xmax; x++) {
      int xx = dh ? im->w - x - 1 : x;
      OPx a = o_get(im, x, y), b = o_get(im, xx, yy);
      o_set(im, x, y, b); o_set(im, xx, yy, a);
    }
  }
}

OImg o_shift(const OImg *s, int dx, int dy) {                                    /* blit.c:360-368 */
  OImg t = o_new(s->w, s->h, s->fmt, s);
  ORect full = {0, 0, s->w - 1, s->h - 1}; OPx bg = {s->bg[0], s->bg[1], s->bg[2]};
  o_wipe(&t, full, bg);
  o_copy(s, &t, full, dx, dy);
  return t;
}

/* ---- rectangle statistics (blit.c:91-167) ----------------------------------- */

static uint8_t rect_avg(const OImg *im, ORect in, int which) {
  ORect a = o_clip(im, in);
  uint64_t n = o_count(a), s = 0;        /* abs()-based: an inverted clip still has n > 0 and scans nothing -> 255 */
  if (n == 0) return 0;
  for (int y = a.y0; y <= a.y1; y++) for (int x = a.x0; x <= a.x1; x++) {
    OPx p = o_get(im, x, y);
    s += which == 0 ? o_gray(p) : which == 1 ? o_min(p) : o_max(p);
  }
  return (uint8_t)(0xFF - (s / n));
}
uint8_t o_inverse_brightness(const OImg *im, ORect r) { return rect_avg(im, r, 0); }
uint8_t o_inverse_lightness(const OImg *im, ORect r) { return rect_avg(im, r, 1); }
uint8_t o_darkness(const OImg *im, ORect r) { return rect_avg(im, r, 2); }
uint64_t o_count_brightness(const OImg *im, ORect a, int lo, int hi) {          /* blit.c:148-167: no clipping */
  uint64_t c = 0;
  for (int y = a.y0; y <= a.y1; y++) for (int x = a.x0; x <= a.x1; x++) { int g = o_gray(o_get(im, x, y)); if (g >= lo && g <= hi) c++; }
  return c;
}

/* ---- masks (imageprocess/masks.c) ------------------------------------------- */

void o_apply_masks(OImg *im, const ORect *m, size_t n, OPx c) {                  /* masks.c:311-325 */
  if (n == 0) return;
  for (int y = 0; y < im->h; y++) for (int x = 0; x < im->w; x++) {
    int in = 0;
    for (size_t k = 0; k < n && !in; k++) in = o_pt_in(x, y, m[k]);
    if (!in) o_set(im, x, y, c);
  }
}
void o_apply_wipes(OImg *im, const ORect *w, size_t n, OPx c) {                   /* masks.c:337-345 */
  for (size_t k = 0; k < n; k++)
    for (int y = w[k].y0; y <= w[k].y1; y++) for (int x = w[k].x0; x <= w[k].x1; x++) o_set(im, x, y, c);
}
ORect o_border_to_mask(const OImg *im, Border b) { ORect r = {b.left, b.top, im->w - b.right - 1, im->h - b.bottom - 1}; return r; }
void o_apply_border(OImg *im, Border b, OPx c) {                                  /* masks.c:370-382 */
  if (!b.left && !b.top && !b.right && !b.bottom) return;
  ORect m = o_border_to_mask(im, b);
  o_apply_masks(im, &m, 1, c);
}

static uint32_t detect_edge(const OImg *im, int ox, int oy, int sx, int sy, int size, int depth, float thr) { /* masks.c:54-100 */
  ORect a;
  if (sy == 0) { if (depth == -1) depth = im->h; a.x0 = ox + (-size / 2); a.y0 = oy + (-depth / 2); a.x1 = a.x0 + size - 1; a.y1 = a.y0 + depth - 1; }
  else { if (depth == -1) depth = im->w; a.x0 = ox + (-depth / 2); a.y0 = oy + (-size / 2); a.x1 = a.x0 + depth - 1; a.y1 = a.y0 + size - 1; }
  uint32_t total = 0, count = 0;
  uint8_t blackness;
  do {
    blackness = o_inverse_brightness(im, a);
    total += blackness; count++;
    a.x0 += sx; a.x1 += sx; a.y0 += sy; a.y1 += sy;
  } while ((blackness >= ((thr * total) / count)) && blackness != 0);
  return count;
}

size_t o_detect_masks(const OImg *im, const MaskDetectionParameters *p, const Point *pts, size_t n, ORect *out) { /* masks.c:107-205 */
  size_t cnt = 0;
  if (!p->scan_direction.horizontal && !p->scan_direction.vertical) return 0;
  for (size_t i = 0; i < n; i++) {
    ORect m; int ox = pts[i].x, oy = pts[i].y;
    if (p->scan_direction.horizontal) {
      int l = (int)detect_edge(im, ox, oy, -p->scan_step.horizontal, 0, p->scan_size.width, p->scan_depth.horizontal, p->scan_threshold.horizontal);
      int r = (int)detect_edge(im, ox, oy, p->scan_step.horizontal, 0, p->scan_size.width, p->scan_depth.horizontal, p->scan_threshold.horizontal);
      m.x0 = ox - p->scan_step.horizontal * l - p->scan_size.width / 2; m.x1 = ox + p->scan_step.horizontal * r + p->scan_size.width / 2;
    } else { m.x0 = 0; m.x1 = im->w - 1; }
    if (p->scan_direction.vertical) {
      int t = (int)detect_edge(im, ox, oy, 0, -p->scan_step.vertical, p->scan_size.height, p->scan_depth.vertical, p->scan_threshold.vertical);
      int b = (int)detect_edge(im, ox, oy, 0, p->scan_step.vertical, p->scan_size.height, p->scan_depth.vertical, p->scan_threshold.vertical);
      m.y0 = oy - p->scan_step.vertical * t - p->scan_size.height / 2; m.y1 = oy + p->scan_step.vertical * b + p->scan_size.height / 2;
    } else { m.y0 = 0; m.y1 = im->h - 1; }
    int w = abs(m.x0 - m.x1) + 1, h = abs(m.y0 - m.y1) + 1;
    if ((p->minimum_width != -1 && w < p->minimum_width) || (p->maximum_width != -1 && w > p->maximum_width)) { m.x0 = ox - p->maximum_width / 2; m.x1 = ox + p->maximum_width / 2; }
    if ((p->minimum_height != -1 && h < p->minimum_height) || (p->maximum_height != -1 && h > p->maximum_height)) { m.y0 = oy - p->maximum_height / 2; m.y1 = oy + p->maximum_height / 2; }
    out[i] = m;
    if (!(m.x0 == -1 && m.y0 == -1 && m.x1 == -1 && m.y1 == -1)) cnt++;
  }
  return cnt;
}

static void move_area(OImg *im, ORect area, int tx, int ty) {                     /* masks.c:239-243, :296-300 */
  int w = abs(area.x0 - area.x1) + 1, h = abs(area.y0 - area.y1) + 1;
  OImg tmp = o_new(w, h, im->fmt, im);
  ORect full = {0, 0, w - 1, h - 1}; OPx bg = {im->bg[0], im->bg[1], im->bg[2]};
  o_wipe(&tmp, full, bg);
  o_copy(im, &tmp, area, 0, 0);
  o_wipe(im, area, bg);
  o_copy(&tmp, im, full, tx, ty);
  o_free(&tmp);
}
int o_center_mask(OImg *im, int cx, int cy, ORect area) {                          /* masks.c:222-249 */
  int w = abs(area.x0 - area.x1) + 1, h = abs(area.y0 - area.y1) + 1;
  int tx = cx + (-w / 2), ty = cy + (-h / 2);
  ORect full = {0, 0, im->w - 1, im->h - 1};
  if (!(o_pt_in(tx, ty, full) && o_pt_in(tx + w - 1, ty + h - 1, full))) return 0;
  move_area(im, area, tx, ty);
  return 1;
}
void o_align_mask(OImg *im, ORect inside, ORect outside, const MaskAlignmentParameters *p) { /* masks.c:265-300 */
  int w = abs(inside.x0 - inside.x1) + 1, h = abs(inside.y0 - inside.y1) + 1, tx, ty;
  if (p->alignment.left) tx = outside.x0 + p->margin.horizontal; else if (p->alignment.right) tx = outside.x1 - w - p->margin.horizontal; else tx = (outside.x0 + outside.x1 - w) / 2;
  if (p->alignment.top) ty = outside.y0 + p->margin.vertical; else if (p->alignment.bottom) ty = outside.y1 - h - p->margin.vertical; else ty = (outside.y0 + outside.y1 - h) / 2;
  move_area(im, inside, tx, ty);
}

static uint32_t border_edge(const OImg *im, ORect om, int sx, int sy, int size, int32_t thr) { /* masks.c:410-448 */
  ORect a = om;
  int32_t max_step;
  if (sy == 0) { if (sx > 0) a.x1 = om.x0 + size; else a.x0 = om.x1 - size; max_step = abs(om.x0 - om.x1) + 1; }
  else { if (sy > 0) a.y1 = om.y0 + size; else a.y0 = om.y1 - size; max_step = abs(om.y0 - om.y1) + 1; }
  uint32_t result = 0;
  while (result < (uint32_t)max_step) {
    uint32_t cnt = (uint32_t)o_count_brightness(im, a, 0, im->abt);
    if (cnt >= (uint32_t)thr) return result;
    a.x0 += sx; a.x1 += sx; a.y0 += sy; a.y1 += sy;
    result += (uint32_t)abs(sx + sy);
  }
  return 0;
}
Border o_detect_border(const OImg *im, const BorderScanParameters *p, ORect om) {   /* masks.c:454-488 */
  Border b = {om.x0, om.y0, im->w - om.x1, im->h - om.y1};
  if (p->scan_direction.horizontal) {
    b.left += (int32_t)border_edge(im, om, p->scan_step.horizontal, 0, p->scan_size.width, p->scan_threshold.horizontal);
    b.right += (int32_t)border_edge(im, om, -p->scan_step.horizontal, 0, p->scan_size.width, p->scan_threshold.horizontal);
  }
  if (p->scan_direction.vertical) {
    b.top += (int32_t)border_edge(im, om, 0, p->scan_step.vertical, p->scan_size.height, p->scan_threshold.vertical);
    b.bottom += (int32_t)border_edge(im, om, 0, -p->scan_step.vertical, p->scan_size.height, p->scan_threshold.vertical);
  }
  return b;
}

/* ---- flood fill (imageprocess/fill.c:16-107), recursion made explicit --------- */

static uint64_t fill_line(OImg *im, int x, int y, int dx, int dy, int lo, int hi, uint64_t intensity) {
  uint64_t dist = 0, cnt = 1;
  for (;;) {
    x += dx; y += dy;
    int g = o_gray(o_get(im, x, y));
    if (g >= lo && g <= hi) cnt = intensity; else cnt--;
    if (cnt <= 0 || !o_inside(im, x, y)) return dist;
    o_set(im, x, y, WHITE_PX);
    dist++;
  }
}
typedef struct { int x, y; uint64_t len[4]; int line; uint64_t d; int sub; } Frame;
void o_flood_fill(OImg *im, int x, int y, int lo, int hi, uint64_t intensity) {
  static const int DX[4] = {-1, 0, 1, 0}, DY[4] = {0, -1, 0, 1};       /* left, up, right, down */
  size_t cap = 1024, sp = 0;
  Frame *st = (Frame *)malloc(cap * sizeof(Frame));
  int cx = x, cy = y, pending = 1;
  for (;;) {
    if (pending) {
      pending = 0;
      int g = o_gray(o_get(im, cx, cy));
      if (g >= lo && g <= hi) {
        if (sp == cap) { cap *= 2; st = (Frame *)realloc(st, cap * sizeof(Frame)); }
        Frame *f = &st[sp++];
        f->x = cx; f->y = cy; f->line = 0; f->d = 0; f->sub = 0;
        o_set(im, cx, cy, WHITE_PX);
        for (int k = 0; k < 4; k++) f->len[k] = fill_line(im, cx, cy, DX[k], DY[k], lo, hi, intensity);
      }
    }
    if (sp == 0) break;
    Frame *f = &st[sp - 1];
    /* next neighbour of the painted cross, in the order of fill.c:54-74 + :98-105 */
    while (f->line < 4 && f->d >= f->len[f->line]) { f->line++; f->d = 0; f->sub = 0; }
    if (f->line == 4) { sp--; continue; }
    int qx = f->x + DX[f->line] * (int)(f->d + 1), qy = f->y + DY[f->line] * (int)(f->d + 1);
    if (DX[f->line] != 0) { cx = qx; cy = qy + (f->sub == 0 ? 1 : -1); }   /* below, then above */
    else { cx = qx + (f->sub == 0 ? 1 : -1); cy = qy; }                    /* right, then left */
    if (++f->sub == 2) { f->sub = 0; f->d++; }
    pending = 1;
  }
  free(st);
}

/* ---- filters (imageprocess/filters.c) ------------------------------------------ */

static int excluded(ORect a, const BlackfilterParameters *p) {                      /* primitives.c:117-133 */
  for (size_t i = 0; i < p->exclusions_count; i++) {
    ORect e = {p->exclusions[i].vertex[0].x, p->exclusions[i].vertex[0].y, p->exclusions[i].vertex[1].x, p->exclusions[i].vertex[1].y};
    ORect n = o_norm(a);
    if (o_pt_in(n.x0, n.y0, e) || o_pt_in(n.x1, n.y1, e)) return 1;
  }
  return 0;
}
static int bf_scan(OImg *im, const BlackfilterParameters *p, int sx, int sy, int sw, int sh, int shx, int shy) { /* filters.c:49-104 */
  int fills = 0;
  ORect a = {0, 0, sw - 1, sh - 1};
  while (o_inside(im, a.x0, a.y0)) {
    if (!o_inside(im, a.x1, a.y1)) { int dx = im->w - 1 - a.x1, dy = im->h - 1 - a.y1; a.x0 += dx; a.x1 += dx; a.y0 += dy; a.y1 += dy; }
    do {
      if (o_darkness(im, a) >= p->abs_threshold && !excluded(a, p)) {
        fills++;
        for (int y = a.y0; y <= a.y1; y++) for (int x = a.x0; x <= a.x1; x++) o_flood_fill(im, x, y, 0, im->abt, (uint64_t)p->intensity);
      }
      a.x0 += sx; a.x1 += sx; a.y0 += sy; a.y1 += sy;
    } while (o_inside(im, a.x0, a.y0));
    a.x0 += shx; a.x1 += shx; a.y0 += shy; a.y1 += shy;
  }
  return fills;
}
int o_blackfilter(OImg *im, const BlackfilterParameters *p) {                         /* filters.c:111-127 */
  int fills = 0;
  if (p->scan_direction.horizontal) fills += bf_scan(im, p, p->scan_step.horizontal, 0, p->scan_size.width, (int)p->scan_depth.vertical, 0, (int)p->scan_depth.vertical);
  if (p->scan_direction.vertical) fills += bf_scan(im, p, 0, p->scan_step.vertical, (int)p->scan_depth.horizontal, p->scan_size.height, (int)p->scan_depth.horizontal, 0);
  return fills;
}

void o_blurfilter(OImg *im, const BlurfilterParameters *p, int white) {               /* filters.c:149-232 */
  int bw = p->scan_size.width, bh = p->scan_size.height;
  uint32_t n = (uint32_t)(im->w / bw);
  uint64_t T = (uint64_t)(bw * bh);
  /* the reference's three count rows alias one another (prev=&buf[0][0], cur=&buf[0][1],
   * next=&buf[0][2], filters.c:160-167) inside a never-initialised VLA; zero-initialised here,
   * matching how oracle/_ref is compiled (-ftrivial-auto-var-init=zero) */
  uint64_t *flat = (uint64_t *)calloc(3 * ((size_t)n + 2) + 4, sizeof(uint64_t));
  uint64_t *prev = flat, *cur = flat + 1, *next = flat + 2;
  cur[0] = T; cur[n] = T; next[0] = T; next[n] = T;
  int max_left = im->w - bw, max_top = im->h - bh;
  for (int left = 0, b = 1; left <= max_left; left += bw) { ORect r = {left, 0, left + bw - 1, bh - 1}; cur[b++] = o_count_brightness(im, r, 0, white); }
  for (int top = 0; top <= max_top; top += bh) {
    ORect r0 = {0, top + p->scan_step.vertical, bw - 1, top + p->scan_step.vertical + bh - 1};
    next[0] = o_count_brightness(im, r0, 0, white);
    for (int left = 0, b = 1; left <= max_left; left += bw, b++) {
      ORect rn = {left + bw, top + p->scan_step.vertical, left + 2 * bw - 1, top + p->scan_step.vertical + bh - 1};
      next[b + 1] = o_count_brightness(im, rn, 0, white);
      uint64_t m = prev[b - 1]; if (prev[b + 1] > m) m = prev[b + 1]; if (cur[b] > m) m = cur[b];
      if (next[b - 1] > m) m = next[b - 1]; if (next[b + 1] > m) m = next[b + 1];
      if ((((float)m) / T) <= p->intensity) { ORect w = {left, top, left + bw - 1, top + bh - 1}; o_wipe(im, w, WHITE_PX); cur[b] = T; }
    }
    uint64_t *t = prev; prev = cur; cur = next; next = t;
  }
  free(flat);
}

static int nf_cc(OImg *im, int x, int y, int clear, int white) {                        /* filters.c:243-254 */
  if (o_min(o_get(im, x, y)) >= white) return 0;
  if (clear) o_set(im, x, y, WHITE_PX);
  return 1;
}
static uint64_t nf_level(OImg *im, int px, int py, uint32_t level, int clear, int white) {  /* filters.c:256-285 */
  uint64_t c = 0;
  /* the reference compares a signed loop variable with an unsigned bound: a ring whose
   * first coordinate would be negative is skipped entirely */
  for (int32_t xx = px - level; (uint32_t)xx <= px + level; xx++) { c += nf_cc(im, xx, py - (int)level, clear, white); c += nf_cc(im, xx, py + (int)level, clear, white); }
  for (int32_t yy = py - (level - 1); (uint32_t)yy <= py + (level - 1); yy++) { c += nf_cc(im, px - (int)level, yy, clear, white); c += nf_cc(im, px + (int)level, yy, clear, white); }
  return c;
}
uint64_t o_noisefilter(OImg *im, uint64_t intensity, int white) {                        /* filters.c:324-348 */
  uint64_t clusters = 0;
  for (int y = 0; y < im->h; y++) for (int x = 0; x < im->w; x++) {
    if (o_max(o_get(im, x, y)) >= white) continue;
    uint64_t count = 1, l; uint32_t level = 1;
    do { l = nf_level(im, x, y, level, 0, white); count += l; level++; } while (l != 0 && level <= intensity);
    if (count <= intensity) {
      o_set(im, x, y, WHITE_PX);
      level = 1;
      do { l = nf_level(im, x, y, level, 1, white); level++; } while (l != 0);
      clusters++;
    }
  }
  return clusters;
}

void o_grayfilter(OImg *im, const GrayfilterParameters *p) {                             /* filters.c:370-402 */
  int x = 0, y = 0;
  do {
    ORect a = {x, y, x + p->scan_size.width - 1, y + p->scan_size.height - 1};
    if (o_count_brightness(im, a, 0, im->abt) == 0) {
      if (o_inverse_lightness(im, a) < p->abs_threshold) o_wipe(im, a, WHITE_PX);
    }
    if (x < im->w) x += p->scan_step.horizontal; else { x = 0; y += p->scan_step.vertical; }
  } while (y <= im->h);
}

/* ---- deskew (imageprocess/deskew.c) ------------------------------------------------ */

static int rot_peak(const OImg *im, ORect mask, const DeskewParameters *p, int shx, int shy, float m) { /* deskew.c:48-142 */
  int sw = abs(mask.x0 - mask.x1) + 1, sh = abs(mask.y0 - mask.y1) + 1;
  int scan = p->deskewScanSize, maxDepth;
  float X, Y, stepX, stepY;
  int maxAbs = 255 * p->deskewScanSize * p->deskewScanDepth;
  if (shy == 0) {
    if (scan == -1) scan = sh;
    scan = imin(scan, imin(10000, sh));
    maxDepth = sw / 2;
    int half = scan / 2, outer = (int)(fabsf(m) * half), mid = sh / 2;
    int side = shx > 0 ? mask.x0 - outer : mask.x1 + outer;
    X = side + half * m; Y = mask.y0 + mid - half; stepX = -m; stepY = 1.0;
  } else {
    if (scan == -1) scan = sw;
    scan = imin(scan, imin(10000, sw));
    maxDepth = sh / 2;
    int half = scan / 2, outer = (int)(fabsf(m) * half), mid = sw / 2;
    int side = shy > 0 ? mask.x0 - outer : mask.x1 + outer;      /* deskew.c:96-97: .x where .y is meant */
    X = mask.x0 + mid - half; Y = side - (half * m); stepX = 1.0; stepY = -m;
  }
  if (scan <= 0) return 0;
  int *px = (int *)malloc(sizeof(int) * (size_t)scan), *py = (int *)malloc(sizeof(int) * (size_t)scan);
  for (int k = 0; k < scan; k++) { px[k] = (int)X; py[k] = (int)Y; X += stepX; Y += stepY; }
  int dep, last = 0, maxDiff = 0, acc = 0;
  for (dep = 0; acc < maxAbs && dep < maxDepth; dep++) {
    int blackness = 0;
    for (int k = 0; k < scan; k++) {
      int x = px[k], y = py[k];
      px[k] += shx; py[k] += shy;
      if (o_pt_in(x, y, mask)) blackness += 255 - o_max(o_get(im, x, y));
    }
    int diff = blackness - last;
    last = blackness;
    if (diff >= maxDiff) maxDiff = diff;
    acc += blackness;
  }
  free(px); free(py);
  return dep < maxDepth ? maxDiff : 0;
}
static float edge_rotation(const OImg *im, ORect mask, const DeskewParameters *p, int shx, int shy) {  /* deskew.c:148-171 */
  int max_peak = 0; float detected = 0.0;
  for (float rotation = 0.0; rotation <= p->deskewScanRangeRad; rotation = (rotation >= 0.0) ? -(rotation + p->deskewScanStepRad) : -rotation) {
    int peak = rot_peak(im, mask, p, shx, shy, tanf(rotation));
    if (peak > max_peak) { detected = rotation; max_peak = peak; }
  }
  return detected;
}
float o_detect_rotation(const OImg *im, ORect mask, const DeskewParameters *p) {          /* deskew.c:178-241 */
  float r[4]; int n = 0;
  if (p->scan_edges.left) r[n++] = edge_rotation(im, mask, p, 1, 0);
  if (p->scan_edges.top) r[n++] = -edge_rotation(im, mask, p, 0, 1);
  if (p->scan_edges.right) r[n++] = edge_rotation(im, mask, p, -1, 0);
  if (p->scan_edges.bottom) r[n++] = -edge_rotation(im, mask, p, 0, -1);
  float total = 0.0;
  for (int i = 0; i < n; i++) total += r[i];
  float average = total / n;
  total = 0.0;
  for (int i = 0; i < n; i++) total += powf(r[i] - average, 2);
  float deviation = sqrtf(total);
  return deviation <= p->deskewScanDeviationRad ? average : 0.0f;
}
void o_deskew(OImg *im, ORect mask, float radians, int type) {                           /* deskew.c:253-286 */
  int w = abs(mask.x0 - mask.x1) + 1, h = abs(mask.y0 - mask.y1) + 1;
  OImg rot = o_new(w, h, im->fmt, im);
  ORect nm = o_norm(mask);
  float scx = nm.x0 + w / 2.0f, scy = nm.y0 + h / 2.0f, tcx = 0 + w / 2.0f, tcy = 0 + h / 2.0f;
  const float sinval = sinf(-radians), cosval = cosf(-radians);
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    const float srcX = scx + (x - tcx) * cosval + (y - tcy) * sinval;
    const float srcY = scy + (y - tcy) * cosval - (x - tcx) * sinval;
    o_set(&rot, x, y, o_interp(im, srcX, srcY, type));
  }
  ORect full = {0, 0, w - 1, h - 1};
  o_copy(&rot, im, full, mask.x0, mask.y0);
  o_free(&rot);
}
