/* oracle.h — TEST INFRASTRUCTURE: declarations of the CPU restatement
 * (see oracle_ops.c for what it is pinned against). */
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "unpaper_b200.h"

typedef struct { uint8_t *d; int w, h, ls, fmt; uint8_t bg[3]; uint8_t abt; } OImg;
typedef struct { int r, g, b; } OPx;
typedef struct { int x0, y0, x1, y1; } ORect;

int o_inside(const OImg *im, int x, int y);
OPx o_get(const OImg *im, int x, int y);
void o_set(OImg *im, int x, int y, OPx p);
int o_gray(OPx p); int o_min(OPx p); int o_max(OPx p);
ORect o_norm(ORect r); ORect o_clip(const OImg *im, ORect r); int o_pt_in(int x, int y, ORect r);
uint64_t o_count(ORect r);
OImg o_new(int w, int h, int fmt, const OImg *like);
void o_free(OImg *im);

void o_wipe(OImg *im, ORect area, OPx c);
void o_copy(const OImg *s, OImg *t, ORect area, int tx, int ty);
void o_center_image(const OImg *s, OImg *t, int ox, int oy, int tw, int th);
OPx o_interp(const OImg *im, float fx, float fy, int type);
OImg o_stretch(const OImg *s, int w, int h, int type);
OImg o_resize(const OImg *s, int w, int h, int type);
OImg o_rotate90(const OImg *s, int dir);
void o_mirror(OImg *im, int dh, int dv);
OImg o_shift(const OImg *s, int dx, int dy);
uint8_t o_inverse_brightness(const OImg *im, ORect r);
uint8_t o_inverse_lightness(const OImg *im, ORect r);
uint8_t o_darkness(const OImg *im, ORect r);
uint64_t o_count_brightness(const OImg *im, ORect a, int lo, int hi);
void o_apply_masks(OImg *im, const ORect *m, size_t n, OPx c);
void o_apply_wipes(OImg *im, const ORect *w, size_t n, OPx c);
ORect o_border_to_mask(const OImg *im, Border b);
void o_apply_border(OImg *im, Border b, OPx c);
size_t o_detect_masks(const OImg *im, const MaskDetectionParameters *p, const Point *pts, size_t n, ORect *out);
int o_center_mask(OImg *im, int cx, int cy, ORect area);
void o_align_mask(OImg *im, ORect inside, ORect outside, const MaskAlignmentParameters *p);
Border o_detect_border(const OImg *im, const BorderScanParameters *p, ORect om);
void o_flood_fill(OImg *im, int x, int y, int lo, int hi, uint64_t intensity);
int o_blackfilter(OImg *im, const BlackfilterParameters *p);
void o_blurfilter(OImg *im, const BlurfilterParameters *p, int white);
uint64_t o_noisefilter(OImg *im, uint64_t intensity, int white);
void o_grayfilter(OImg *im, const GrayfilterParameters *p);
float o_detect_rotation(const OImg *im, ORect mask, const DeskewParameters *p);
void o_deskew(OImg *im, ORect mask, float radians, int type);
