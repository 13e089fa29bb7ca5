/* dev.h — descriptors shared by the C host layer and the CUDA kernels.
 *
 * Design: every kernel works on a GROUP of independent pages (gridDim.z or
 * gridDim.y = page).  Whatever a stage decides (masks, rotation, border, fill
 * candidates, wipe flags) stays in the page's DPage record in device memory
 * and is consumed from there by the next kernel, so a whole sheet runs without
 * a host round-trip.  The reference-facing vtable is the same kernels with a
 * group of one and a readback after each op.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* internal pixel-format codes */
#define D_INK_CELL 8   /* cell size of the ink map used by rotate() to skip white tiles */
enum { DF_GRAY8 = 0, DF_Y400A = 1, DF_RGB24 = 2, DF_MONOWHITE = 3, DF_MONOBLACK = 4 };

typedef struct { int32_t x0, y0, x1, y1; } DRect;      /* inclusive, like Rectangle */
typedef struct { int32_t left, top, right, bottom; } DBorder;

typedef struct {
  uint8_t *data;
  int32_t w, h, pitch, fmt;
  uint8_t abt;          /* Image.abs_black_threshold (set_pixel on mono formats) */
  uint8_t bg[3];        /* Image.background */
} DImg;

#define D_MAX_MASKS 8
#define D_MAX_BORDERS 2
#define D_MAX_RECTS 100

/* statistic selectors for the line-sum kernels */
enum { ST_GRAY = 0, ST_MAXCH = 1, ST_COUNT_GRAY_RANGE = 2 };

/* a rectangle move: copy src_area of img -> aux(0,0); wipe src_area with bg;
 * copy aux -> img at (tx,ty).  (center_mask masks.c:222-249, align_mask
 * masks.c:265-305) */
typedef struct {
  DRect area; int32_t tx, ty; int32_t enabled; int32_t use_masks;
  /* sorted byte positions along a row where the source of a pixel can change (k_move_pass) */
  int32_t nseg, pad;
  int32_t bnd[16];
} DMove;

/* one line-sum job: sums `stat` along a band.  axis 0: out[x-xa] = sum over
 * y in [ya,yb] (column sums); axis 1: out[y-ya] = sum over x in [xa,xb]. */
typedef struct {
  int32_t xa, xb, ya, yb;   /* clipped, inclusive; empty if xa>xb or ya>yb */
  int32_t axis;
  int32_t out_off;          /* offset (u32 elements) into DPage.u32 */
} DLineJob;

typedef struct DPage {
  DImg img;              /* working image */
  DImg aux;              /* scratch image, same format, capacity >= img (vtable ops; unused by the sheet engine) */
  /* sheet engine: every sheet slot has two buffers of equal geometry.  Passes that move
   * pixels (deskew, centre / align a mask, shift) render img -> other in one sweep and
   * then swap the two pointers; buf[] are the fixed addresses (img.data = buf[0] at the
   * start of every sheet). */
  uint8_t *other;
  uint8_t *buf[2];
  uint8_t *cls;          /* [h*w] noisefilter classes */
  uint32_t *list;        /* noisefilter mutable list */
  uint32_t *u32;         /* general u32 scratch */
  uint64_t *stack;       /* flood-fill frame stack (4 x u64 per frame) */
  uint8_t *ink;          /* D_INK_CELL x D_INK_CELL pixel cells: 1 = every pixel of the cell is pure white */
  int32_t ink_ncx, ink_ncy, ink_ok, ink_cap;
  uint32_t *pre;         /* column prefix sums for the rotation scan: [rows+1][img.w] */
  int64_t pre_cap;       /* capacity of `pre` in u32 elements */
  int32_t list_cap, u32_cap, stack_cap, pad0;

  /* counters (zeroed per sheet) */
  uint32_t list_n;
  uint32_t nf_clusters;
  uint32_t bf_fills;
  uint32_t error;        /* sticky error bits, see DERR_* */

  /* detector inputs that can come from the host (vtable) or a previous stage */
  int32_t point_count;
  int32_t px[D_MAX_MASKS], py[D_MAX_MASKS];
  DRect outside[D_MAX_BORDERS];
  int32_t outside_count;

  /* detector results */
  int32_t edge_count[D_MAX_MASKS][4];   /* detect_edge steps: left,right,top,bottom */
  int32_t mask_count;
  int32_t mask_valid[D_MAX_MASKS];
  DRect masks[D_MAX_MASKS];
  int32_t mask_count_deskew;             /* snapshot taken by the rotation stage */
  DRect masks_deskew[D_MAX_MASKS];
  int32_t rot_angle_idx[D_MAX_MASKS][4]; /* winning angle index per edge */
  float rotation[D_MAX_MASKS];
  float rot_sin[D_MAX_MASKS], rot_cos[D_MAX_MASKS];  /* of -rotation */
  int32_t rot_apply[D_MAX_MASKS];
  int32_t rot_more[D_MAX_MASKS];         /* a scan line of this mask did not end within the depth-limited prefix table */
  int32_t centered[D_MAX_MASKS];
  DBorder border[D_MAX_BORDERS];
  DRect border_mask[D_MAX_BORDERS];
  DMove move;
} DPage;

/* blit job descriptors (device memory; written by the host for vtable calls or
 * by a prep kernel inside the sheet engine) */
typedef struct { DImg img; DRect r; uint8_t c[3]; uint8_t pad; int32_t enabled; } DFillJob;
typedef struct { DImg src, dst; DRect area; int32_t tx, ty; int32_t enabled; int32_t pad; } DCopyJob;
typedef struct { DImg img; const DRect *rects; int32_t nrects; uint8_t c[3]; uint8_t pad; int32_t enabled; int32_t pad2; } DMaskJob;

/* one blackfilter scan position (filters.c:60-103), in scan order; its sum on
 * the untouched image = sum of `axis`-line sums at u32[sum_off + coord]. */
typedef struct { DRect r; int32_t sum_off; int32_t axis; } DBfPos;

enum {
  DERR_LIST_OVERFLOW = 1,    /* noisefilter mutable list too small */
  DERR_STACK_OVERFLOW = 2,   /* flood-fill frame stack too small */
  DERR_EDGE_RUNAWAY = 4,     /* detect_edge left the image without stopping */
  DERR_UNSUPPORTED = 8,
};

#ifdef __cplusplus
}
#endif
