/* host.h — internal declarations of the C host layer. */
#pragma once
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "dev.h"
#include "launch.h"
#include "rt.h"
#include "unpaper_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int b200_fmt_to_dev(int av_pix_fmt);            /* -1 if unsupported */
int b200_fmt_row_bytes(int av_pix_fmt, int width);
bool b200_image_view(Image *image, DImg *out);  /* image_res.c */

/* ---- plans: static, geometry-dependent tables in device memory ----------- */

typedef struct {
  int npos, njobs, sums_len, flag_off, u32_need;
  DBfPos *pos_dev;
  DLineJob *jobs_host, *jobs_dev;
  int abs_threshold, mask_hi;
  long long intensity;
} BfPlan;

typedef struct {
  int n, nrows, bw, bh, nrects;
  DRect *rects_dev;
  int cnt_off, state_off, flag_off, u32_need;
  unsigned long long T;
  float intensity;
  int white;
} BlurPlan;

typedef struct {
  int gp[18];
  int white_off, u32_need;
  int dark_max;
  int ok;
} GrayPlan;

typedef struct {
  int njobs, stride, u32_need, max_points;
  DLineJob *jobs_host, *jobs_dev;
  MaskDetectionParameters p;
} MaskPlan;

typedef struct {
  int njobs, stride, u32_need, oob_dark, abt;
  DLineJob *jobs_host, *jobs_dev;
  BorderScanParameters p;
} BorderPlan;

typedef struct {
  int nangles;
  float *rot_host, *tan_host;      /* the reference's angle sequence (deskew.c:156-160) */
  float *rot_dev, *tan_dev;
  float *pair_dev;                 /* optional [2n][2n][4] table: rotation, sin, cos of -rotation */
  int edges[4];
  int peak_off, u32_need, scan_cap;
  long long pre_need;              /* u32 elements of column-prefix scratch per page */
  int run_cap;                     /* max vertical runs of a scan line (0: too many for the warp kernel) */
  int host_tail;                   /* no pair table for this case: average/deviation/sin/cos come from the host's libm */
  DeskewParameters p;
} RotPlan;

/* stream-ordered host evaluation of detect_rotation's float tail (stages.c:rot_host_tail) */
typedef struct { const RotPlan *pl; const DPage *pulled; float *tab; int n, mi; } RotHostJob;

int bf_plan_build(BfPlan *pl, int w, int h, const BlackfilterParameters *p, int abs_black_threshold);
void bf_plan_free(BfPlan *pl);
int blur_plan_build(BlurPlan *pl, int w, int h, const BlurfilterParameters *p, int abs_white);
void blur_plan_free(BlurPlan *pl);
int gray_plan_build(GrayPlan *pl, int w, int h, const GrayfilterParameters *p, int abs_black_threshold);
int mask_plan_build(MaskPlan *pl, int w, int h, const MaskDetectionParameters *p, const Point *pts, int npts);
void mask_plan_free(MaskPlan *pl);
int border_plan_build(BorderPlan *pl, int w, int h, const BorderScanParameters *p,
                      const Rectangle *outside, int n, int abs_black_threshold);
void border_plan_free(BorderPlan *pl);
int rot_plan_build(RotPlan *pl, int w, int h, const DeskewParameters *p, int max_masks, bool with_pair_table);
void rot_plan_free(RotPlan *pl);
/* detect_rotation_cpu's float tail (deskew.c:218-240) on host */
float rot_finalize_host(const RotPlan *pl, const int angle_idx[4]);

/* ---- stages: enqueue one pipeline step for a group of pages -------------- */

typedef struct {
  cudaStream_t st;
  int npages;
  DPage *pages;               /* device array */
  int w, h, fmt;              /* geometry shared by the group (device fmt code) */
  int rows_aligned16;         /* every page: data and pitch are multiples of 16 bytes */
  DFillJob *fillA, *fillB, *fillC; /* device job arrays, npages each */
  DCopyJob *copyA, *copyB;
  DMaskJob *maskJ;
  /* engine only (NULL in the vtable, which finishes detect_rotation on the host itself) */
  RotHostJob *rot_jobs;       /* host memory, D_MAX_MASKS entries that stay valid until the group is done */
  DPage *rot_pull;            /* pinned, npages records */
  float *rot_tab_host, *rot_tab_dev;   /* pinned / device, 4 floats per page */
  int want_ink, ink_fresh;    /* engine: let detect_masks' pass over the sheet also build the ink map; it is current */
  int parity;                 /* engine: which of the slot's two sheet buffers is the working image (same for the whole group) */
  uint64_t launches;
} StageCtx;

void stage_blackfilter(StageCtx *c, const BfPlan *pl);
int stage_noisefilter(StageCtx *c, uint64_t intensity, int white);
void stage_blurfilter(StageCtx *c, const BlurPlan *pl);
int stage_grayfilter(StageCtx *c, const GrayPlan *pl);
void stage_detect_masks(StageCtx *c, const MaskPlan *pl);
int stage_detect_rotation(StageCtx *c, const RotPlan *pl, int max_masks);
int stage_detect_rotation_mask(StageCtx *c, const RotPlan *pl, int mi);
void stage_deskew(StageCtx *c, int interp, int max_masks);
void stage_deskew_mask(StageCtx *c, int interp, int mi);
void stage_center_masks(StageCtx *c, int max_masks);
void stage_detect_border(StageCtx *c, const BorderPlan *pl);
void stage_apply_border_masks(StageCtx *c, Pixel color);
void stage_align_masks(StageCtx *c, const MaskAlignmentParameters *p, int n_outside);
/* sheet-engine forms (two sheet buffers per slot, one sweep per move, see dev.h DPage.other) */
void stage_deskew_mask_pass(StageCtx *c, int interp, int mi);
void stage_center_masks_pass(StageCtx *c, int max_masks);
void stage_align_masks_pass(StageCtx *c, const MaskAlignmentParameters *p, int n_outside, Pixel mask_color,
                            uint8_t *final_dst, size_t final_stride);
void stage_shift_pass(StageCtx *c, Delta d);

/* scratch sizing for one page of w x h in device format fmt */
typedef struct { size_t aux_bytes; int aux_pitch, aux_h; size_t cls_bytes; int list_cap, u32_cap, stack_cap; long long pre_cap; } ScratchNeed;
void scratch_need_all(ScratchNeed *n, int w, int h, int fmt);
int nf_list_cap(int w, int h, uint64_t intensity);   /* noisefilter list entries for this intensity */

void *blob_upload(const void *host, size_t bytes);   /* synchronous H2D into cached device memory */

#ifdef __cplusplus
}
#endif
