/* launch.h — the thin C-ABI between the C host layer and the sm_100a kernels.
 * Every launcher enqueues on `st` and returns immediately; nothing here
 * synchronises or allocates. */
#pragma once
#include <cuda_runtime_api.h>
#include <stdint.h>

#include "dev.h"

#ifdef __cplusplus
extern "C" {
#endif

/* k_blit.cu */
void b200k_fill_jobs(cudaStream_t st, const DFillJob *jobs, int njobs, int maxw, int maxh);
void b200k_copy_jobs(cudaStream_t st, const DCopyJob *jobs, int njobs, int maxw_bytes, int maxh);
void b200k_apply_masks(cudaStream_t st, const DMaskJob *jobs, int njobs, int maxw, int maxh);
void b200k_mirror(cudaStream_t st, DImg im, int dir_h, int dir_v);
void b200k_mirror_pages(cudaStream_t st, DPage *pages, int npages, int maxw, int maxh, int dir_h, int dir_v);
void b200k_rotate90(cudaStream_t st, DImg src, DImg dst, int dir);

/* k_stats.cu */
void b200k_zero_u32(cudaStream_t st, DPage *pages, int npages, int off, int n);
/* gray8_aligned: every page is GRAY8 with 16-byte aligned rows (vector path); want_ink: also leave the
 * ink map of k_inkmap behind when one job covers the whole img_w x img_h image; returns 1 if it did */
int b200k_linesums(cudaStream_t st, DPage *pages, int npages, const DLineJob *jobs_dev,
                   const DLineJob *jobs_host, int njobs, int stat, int lo, int hi, int gray8_aligned, int img_w, int img_h,
                   int want_ink);
void b200k_rect_count(cudaStream_t st, DPage *pages, int npages, const DRect *rects_dev, int nrects,
                      int lo, int hi, int out_off);
int b200k_cellstats(cudaStream_t st, DPage *pages, int npages, int gx, int gy, int ncx, int ncy,
                    int dark_max, int out_off);

/* k_masks.cu */
void b200k_detect_masks(cudaStream_t st, DPage *pages, int npages, int max_points,
                        const int scan_size[2], const int scan_depth[2], const int scan_step[2],
                        const float threshold[2], int dir_h, int dir_v, int sum_off, int sum_stride,
                        int min_w, int max_w, int min_h, int max_h);
void b200k_detect_border(cudaStream_t st, DPage *pages, int npages, int size_w, int size_h, int step_h,
                         int step_v, int thr_h, int thr_v, int dir_h, int dir_v, int sum_off,
                         int sum_stride, int oob_dark);
void b200k_prep_center(cudaStream_t st, DPage *pages, int npages, int i, DFillJob *fill_aux,
                       DCopyJob *copy_out, DFillJob *wipe, DCopyJob *copy_in, DFillJob *wipe2);
void b200k_prep_align(cudaStream_t st, DPage *pages, int npages, int i, int left, int top, int right,
                      int bottom, int margin_h, int margin_v, DFillJob *fill_aux, DCopyJob *copy_out,
                      DFillJob *wipe, DCopyJob *copy_in, DFillJob *wipe2);
void b200k_prep_border_maskjob(cudaStream_t st, DPage *pages, int npages, DMaskJob *jobs, int r, int g, int b);

/* k_filters.cu */
void b200k_bf_scan(cudaStream_t st, DPage *pages, int npages, const DBfPos *pos_dev, int npos,
                   int abs_threshold, long long intensity, int mask_lo, int mask_hi, int flag_off,
                   int maxh /* tallest page of the group: sizes the shared-memory column strip */);
int b200k_noisefilter(cudaStream_t st, DPage *pages, int npages, int maxw, int maxh, int fmt,
                      unsigned long long intensity, int white, int flags /* bit0: rows 16-byte aligned */);
void b200k_blur_decide(cudaStream_t st, DPage *pages, int npages, int n, int nrows,
                       unsigned long long T, float intensity, int cnt_off, int state_off, int flag_off);
void b200k_blur_wipe(cudaStream_t st, DPage *pages, int npages, int n, int nrows, int bw, int bh, int flag_off);
void b200k_gray_cascade(cudaStream_t st, DPage *pages, int npages, const int *gp18, int white_off);

/* k_deskew.cu */
int b200k_rot_peaks(cudaStream_t st, DPage *pages, int npages, int mi_first, int mi_count, const float *tan_tab_dev,
                    int nangles, int scan_size_param, float scan_depth, const int edges[4],
                    int peak_off, int scan_cap, int maxw, int use_prefix, int run_cap /* 0: one CTA per angle */);
void b200k_rot_finalize(cudaStream_t st, DPage *pages, int npages, const float *rot_tab_dev,
                        const float *pair_tab_dev, int nangles, const int edges[4], int peak_off,
                        float deviation, int mi_first, int mi_count);
/* sin/cos of -rotation for mask `mi` of every page from a host-computed table
 * sc[page] = {sin, cos} (only with 3-4 scan edges, where no pair table exists) */
void b200k_rot_set_sincos(cudaStream_t st, DPage *pages, int npages, int mi, const float *sc_dev);
void b200k_rotate(cudaStream_t st, DPage *pages, int npages, int mi, int interp, int maxw, int maxh,
                  DCopyJob *back_jobs);
void b200k_stretch(cudaStream_t st, DImg src, DImg dst, float hr, float vr, int interp);

/* k_engine.cu — sheet-engine helpers */
void b200k_swap_sheets(cudaStream_t st, DPage *pages, int npages);
void b200k_retarget_jobs(cudaStream_t st, const DPage *pages, int npages, DFillJob *fills, int nfill,
                         DMaskJob *masks, int nmask, int stride);
void b200k_set_geometry(cudaStream_t st, DPage *pages, int npages, int w, int h, int pitch);
/* size-changing operations for a group (k_blit.cu / k_deskew.cu): img -> other with a new geometry */
void b200k_rotate90_batch(cudaStream_t st, DImg src, DImg dst, int dir, int nimages, size_t src_stride, size_t dst_stride);
void b200k_rotate90_pages(cudaStream_t st, DPage *pages, int npages, int sw, int sh, int dir, int dpitch);
void b200k_center_pages(cudaStream_t st, DPage *pages, int npages, int dw, int dh, int dpitch);
void b200k_stretch_pages(cudaStream_t st, DPage *pages, int npages, int sw, int sh, int dw, int dh, int dpitch, int interp);
/* one-sweep rectangle move img -> other of every page's DPage.move (k_blit.cu) */
void b200k_move_pass(cudaStream_t st, DPage *pages, int npages, int maxw_bytes, int maxh, int mc_r, int mc_g, int mc_b);
/* k_masks.cu: fill DPage.move for center_mask(i) / align_mask(i) / shift_image */
void b200k_prep_center_move(cudaStream_t st, DPage *pages, int npages, int i);
void b200k_prep_align_move(cudaStream_t st, DPage *pages, int npages, int i, int left, int top, int right,
                           int bottom, int margin_h, int margin_v, int use_masks);
void b200k_prep_shift_move(cudaStream_t st, DPage *pages, int npages, int dx, int dy);
/* k_deskew.cu: deskew() of mask `mi` as one full-sheet pass img -> other */
void b200k_rotate_sheet(cudaStream_t st, DPage *pages, int npages, int mi, int interp, int fmt, int maxw, int maxh,
                        int ink_fresh /* the ink map of the current sheet contents already exists */);
void b200k_page_reset(cudaStream_t st, DPage *pages, int npages);
/* the next pass img -> other of sheet p writes to base + p * stride instead of the slot's other buffer */
void b200k_set_other(cudaStream_t st, DPage *pages, int npages, uint8_t *base, size_t stride);
void b200k_pack_rows(cudaStream_t st, const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                     int row_bytes, int rows, int nimages, size_t src_stride, size_t dst_stride);

/* saveImage()'s output pixel-format conversion (file.c:197-260) of nimages images */
void b200k_convert_out(cudaStream_t st, DImg src, DImg dst, int nimages, size_t src_stride, size_t dst_stride);

#ifdef __cplusplus
}
#endif
