/* engine.c — the sheet engine: process_sheet()'s stage order
 * (reference src/core/sheet_stages.c:44-696) for GROUPS of independent sheets.
 *
 * Replaces the reference's page scheduler (lib/batch_worker.c:79-296: one
 * pthread + one stream per sheet, each blocking on dozens of tiny D2H reads)
 * with asynchronous lanes: a lane owns a stream, a device workspace for
 * `group_pages` sheets and the job/result records; every kernel covers the
 * whole group and every data-dependent decision (masks, rotation, centring,
 * border) stays in the pages' device records.  One host thread keeps all lanes
 * busy; there is no host round-trip inside a sheet.
 *
 * Working-sheet format: the reference always builds an RGB24 sheet
 * (sheet_stages.c:153-155).  For GRAY8 pages with gray background / mask
 * colours every RGB24 pixel would carry r=g=b through all stages, so the
 * engine keeps a 1 byte/pixel sheet instead — same results, a third of the
 * traffic.  RGB24 pages use an RGB24 sheet.
 */
#define _GNU_SOURCE
#include <libavutil/pixfmt.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "host.h"

/* stage switches: bit k of a sheet's skip mask = isExcluded(sheet_nr, no_<k>_multi_index, ignore_multi_index) */
enum { SW_BLACK, SW_NOISE, SW_BLUR, SW_GRAY, SW_MASK_SCAN, SW_MASK_CENTER, SW_DESKEW, SW_WIPE, SW_BORDER,
       SW_BORDER_SCAN, SW_BORDER_ALIGN, SW_COUNT };

enum { STG_DECODE, STG_BLACK, STG_NOISE, STG_BLUR, STG_GRAY, STG_MASKS, STG_ROTDET, STG_DESKEW,
       STG_CENTER, STG_BORDER, STG_OUTPUT, STG_COUNT };
static const char *STG_NAME[STG_COUNT] = {"decode", "blackfilter", "noisefilter", "blurfilter", "grayfilter",
                                          "detect_masks", "detect_rotation", "deskew", "center_mask",
                                          "border", "output"};

typedef struct { int w, h, pitch; } Geo;
enum { GS_ROT90, GS_STRETCH, GS_CENTER };
typedef struct { int op, dir; Geo to; } GeoStep;

typedef struct {
  cudaStream_t st;
  /* up to QD groups are queued on a lane's stream, so the stream never runs dry
   * while the host is busy with another lane; the device workspace is shared
   * (stream order protects it), only the result records are per flight */
  struct Flight {
    cudaEvent_t done;
    cudaEvent_t done_t;   /* timing-enabled twin of `done` */
    cudaEvent_t ev[STG_COUNT + 1];
    int ev_mask;          /* which stage boundaries were recorded */
    DPage *pages_res;     /* pinned */
    const uint8_t *out;   /* where this flight's sheets land (caller memory) */
    int busy, first, n;   /* first: index of the group's first sheet within its stream */
    B200SheetResult *res; /* where this group's results go (NULL: not wanted) */
    unsigned skip;        /* stage switches of this group's sheets */
  } fl[2];
  int slot;               /* flight being issued / collected */
  int ran;
  cudaEvent_t last_done_t;
  uint8_t *sheets, *sheets2, *cls;   /* two sheet buffers per slot: moves render from one into the other */
  uint32_t *list, *u32;
  uint64_t *stack;
  uint32_t *pre;
  uint8_t *ink;
  uint8_t *page_stage;    /* device staging for host-mode pages */
  uint8_t *page_stage2;   /* pre_rotate: the pages after flip_rotate_90 */
  uint8_t *out_stage;     /* device staging for converted output sheets (host mode) */
  DPage *pages_dev, *pages_tmpl /* host */;
  DFillJob *fillA, *fillB, *fillC, *decode_fill;
  DCopyJob *copyA, *copyB, *decode_copy, *decode_copy_host_tmpl;
  DMaskJob *maskJ;
  DFillJob *static_fill[3];   /* pre / mid / post wipe+border rectangles, per page */
  int static_fill_n[3];
  DMaskJob *static_mask[3];
  DRect *static_mask_rects[3];
  uint8_t *out_host; uint8_t *out_dev;
  int host_mode;
  /* detect_rotation's host tail (3-4 scan edges): see stages.c:rot_host_tail */
  RotHostJob rot_jobs[2][D_MAX_MASKS];
  DPage *rot_pull; float *rot_tab_host, *rot_tab_dev;
} Lane;

struct B200Engine {
  B200SheetConfig cfg;
  int device, page_w, page_h, page_fmt, dfmt, bpp;   /* dfmt/bpp: the working sheet */
  int page_dfmt;                                      /* the pages as they arrive */
  int sheet_w, sheet_h, sheet_pitch, page_row, sheet_row;   /* the PIPELINE geometry (after the pre-stage stretch / resize) */
  size_t sheet_stride, page_bytes;
  /* size-changing options (sheet_stages.c:134-145, :216-230, :511-531): every sheet of an engine goes through the
   * same chain of geometries, fixed at creation.  in: the decoded sheet; steps: passes img -> other with a new size */
  Geo g_in, g_out;
  GeoStep pre_steps[3], post_steps[4];
  int n_pre, n_post, direct_upload;
  int pre_rot_dir, rp_w, rp_h, rp_row;      /* pre_rotate: the pages after flip_rotate_90 */
  size_t rp_bytes;
  int group, nlanes;
  int npoints, noutside;
  Point points[D_MAX_MASKS];
  Rectangle outside[D_MAX_BORDERS];
  Rectangle bf_excl[MAX_MASKS];
  BfPlan bf; BlurPlan blur; GrayPlan gray; MaskPlan mask; BorderPlan border; RotPlan rot;
  ScratchNeed need;
  Lane *lanes;
  uint64_t launches;
  int profiling;
  int n_static_mask_jobs[3];
  int bad_sheets, bad_first;
  unsigned bad_flags;
  cudaEvent_t ev_begin;
  int out_fmt, out_dfmt, out_row;   /* output conversion (av format, -1 = none); out_row: tight row of ONE output image */
  int out_count, out_w;             /* sheet split (sheet_stages.c:606-621): images per sheet, width of each */
  /* per-sheet stage switches (options->no_*_multi_index + ignore_multi_index) */
  int first_sheet_nr;
  int32_t *sw_idx[SW_COUNT + 1]; int sw_n[SW_COUNT + 1];
  B200SheetDoneFn done_fn; void *done_user; int done_failed;
  double last_device_ms;
  double stage_ms[STG_COUNT], stage_ms_min[STG_COUNT], stage_ms_max[STG_COUNT];
  uint64_t stage_groups[STG_COUNT];
  /* stream state: groups [g_done, g) are in flight, at most two per lane */
  int streaming, host_mode, g, g_done, fed;
  int issue_failed;   /* a stage launcher refused (limits are checked at creation, so this is a bug trap) */
};

static int imax(int a, int b) { return a > b ? a : b; }

int unpaper_b200_engine_sheet_width(const B200Engine *e) { return e->g_out.w; }
int unpaper_b200_engine_sheet_height(const B200Engine *e) { return e->g_out.h; }
size_t unpaper_b200_engine_sheet_bytes(const B200Engine *e) {
  return (size_t)e->out_row * e->g_out.h * e->out_count;
}
int unpaper_b200_engine_output_width(const B200Engine *e) { return e->out_w; }
int unpaper_b200_engine_output_count(const B200Engine *e) { return e->out_count; }
void unpaper_b200_engine_set_sheet_callback(B200Engine *e, B200SheetDoneFn fn, void *user) {
  if (!e) return;
  e->done_fn = fn; e->done_user = user;
}
int unpaper_b200_engine_output_format(const B200Engine *e) {
  return e->out_fmt >= 0 ? e->out_fmt : (e->dfmt == DF_GRAY8 ? AV_PIX_FMT_GRAY8 : AV_PIX_FMT_RGB24);
}

/* sheet_stage_output's conversion (sheet_stages.c:536-631 -> file.c:197-260) moved onto
 * the device, so that the D2H (or the device-side sink) carries the output format */
int unpaper_b200_engine_set_output_format(B200Engine *e, int av_pix_fmt) {
  if (!e) return -1;
  unpaper_b200_set_device(e->device);
  CUDA_OK(cudaDeviceSynchronize());
  /* -1: what sheet_stage_output picks without an explicit type, the page format
   * (sheet_stages.c:130-131), as saveImage() writes it (file.c:201-208) */
  int fmt = unpaper_b200_output_format(av_pix_fmt < 0 ? e->page_fmt : av_pix_fmt);
  if (fmt == (e->dfmt == DF_GRAY8 ? AV_PIX_FMT_GRAY8 : AV_PIX_FMT_RGB24)) fmt = -1;   /* the working sheet as it is */
  int df = e->dfmt, row = e->out_w * e->bpp;
  if (fmt >= 0) {
    df = b200_fmt_to_dev(fmt); row = b200_fmt_row_bytes(fmt, e->out_w);
    if (df < 0 || row <= 0) { b200_set_error("engine: unsupported output format %d", av_pix_fmt); return -1; }
  }
  e->out_dfmt = df; e->out_row = row;
  e->out_fmt = fmt;
  for (int i = 0; i < e->nlanes; i++) {
    Lane *ln = &e->lanes[i];
    if (ln->out_stage) { b200_dev_free(ln->out_stage); ln->out_stage = NULL; }
    if (fmt >= 0 || e->out_count > 1)
      ln->out_stage = (uint8_t *)b200_dev_alloc((size_t)row * e->g_out.h * e->out_count * e->group + 64);
  }
  return 0;
}
uint64_t unpaper_b200_engine_launch_count(const B200Engine *e) { return e->launches; }
double unpaper_b200_engine_last_device_ms(const B200Engine *e) { return e->last_device_ms; }
int unpaper_b200_engine_set_profiling(B200Engine *e, int enabled) {
  e->profiling = enabled;
  memset(e->stage_ms, 0, sizeof(e->stage_ms));
  memset(e->stage_groups, 0, sizeof(e->stage_groups));
  return 0;
}
int unpaper_b200_engine_get_profile(const B200Engine *e, int max_entries, const char **names, double *ms,
                                    uint64_t *launches, double *alg_bytes) {
  int n = 0;
  for (int s = 0; s < STG_COUNT && n < max_entries; s++) {
    names[n] = STG_NAME[s]; ms[n] = e->stage_ms[s]; launches[n] = e->stage_groups[s];
    if (alg_bytes) alg_bytes[n] = 0;
    n++;
  }
  return n;
}

int unpaper_b200_engine_get_profile_spread(const B200Engine *e, int max_entries, double *min_ms, double *max_ms) {
  int n = 0;
  for (int s = 0; s < STG_COUNT && n < max_entries; s++, n++) { min_ms[n] = e->stage_ms_min[s]; max_ms[n] = e->stage_ms_max[s]; }
  return n;
}

/* fill rectangles that the reference applies with apply_wipes / apply_border
 * as static jobs (geometry does not depend on the page contents) */
static int build_static(B200Engine *e, Lane *ln, int slot, const Rectangle *wipes, int nw, Border border,
                        const Rectangle *masks, int nm, const Rectangle *extra_wipe) {
  int P = e->group;
  int nfill = nw + (extra_wipe ? 1 : 0);
  ln->static_fill_n[slot] = nfill;
  ln->static_fill[slot] = NULL; ln->static_mask[slot] = NULL; ln->static_mask_rects[slot] = NULL;
  if (nfill > 0) {
    DFillJob *h = (DFillJob *)calloc((size_t)nfill * P, sizeof(DFillJob));
    for (int k = 0; k < nfill; k++) {
      Rectangle r = k < nw ? wipes[k] : *extra_wipe;
      for (int p = 0; p < P; p++) {
        DFillJob *j = &h[(size_t)k * P + p];
        j->img = ln->pages_tmpl[p].img;
        j->r = (DRect){r.vertex[0].x, r.vertex[0].y, r.vertex[1].x, r.vertex[1].y};   /* apply_wipes: as given */
        j->c[0] = e->cfg.mask_color.r; j->c[1] = e->cfg.mask_color.g; j->c[2] = e->cfg.mask_color.b;
        j->enabled = 1;
      }
    }
    ln->static_fill[slot] = (DFillJob *)blob_upload(h, (size_t)nfill * P * sizeof(DFillJob));
    free(h);
  }
  bool has_border = border.left || border.top || border.right || border.bottom;
  int nrect = nm + (has_border ? 1 : 0);
  if (nrect > 0) {
    /* apply_masks(pre_masks) and apply_border are separate calls in the
     * reference; both are "paint outside these rectangles" and run here as two
     * consecutive jobs when both are present */
    DRect *hr = (DRect *)calloc((size_t)nrect, sizeof(DRect));
    for (int k = 0; k < nm; k++) hr[k] = (DRect){masks[k].vertex[0].x, masks[k].vertex[0].y, masks[k].vertex[1].x, masks[k].vertex[1].y};
    if (has_border) hr[nm] = (DRect){border.left, border.top, e->sheet_w - border.right - 1, e->sheet_h - border.bottom - 1};
    ln->static_mask_rects[slot] = (DRect *)blob_upload(hr, (size_t)nrect * sizeof(DRect));
    free(hr);
    int njob = (nm > 0 ? 1 : 0) + (has_border ? 1 : 0);
    DMaskJob *hj = (DMaskJob *)calloc((size_t)njob * P, sizeof(DMaskJob));
    int q = 0;
    if (nm > 0) { for (int p = 0; p < P; p++) { DMaskJob *j = &hj[(size_t)q * P + p]; j->img = ln->pages_tmpl[p].img; j->rects = ln->static_mask_rects[slot]; j->nrects = nm; j->enabled = 1;
        j->c[0] = e->cfg.mask_color.r; j->c[1] = e->cfg.mask_color.g; j->c[2] = e->cfg.mask_color.b; } q++; }
    if (has_border) { for (int p = 0; p < P; p++) { DMaskJob *j = &hj[(size_t)q * P + p]; j->img = ln->pages_tmpl[p].img; j->rects = ln->static_mask_rects[slot] + nm; j->nrects = 1; j->enabled = 1;
        j->c[0] = e->cfg.mask_color.r; j->c[1] = e->cfg.mask_color.g; j->c[2] = e->cfg.mask_color.b; } q++; }
    ln->static_mask[slot] = (DMaskJob *)blob_upload(hj, (size_t)njob * P * sizeof(DMaskJob));
    free(hj);
    return njob;
  }
  return 0;
}

static void lane_free(Lane *ln) {
  void *ptrs[] = {ln->sheets, ln->sheets2, ln->cls, ln->list, ln->u32, ln->stack, ln->pre, ln->ink, ln->page_stage, ln->page_stage2, ln->out_stage, ln->pages_dev,
                  ln->fillA, ln->fillB, ln->fillC, ln->decode_fill, ln->copyA, ln->copyB, ln->decode_copy, ln->maskJ,
                  ln->static_fill[0], ln->static_fill[1], ln->static_fill[2], ln->static_mask[0], ln->static_mask[1],
                  ln->static_mask[2], ln->static_mask_rects[0], ln->static_mask_rects[1], ln->static_mask_rects[2]};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) if (ptrs[i]) b200_dev_free(ptrs[i]);
  free(ln->pages_tmpl); free(ln->decode_copy_host_tmpl);
  for (ln->slot = 0; ln->slot < 2; ln->slot++) {
    if (ln->fl[ln->slot].pages_res) b200_pinned_free(ln->fl[ln->slot].pages_res);
    if (ln->fl[ln->slot].done) cudaEventDestroy(ln->fl[ln->slot].done);
    if (ln->fl[ln->slot].done_t) cudaEventDestroy(ln->fl[ln->slot].done_t);
    for (int i = 0; i <= STG_COUNT; i++) if (ln->fl[ln->slot].ev[i]) cudaEventDestroy(ln->fl[ln->slot].ev[i]);
  }
  if (ln->rot_pull) b200_pinned_free(ln->rot_pull);
  if (ln->rot_tab_host) b200_pinned_free(ln->rot_tab_host);
  if (ln->rot_tab_dev) b200_dev_free(ln->rot_tab_dev);
  if (ln->st) b200_stream_release(ln->st);
  memset(ln, 0, sizeof(*ln));
}

void unpaper_b200_engine_destroy(B200Engine *e) {
  if (!e) return;
  unpaper_b200_set_device(e->device);
  if (e->lanes) { for (int i = 0; i < e->nlanes; i++) lane_free(&e->lanes[i]); free(e->lanes); }
  if (e->ev_begin) cudaEventDestroy(e->ev_begin);
  for (int k = 0; k <= SW_COUNT; k++) free(e->sw_idx[k]);
  bf_plan_free(&e->bf); blur_plan_free(&e->blur); mask_plan_free(&e->mask); border_plan_free(&e->border); rot_plan_free(&e->rot);
  free(e);
}


B200Engine *unpaper_b200_engine_create(const B200SheetConfig *cfg, int device, int page_w, int page_h,
                                       int page_format, int group_pages, int lanes) {
  if (!cfg || page_w <= 0 || page_h <= 0 || group_pages <= 0 || lanes <= 0) { b200_set_error("engine: bad arguments"); return NULL; }
  if (unpaper_b200_set_device(device)) return NULL;
  /* 1-bit pages (pbm scans, the PDF path's expand_1bit_to_8bit) are expanded into the
   * 1 B/px working sheet by the decode stage's copy and leave as MONOWHITE again */
  bool mono_pages = page_format == AV_PIX_FMT_MONOWHITE || page_format == AV_PIX_FMT_MONOBLACK;
  if (page_format != AV_PIX_FMT_GRAY8 && page_format != AV_PIX_FMT_RGB24 && !mono_pages) {
    b200_set_error("engine: page format must be GRAY8, RGB24, MONOWHITE or MONOBLACK"); return NULL;
  }
  if (cfg->input_count < 1 || cfg->input_count > 2) { b200_set_error("engine: input_count must be 1 or 2"); return NULL; }
  if (cfg->output_count < 0 || cfg->output_count > 2) { b200_set_error("engine: output_count must be 1 or 2"); return NULL; }
  /* the fixed arrays of B200SheetConfig and of the device page record */
  if (cfg->point_count < 0 || cfg->point_count > 8 || cfg->pre_mask_count < 0 || cfg->pre_mask_count > 8 ||
      cfg->pre_wipe_count < 0 || cfg->pre_wipe_count > 8 || cfg->wipe_count < 0 || cfg->wipe_count > 8 ||
      cfg->post_wipe_count < 0 || cfg->post_wipe_count > 8) {
    b200_set_error("engine: point / pre-mask / wipe counts must be within 0..8"); return NULL;
  }
  if (cfg->layout == LAYOUT_DOUBLE && cfg->point_count == 0 && D_MAX_MASKS < 2) { b200_set_error("engine: too many points"); return NULL; }
  if (cfg->blackfilter.exclusions_count > MAX_MASKS || (cfg->blackfilter.exclusions_count > 0 && !cfg->blackfilter.exclusions)) {
    b200_set_error("engine: bad blackfilter exclusion list"); return NULL;
  }
  {
    const B200MultiIndex *mi[SW_COUNT + 1] = {&cfg->no_blackfilter_sheets, &cfg->no_noisefilter_sheets, &cfg->no_blurfilter_sheets,
      &cfg->no_grayfilter_sheets, &cfg->no_mask_scan_sheets, &cfg->no_mask_center_sheets, &cfg->no_deskew_sheets, &cfg->no_wipe_sheets,
      &cfg->no_border_sheets, &cfg->no_border_scan_sheets, &cfg->no_border_align_sheets, &cfg->ignore_sheets};
    for (int k = 0; k <= SW_COUNT; k++)
      if (mi[k]->count < -1 || (mi[k]->count > 0 && !mi[k]->indexes)) { b200_set_error("engine: bad per-sheet switch list %d", k); return NULL; }
  }
  bool gray_colors = cfg->sheet_background.r == cfg->sheet_background.g && cfg->sheet_background.g == cfg->sheet_background.b &&
                     cfg->mask_color.r == cfg->mask_color.g && cfg->mask_color.g == cfg->mask_color.b;
  if (page_format != AV_PIX_FMT_RGB24 && !gray_colors) { b200_set_error("engine: gray and 1-bit pages need gray background and mask colours"); return NULL; }
  B200Engine *e = (B200Engine *)calloc(1, sizeof(*e));
  e->cfg = *cfg; e->device = device;
  e->page_w = page_w; e->page_h = page_h; e->page_fmt = page_format; e->page_dfmt = b200_fmt_to_dev(page_format);
  e->dfmt = page_format == AV_PIX_FMT_RGB24 ? DF_RGB24 : DF_GRAY8;
  e->out_fmt = -1;
  e->bpp = e->dfmt == DF_GRAY8 ? 1 : 3;
  e->first_sheet_nr = cfg->first_sheet_nr > 0 ? cfg->first_sheet_nr : 1;
  {
    /* the lists are copied; a switch that is off for every sheet (uint8 flag, or count -1) stays in cfg->no_* */
    const B200MultiIndex *mi[SW_COUNT + 1] = {&cfg->no_blackfilter_sheets, &cfg->no_noisefilter_sheets, &cfg->no_blurfilter_sheets,
      &cfg->no_grayfilter_sheets, &cfg->no_mask_scan_sheets, &cfg->no_mask_center_sheets, &cfg->no_deskew_sheets, &cfg->no_wipe_sheets,
      &cfg->no_border_sheets, &cfg->no_border_scan_sheets, &cfg->no_border_align_sheets, &cfg->ignore_sheets};
    uint8_t *flag[SW_COUNT] = {&e->cfg.no_blackfilter, &e->cfg.no_noisefilter, &e->cfg.no_blurfilter, &e->cfg.no_grayfilter,
      &e->cfg.no_mask_scan, &e->cfg.no_mask_center, &e->cfg.no_deskew, &e->cfg.no_wipe, &e->cfg.no_border,
      &e->cfg.no_border_scan, &e->cfg.no_border_align};
    for (int k = 0; k <= SW_COUNT; k++) {
      if (mi[k]->count == -1) {
        if (k < SW_COUNT) *flag[k] = 1; else for (int q = 0; q < SW_COUNT; q++) *flag[q] = 1;
      } else if (mi[k]->count > 0) {
        e->sw_n[k] = mi[k]->count;
        e->sw_idx[k] = (int32_t *)malloc(sizeof(int32_t) * (size_t)mi[k]->count);
        memcpy(e->sw_idx[k], mi[k]->indexes, sizeof(int32_t) * (size_t)mi[k]->count);
      }
    }
    memset(&e->cfg.no_blackfilter_sheets, 0, sizeof(B200MultiIndex) * (SW_COUNT + 1));   /* no dangling caller pointers */
    cfg = &e->cfg;
  }
  e->group = group_pages; e->nlanes = lanes;
  e->page_row = b200_fmt_row_bytes(page_format, page_w);
  e->page_bytes = (size_t)e->page_row * page_h;
  /* pages after options->pre_rotate (sheet_stages.c:134-137) */
  e->pre_rot_dir = cfg->pre_rotate / 90;
  e->rp_w = e->pre_rot_dir ? page_h : page_w; e->rp_h = e->pre_rot_dir ? page_w : page_h;
  e->rp_row = b200_fmt_row_bytes(page_format, e->rp_w); e->rp_bytes = (size_t)e->rp_row * e->rp_h;
  {
    /* the chain of sheet geometries: decoded sheet = input pages side by side unless --sheet-size says
     * otherwise (:139-145); pre stage: stretch (+ zoom), resize (:216-230); post stage: rotate, stretch
     * (+ zoom), resize (:511-531) */
    const int bpp = e->bpp;
#define GEO(W_, H_) ((Geo){(W_), (H_), (((W_) * bpp + 15) & ~15)})
    RectangleSize cur = {cfg->sheet_size.width == -1 ? e->rp_w * cfg->input_count : cfg->sheet_size.width,
                         cfg->sheet_size.height == -1 ? e->rp_h : cfg->sheet_size.height};
    e->g_in = GEO(cur.width, cur.height);
    for (int stage = 0; stage < 2; stage++) {
      GeoStep *steps = stage ? e->post_steps : e->pre_steps;
      int ns = 0;
      RectangleSize st_size = stage ? cfg->post_stretch_size : cfg->stretch_size;
      RectangleSize pg_size = stage ? cfg->post_page_size : cfg->page_size;
      float zoom = stage ? cfg->post_zoom_factor : cfg->pre_zoom_factor;
      if (zoom == 0.0f) zoom = 1.0f;
      if (stage && cfg->post_rotate / 90 != 0) {
        steps[ns++] = (GeoStep){GS_ROT90, cfg->post_rotate / 90, GEO(cur.height, cur.width)};
        cur = (RectangleSize){cur.height, cur.width};
      }
      RectangleSize sz = {st_size.width == -1 ? cur.width : st_size.width, st_size.height == -1 ? cur.height : st_size.height};
      sz.width *= zoom; sz.height *= zoom;                       /* int *= float, like :219-220 */
      if (sz.width != cur.width || sz.height != cur.height) { steps[ns++] = (GeoStep){GS_STRETCH, 0, GEO(sz.width, sz.height)}; cur = sz; }
      if (pg_size.width != -1 || pg_size.height != -1) {
        RectangleSize size = {pg_size.width == -1 ? cur.width : pg_size.width, pg_size.height == -1 ? cur.height : pg_size.height};
        if (size.width != cur.width || size.height != cur.height) {
          /* resize_and_replace (blit.c:244-283): stretch keeping the aspect, then centre on a new sheet */
          const float hr = (float)size.width / (float)cur.width, vr = (float)size.height / (float)cur.height;
          RectangleSize ss;
          if (hr < vr) ss = (RectangleSize){size.width, cur.height * hr};
          else if (vr < hr) ss = (RectangleSize){cur.width * vr, size.height};
          else ss = size;
          if (ss.width != cur.width || ss.height != cur.height) { steps[ns++] = (GeoStep){GS_STRETCH, 0, GEO(ss.width, ss.height)}; cur = ss; }
          if (size.width != ss.width || size.height != ss.height) { steps[ns++] = (GeoStep){GS_CENTER, 0, GEO(size.width, size.height)}; cur = size; }
        }
      }
      if (stage) e->n_post = ns; else e->n_pre = ns;
      if (cur.width <= 0 || cur.height <= 0) { b200_set_error("engine: a size option gives an empty sheet"); free(e); return NULL; }
      if (!stage) { e->sheet_w = cur.width; e->sheet_h = cur.height; }
    }
    e->g_out = GEO(cur.width, cur.height);
#undef GEO
  }
  e->sheet_row = e->sheet_w * e->bpp;
  e->out_count = cfg->output_count == 2 ? 2 : 1;
  e->out_w = e->g_out.w / e->out_count;          /* sheet_stages.c:612 */
  e->out_dfmt = e->dfmt; e->out_row = e->out_w * e->bpp;
  e->sheet_pitch = (e->sheet_row + 15) & ~15;
  {
    size_t big = (size_t)e->g_in.pitch * e->g_in.h;
    for (int i = 0; i < e->n_pre; i++) { size_t b = (size_t)e->pre_steps[i].to.pitch * e->pre_steps[i].to.h; if (b > big) big = b; }
    for (int i = 0; i < e->n_post; i++) { size_t b = (size_t)e->post_steps[i].to.pitch * e->post_steps[i].to.h; if (b > big) big = b; }
    e->sheet_stride = ((big + 64) + 255) & ~(size_t)255;
  }
  /* the upload can be the decode stage's centre copy when a page IS the decoded sheet */
  e->direct_upload = cfg->input_count == 1 && !e->pre_rot_dir && e->g_in.w == page_w && e->g_in.h == page_h &&
                     e->g_in.pitch == e->page_row && e->page_dfmt == e->dfmt;
  int W = e->sheet_w, H = e->sheet_h;

  /* layout-derived points, mask maxima, border-scan areas (sheet_stages.c:232-279) */
  MaskDetectionParameters mp = cfg->mask_detection;
  e->npoints = cfg->point_count;
  for (int i = 0; i < e->npoints && i < D_MAX_MASKS; i++) e->points[i] = cfg->points[i];
  if (cfg->layout == LAYOUT_SINGLE) {
    if (e->npoints == 0) e->points[e->npoints++] = (Point){W / 2, H / 2};
    if (mp.maximum_width == -1) mp.maximum_width = W;
    if (mp.maximum_height == -1) mp.maximum_height = H;
    e->outside[e->noutside++] = (Rectangle){{{0, 0}, {W - 1, H - 1}}};
  } else if (cfg->layout == LAYOUT_DOUBLE) {
    if (e->npoints == 0) { e->points[e->npoints++] = (Point){W / 4, H / 2}; e->points[e->npoints++] = (Point){W - W / 4, H / 2}; }
    if (mp.maximum_width == -1) mp.maximum_width = W / 2;
    if (mp.maximum_height == -1) mp.maximum_height = H;
    e->outside[e->noutside++] = (Rectangle){{{0, 0}, {W / 2, H - 1}}};
    e->outside[e->noutside++] = (Rectangle){{{W / 2, 0}, {W - 1, H - 1}}};
  }
  if (mp.maximum_width == -1) mp.maximum_width = W;
  if (mp.maximum_height == -1) mp.maximum_height = H;

  /* blackfilter exclusions (sheet_stages.c:298-322) */
  BlackfilterParameters bfp = cfg->blackfilter;
  size_t nex = 0;
  for (size_t i = 0; i < cfg->blackfilter.exclusions_count && i < MAX_MASKS; i++) e->bf_excl[nex++] = cfg->blackfilter.exclusions[i];
  if (nex == 0 && cfg->layout != LAYOUT_NONE) {
    if (cfg->layout == LAYOUT_SINGLE) {
      e->bf_excl[nex++] = (Rectangle){{{W / 4, H / 4}, {W / 4 + W / 2 - 1, H / 4 + H / 2 - 1}}};
    } else {
      int fw = W / 4, fh = H / 2, ox = W / 8, oy = H / 4;
      e->bf_excl[nex++] = (Rectangle){{{ox, oy}, {ox + fw - 1, oy + fh - 1}}};
      e->bf_excl[nex++] = (Rectangle){{{ox + W / 2, oy}, {ox + W / 2 + fw - 1, oy + fh - 1}}};
    }
  }
  bfp.exclusions = e->bf_excl; bfp.exclusions_count = nex;

  int rc = 0;
  if (!cfg->no_blackfilter) rc |= bf_plan_build(&e->bf, W, H, &bfp, cfg->abs_black_threshold);
  if (!cfg->no_blurfilter) rc |= blur_plan_build(&e->blur, W, H, &cfg->blurfilter, cfg->abs_white_threshold);
  if (!cfg->no_grayfilter) rc |= gray_plan_build(&e->gray, W, H, &cfg->grayfilter, cfg->abs_black_threshold);
  if (!cfg->no_mask_scan) rc |= mask_plan_build(&e->mask, W, H, &mp, e->points, e->npoints);
  if (!cfg->no_border_scan) rc |= border_plan_build(&e->border, W, H, &cfg->border_scan, e->outside, e->noutside, cfg->abs_black_threshold);
  if (!cfg->no_deskew) rc |= rot_plan_build(&e->rot, W, H, &cfg->deskew, imax(e->npoints, 1), true);
  if (!cfg->no_noisefilter && (H >= 32768 || W >= 65536 || cfg->noisefilter_intensity > 4000)) {
    b200_set_error("engine: noisefilter limits exceeded"); rc = -1;
  }
  if (rc) { unpaper_b200_engine_destroy(e); return NULL; }

  scratch_need_all(&e->need, W, H, e->dfmt);
  if (!cfg->no_noisefilter) e->need.list_cap = nf_list_cap(W, H, cfg->noisefilter_intensity);
  int u32 = 64;
  u32 = imax(u32, e->bf.u32_need); u32 = imax(u32, e->blur.u32_need); u32 = imax(u32, e->gray.u32_need);
  u32 = imax(u32, e->mask.u32_need); u32 = imax(u32, e->border.u32_need); u32 = imax(u32, e->rot.u32_need);
  e->need.u32_cap = (u32 + 63) & ~63;
  e->need.pre_cap = cfg->no_deskew ? 0 : (e->rot.pre_need + 63) & ~63LL;

  int P = group_pages;
  e->lanes = (Lane *)calloc((size_t)lanes, sizeof(Lane));
  for (int li = 0; li < lanes; li++) {
    Lane *ln = &e->lanes[li];
    ln->st = b200_stream_acquire();
    for (ln->slot = 0; ln->slot < 2; ln->slot++) {
      CUDA_OK(cudaEventCreateWithFlags(&ln->fl[ln->slot].done, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreate(&ln->fl[ln->slot].done_t));
      for (int i = 0; i <= STG_COUNT; i++) CUDA_OK(cudaEventCreate(&ln->fl[ln->slot].ev[i]));
      ln->fl[ln->slot].pages_res = (DPage *)b200_pinned_alloc(sizeof(DPage) * P);
    }
    ln->slot = 0;
    ln->sheets = (uint8_t *)b200_dev_alloc(e->sheet_stride * P);
    ln->sheets2 = (uint8_t *)b200_dev_alloc(e->sheet_stride * P);
    size_t cls_stride = (e->need.cls_bytes + 255) & ~(size_t)255;
    ln->cls = (uint8_t *)b200_dev_alloc(cls_stride * P);
    ln->list = (uint32_t *)b200_dev_alloc((size_t)e->need.list_cap * 4 * P);
    ln->u32 = (uint32_t *)b200_dev_alloc((size_t)e->need.u32_cap * 4 * P);
    ln->stack = (uint64_t *)b200_dev_alloc((size_t)e->need.stack_cap * 32 * P);
    if (e->need.pre_cap > 0) ln->pre = (uint32_t *)b200_dev_alloc((size_t)e->need.pre_cap * 4 * P);
    int ink_cells = ((W + D_INK_CELL - 1) / D_INK_CELL) * ((H + D_INK_CELL - 1) / D_INK_CELL);
    size_t ink_stride = ((size_t)ink_cells + 255) & ~(size_t)255;
    ln->ink = (uint8_t *)b200_dev_alloc(ink_stride * P);
    ln->page_stage = (uint8_t *)b200_dev_alloc(e->page_bytes * cfg->input_count * P + 64);
    if (e->pre_rot_dir) ln->page_stage2 = (uint8_t *)b200_dev_alloc(e->rp_bytes * cfg->input_count * P + 64);
    ln->pages_dev = (DPage *)b200_dev_alloc(sizeof(DPage) * P);
    if (!cfg->no_deskew && e->rot.host_tail) {
      ln->rot_pull = (DPage *)b200_pinned_alloc(sizeof(DPage) * P);
      ln->rot_tab_host = (float *)b200_pinned_alloc(sizeof(float) * 4 * P);
      ln->rot_tab_dev = (float *)b200_dev_alloc(sizeof(float) * 4 * P);
    }
    ln->pages_tmpl = (DPage *)calloc((size_t)P, sizeof(DPage));
    ln->fillA = (DFillJob *)b200_dev_alloc(sizeof(DFillJob) * P);
    ln->fillB = (DFillJob *)b200_dev_alloc(sizeof(DFillJob) * P);
    ln->fillC = (DFillJob *)b200_dev_alloc(sizeof(DFillJob) * P);
    ln->copyA = (DCopyJob *)b200_dev_alloc(sizeof(DCopyJob) * P);
    ln->copyB = (DCopyJob *)b200_dev_alloc(sizeof(DCopyJob) * P);
    ln->maskJ = (DMaskJob *)b200_dev_alloc(sizeof(DMaskJob) * P);
    ln->decode_fill = (DFillJob *)b200_dev_alloc(sizeof(DFillJob) * P);
    ln->decode_copy = (DCopyJob *)b200_dev_alloc(sizeof(DCopyJob) * P * cfg->input_count);
    ln->decode_copy_host_tmpl = (DCopyJob *)calloc((size_t)P * cfg->input_count, sizeof(DCopyJob));
    DFillJob *dfill = (DFillJob *)calloc((size_t)P, sizeof(DFillJob));
    for (int p = 0; p < P; p++) {
      DPage *pg = &ln->pages_tmpl[p];
      pg->img = (DImg){ln->sheets + e->sheet_stride * p, e->g_in.w, e->g_in.h, e->g_in.pitch, e->dfmt, cfg->abs_black_threshold,
                       {cfg->sheet_background.r, cfg->sheet_background.g, cfg->sheet_background.b}};
      pg->buf[0] = pg->img.data; pg->buf[1] = ln->sheets2 + e->sheet_stride * p;
      pg->other = pg->buf[1];
      pg->cls = ln->cls + cls_stride * p;
      pg->list = ln->list + (size_t)e->need.list_cap * p; pg->list_cap = e->need.list_cap;
      pg->u32 = ln->u32 + (size_t)e->need.u32_cap * p; pg->u32_cap = e->need.u32_cap;
      pg->stack = ln->stack + (size_t)e->need.stack_cap * 4 * p; pg->stack_cap = e->need.stack_cap;
      if (ln->pre) { pg->pre = ln->pre + (size_t)e->need.pre_cap * p; pg->pre_cap = e->need.pre_cap; }
      pg->ink = ln->ink + ink_stride * p; pg->ink_cap = ink_cells;
      pg->point_count = e->npoints;
      for (int i = 0; i < e->npoints; i++) { pg->px[i] = e->points[i].x; pg->py[i] = e->points[i].y; }
      pg->outside_count = e->noutside;
      for (int i = 0; i < e->noutside; i++)
        pg->outside[i] = (DRect){e->outside[i].vertex[0].x, e->outside[i].vertex[0].y, e->outside[i].vertex[1].x, e->outside[i].vertex[1].y};
      for (int i = 0; i < D_MAX_MASKS; i++) pg->rot_cos[i] = 1.0f;
      /* decode: create_image(fill) + center_image per input (sheet_stages.c:150-165).
       * Pages have the slot's size here, so center_image never wipes and the
       * fill is only needed when the sheet pitch has padding columns (none read). */
      dfill[p].img = pg->img; dfill[p].r = (DRect){0, 0, e->g_in.w - 1, e->g_in.h - 1};
      dfill[p].c[0] = cfg->sheet_background.r; dfill[p].c[1] = cfg->sheet_background.g; dfill[p].c[2] = cfg->sheet_background.b;
      /* create_image(fill): only visible where the pages do not cover the sheet */
      dfill[p].enabled = e->rp_w * cfg->input_count != e->g_in.w || e->rp_h != e->g_in.h;
      for (int j = 0; j < cfg->input_count; j++) {
        DCopyJob *cj = &ln->decode_copy_host_tmpl[(size_t)p * cfg->input_count + j];
        cj->src = (DImg){NULL, e->rp_w, e->rp_h, e->rp_row, e->page_dfmt, cfg->abs_black_threshold, {255, 255, 255}};
        cj->dst = pg->img;
        /* center_image(page, sheet, cell origin, cell size) (blit.c:175-207; sheet_stages.c:160-164) */
        int cw = e->g_in.w / cfg->input_count, chh = e->g_in.h;
        int sx = 0, sy = 0, sw2 = e->rp_w, sh2 = e->rp_h, tx = e->g_in.w * j / cfg->input_count, ty = 0;
        if (sw2 <= cw) tx += (cw - sw2) / 2; else { sx += (sw2 - cw) / 2; sw2 = cw; }
        if (sh2 <= chh) ty += (chh - sh2) / 2; else { sy += (sh2 - chh) / 2; sh2 = chh; }
        cj->area = (DRect){sx, sy, sx + sw2 - 1, sy + sh2 - 1};
        cj->tx = tx; cj->ty = ty;
        cj->enabled = 1;
      }
    }
    CUDA_OK(cudaMemcpy(ln->pages_dev, ln->pages_tmpl, sizeof(DPage) * P, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ln->decode_fill, dfill, sizeof(DFillJob) * P, cudaMemcpyHostToDevice));
    free(dfill);
    /* static wipes / borders at the three points where the reference applies them */
    Rectangle mid = {{{W / 2 - cfg->middle_wipe[0], 0}, {W / 2 + cfg->middle_wipe[1], H - 1}}};
    bool use_mid = cfg->layout == LAYOUT_DOUBLE && (cfg->middle_wipe[0] > 0 || cfg->middle_wipe[1] > 0);
    Border none = {0, 0, 0, 0};
    e->n_static_mask_jobs[0] = build_static(e, ln, 0, cfg->pre_wipes, cfg->no_wipe ? 0 : cfg->pre_wipe_count,
                                         cfg->no_border ? none : cfg->pre_border, cfg->pre_masks, cfg->pre_mask_count, NULL);
    e->n_static_mask_jobs[1] = build_static(e, ln, 1, cfg->wipes, cfg->no_wipe ? 0 : cfg->wipe_count,
                                         cfg->no_border ? none : cfg->border, NULL, 0, (use_mid && !cfg->no_wipe) ? &mid : NULL);
    e->n_static_mask_jobs[2] = build_static(e, ln, 2, cfg->post_wipes, cfg->no_wipe ? 0 : cfg->post_wipe_count,
                                         cfg->no_border ? none : cfg->post_border, NULL, 0, NULL);
  }
  /* 1-bit pages leave in their own format unless the caller asks otherwise (this also
   * allocates the output staging a split sheet needs) */
  if ((mono_pages || e->out_count > 1) && unpaper_b200_engine_set_output_format(e, mono_pages ? -1 : e->page_fmt) != 0) {
    unpaper_b200_engine_destroy(e); return NULL;
  }
  return e;
}

static void mark(B200Engine *e, Lane *ln, int stage_boundary) {
  if (!e->profiling) return;
  CUDA_OK(cudaEventRecord(ln->fl[ln->slot].ev[stage_boundary], ln->st));
  ln->fl[ln->slot].ev_mask |= 1 << stage_boundary;
}

/* apply pre-masks first (sheet_stages.c:211-214), then wipes (:282-285), then
 * border (:288-291) — same order for the mid and post slots */
/* part 0: everything; 1: only the pre-masks (they come before the pre-stage stretch, :210-214); 2: the rest */
static void run_static(B200Engine *e, Lane *ln, StageCtx *c, int slot, int n, unsigned skip, int part) {
  int P = e->group;
  int q = 0;
  bool has_masks = slot == 0 && e->cfg.pre_mask_count > 0;
  if (part == 1 && !has_masks) return;
  /* the tables were built for buffer 0; aim them at the buffer that holds the sheet now */
  if (ln->static_fill_n[slot] > 0 || e->n_static_mask_jobs[slot] > 0) {
    b200k_retarget_jobs(c->st, c->pages, n, ln->static_fill[slot], ln->static_fill_n[slot] * P,
                        ln->static_mask[slot], e->n_static_mask_jobs[slot] * P, P);
    c->launches++;
  }
  if (has_masks) { if (part != 2) { b200k_apply_masks(c->st, ln->static_mask[slot] + (size_t)q * P, n, c->w, c->h); c->launches++; } q++; }
  if (part == 1) return;
  if (!(skip >> SW_WIPE & 1))
    for (int k = 0; k < ln->static_fill_n[slot]; k++) { b200k_fill_jobs(c->st, ln->static_fill[slot] + (size_t)k * P, n, c->w, c->h); c->launches++; }
  if (q < e->n_static_mask_jobs[slot] && !(skip >> SW_BORDER & 1)) { b200k_apply_masks(c->st, ln->static_mask[slot] + (size_t)q * P, n, c->w, c->h); c->launches++; }
}

/* isExcluded(sheet_nr, no_<stage>_multi_index, ignore_multi_index) (parse.h:30-34) for every stage switch */
static unsigned sheet_skip(const B200Engine *e, int sheet_index) {
  int nr = e->first_sheet_nr + sheet_index;
  unsigned m = 0;
  for (int i = 0; i < e->sw_n[SW_COUNT]; i++) if (e->sw_idx[SW_COUNT][i] == nr) return (1u << SW_COUNT) - 1;
  for (int k = 0; k < SW_COUNT; k++)
    for (int i = 0; i < e->sw_n[k]; i++) if (e->sw_idx[k][i] == nr) { m |= 1u << k; break; }
  return m;
}
void unpaper_b200_engine_set_first_sheet_nr(B200Engine *e, int sheet_nr) { if (e) e->first_sheet_nr = sheet_nr; }

/* mirror() then shift_image() on every sheet of the group (sheet_stages.c:200-208, :499-508) */
static void run_geometry(B200Engine *e, Lane *ln, StageCtx *c, int k, int n) {
  Direction m = k == 0 ? e->cfg.pre_mirror : e->cfg.post_mirror;
  if (m.horizontal || m.vertical) { b200k_mirror_pages(c->st, c->pages, n, c->w, c->h, m.horizontal, m.vertical); c->launches++; }
  Delta d = k == 0 ? e->cfg.pre_shift : e->cfg.post_shift;
  (void)ln;
  if (d.horizontal != 0 || d.vertical != 0) stage_shift_pass(c, d);   /* shift_image (blit.c:360-368) */
}

/* the size-changing passes of the pre / post stage: img -> other with a new geometry */
static void run_geo_steps(B200Engine *e, StageCtx *c, const GeoStep *steps, int ns, int n) {
  for (int i = 0; i < ns; i++) {
    const GeoStep *g = &steps[i];
    if (g->op == GS_ROT90) b200k_rotate90_pages(c->st, c->pages, n, c->w, c->h, g->dir, g->to.pitch);
    else if (g->op == GS_STRETCH) b200k_stretch_pages(c->st, c->pages, n, c->w, c->h, g->to.w, g->to.h, g->to.pitch, e->cfg.interpolate_type);
    else b200k_center_pages(c->st, c->pages, n, g->to.w, g->to.h, g->to.pitch);
    b200k_swap_sheets(c->st, c->pages, n);
    b200k_set_geometry(c->st, c->pages, n, g->to.w, g->to.h, g->to.pitch);
    c->parity ^= 1; c->w = g->to.w; c->h = g->to.h;
    c->rows_aligned16 = 1;
    c->launches += 3;
  }
}

static void issue_group(B200Engine *e, Lane *ln, const uint8_t *pages_dev_in, int n, unsigned skip) {
  const B200SheetConfig *cfg = &e->cfg;
#define OFF(k, flag) (cfg->flag || (skip >> (k) & 1))
  const bool no_blackfilter = OFF(SW_BLACK, no_blackfilter), no_noisefilter = OFF(SW_NOISE, no_noisefilter);
  const bool no_blurfilter = OFF(SW_BLUR, no_blurfilter), no_grayfilter = OFF(SW_GRAY, no_grayfilter);
  const bool no_mask_scan = OFF(SW_MASK_SCAN, no_mask_scan), no_mask_center = OFF(SW_MASK_CENTER, no_mask_center);
  const bool no_deskew = OFF(SW_DESKEW, no_deskew), no_border_scan = OFF(SW_BORDER_SCAN, no_border_scan);
  const bool no_border_align = OFF(SW_BORDER_ALIGN, no_border_align);
#undef OFF
  StageCtx c;
  bool direct_out = false;
  memset(&c, 0, sizeof(c));
  c.st = ln->st; c.npages = n; c.pages = ln->pages_dev; c.w = e->g_in.w; c.h = e->g_in.h; c.fmt = e->dfmt;
  c.rows_aligned16 = (e->sheet_pitch & 15) == 0;   /* slabs are 256-byte aligned, strides multiples of 256 */
  c.fillA = ln->fillA; c.fillB = ln->fillB; c.fillC = ln->fillC; c.copyA = ln->copyA; c.copyB = ln->copyB; c.maskJ = ln->maskJ;
  if (ln->rot_pull) { c.rot_jobs = ln->rot_jobs[ln->slot]; c.rot_pull = ln->rot_pull; c.rot_tab_host = ln->rot_tab_host; c.rot_tab_dev = ln->rot_tab_dev; }
  ln->fl[ln->slot].ev_mask = 0;
  int ic = cfg->input_count;

  mark(e, ln, STG_DECODE);
  b200k_page_reset(c.st, c.pages, n);
  if (e->n_pre || e->n_post) { b200k_set_geometry(c.st, c.pages, n, e->g_in.w, e->g_in.h, e->g_in.pitch); c.launches++; }
  /* decode stage: page(s) -> sheet (pages_dev_in == NULL: the upload already placed them) */
  if (pages_dev_in) {
    const uint8_t *psrc = pages_dev_in;
    size_t pbytes = e->page_bytes;
    if (e->pre_rot_dir) {
      /* flip_rotate_90 of every input page (sheet_stages.c:134-137) */
      DImg rs = {(uint8_t *)pages_dev_in, e->page_w, e->page_h, e->page_row, e->page_dfmt, cfg->abs_black_threshold, {255, 255, 255}};
      DImg rd = {ln->page_stage2, e->rp_w, e->rp_h, e->rp_row, e->page_dfmt, cfg->abs_black_threshold, {255, 255, 255}};
      b200k_rotate90_batch(c.st, rs, rd, e->pre_rot_dir, n * ic, e->page_bytes, e->rp_bytes);
      psrc = ln->page_stage2; pbytes = e->rp_bytes;
      c.launches += 1;
    }
    for (int p = 0; p < n; p++)
      for (int j = 0; j < ic; j++)
        ln->decode_copy_host_tmpl[(size_t)p * ic + j].src.data = (uint8_t *)psrc + pbytes * ((size_t)p * ic + j);
    /* job records are tiny; one pageable H2D per group */
    CUDA_OK(cudaMemcpyAsync(ln->decode_copy, ln->decode_copy_host_tmpl, sizeof(DCopyJob) * n * ic, cudaMemcpyHostToDevice, c.st));
    b200k_fill_jobs(c.st, ln->decode_fill, n, e->g_in.w, e->g_in.h);   /* jobs disabled unless the pages leave part of the sheet uncovered */
    b200k_copy_jobs(c.st, ln->decode_copy, n * ic, e->rp_row, e->rp_h);
    c.launches += 2;
  }
  c.launches += 1;
  run_geometry(e, ln, &c, 0, n);
  run_static(e, ln, &c, 0, n, skip, 1);                       /* pre-masks */
  run_geo_steps(e, &c, e->pre_steps, e->n_pre, n);          /* stretch, resize */
  run_static(e, ln, &c, 0, n, skip, 2);                       /* pre-wipes, pre-border */

  mark(e, ln, STG_BLACK);
  if (!no_blackfilter) stage_blackfilter(&c, &e->bf);
  mark(e, ln, STG_NOISE);
  if (!no_noisefilter && stage_noisefilter(&c, cfg->noisefilter_intensity, cfg->abs_white_threshold)) e->issue_failed++;
  mark(e, ln, STG_BLUR);
  if (!no_blurfilter) stage_blurfilter(&c, &e->blur);
  mark(e, ln, STG_GRAY);
  /* masks stage: the reference's first detect_masks() result is discarded
   * (sheet_stages.c:368-372) and has no side effect -> not run */
  if (!no_grayfilter && stage_grayfilter(&c, &e->gray)) e->issue_failed++;
  mark(e, ln, STG_MASKS);
  int nm = e->npoints;
  if (!no_deskew) {
    c.want_ink = cfg->interpolate_type == 2;      /* the cubic rotation skips white tiles through the ink map */
    if (!no_mask_scan) stage_detect_masks(&c, &e->mask);
    c.want_ink = 0;
    mark(e, ln, STG_ROTDET);
    /* mask by mask like sheet_stages.c:406-413: detect_rotation(mask i+1) sees the
     * sheet with mask i already deskewed (the masks may share pixels) */
    for (int mi = 0; mi < nm; mi++) {
      if (stage_detect_rotation_mask(&c, &e->rot, mi)) e->issue_failed++;
      if (mi == nm - 1) mark(e, ln, STG_DESKEW);   /* profile split: exact for one mask */
      stage_deskew_mask_pass(&c, cfg->interpolate_type, mi);
    }
    if (nm == 0) mark(e, ln, STG_DESKEW);
  } else { mark(e, ln, STG_ROTDET); mark(e, ln, STG_DESKEW); }
  mark(e, ln, STG_CENTER);
  if (!no_mask_center) {
    if (!no_mask_scan) stage_detect_masks(&c, &e->mask);
    stage_center_masks_pass(&c, nm);
  }
  run_static(e, ln, &c, 1, n, skip, 0);
  mark(e, ln, STG_BORDER);
  if (!no_border_scan) {
    stage_detect_border(&c, &e->border);
    /* apply_masks(border masks) is fused into the first align_mask sweep */
    if (!no_border_align && e->noutside > 0) {
      /* nothing touches the sheet after this sweep and the caller's device buffer has the sheet's own
       * layout: render the last sweep straight into it */
      direct_out = !ln->host_mode && e->out_fmt < 0 && e->out_count == 1 && e->n_post == 0 &&
                   e->g_out.pitch == e->g_out.w * e->bpp && ln->static_fill_n[2] == 0 && e->n_static_mask_jobs[2] == 0 &&
                   !cfg->post_mirror.horizontal && !cfg->post_mirror.vertical && cfg->post_shift.horizontal == 0 &&
                   cfg->post_shift.vertical == 0;
      stage_align_masks_pass(&c, &cfg->mask_alignment, e->noutside, cfg->mask_color, direct_out ? ln->out_dev : NULL,
                             (size_t)e->g_out.pitch * e->g_out.h);
    } else stage_apply_border_masks(&c, cfg->mask_color);
  }
  run_static(e, ln, &c, 2, n, skip, 0);
  run_geometry(e, ln, &c, 1, n);
  run_geo_steps(e, &c, e->post_steps, e->n_post, n);         /* rotate, stretch, resize */
  mark(e, ln, STG_OUTPUT);
  /* output stage (sheet_stages.c:536-631): sheet -> caller, tight rows, + the decisions */
  const int osheet_row = e->g_out.w * e->bpp;
  size_t img_bytes = (size_t)e->out_row * e->g_out.h, out_sheet = img_bytes * e->out_count;
  const uint8_t *cur = c.parity ? ln->sheets2 : ln->sheets;   /* the buffer that holds the finished sheets */
  if (e->out_fmt >= 0 || e->out_count > 1) {
    /* per output image: the sheet split of :606-621 (copy_rectangle of the j-th
     * sheet_w/output_count columns) and saveImage()'s conversion, straight into tight rows */
    uint8_t *dst = ln->host_mode ? ln->out_stage : ln->out_dev;
    for (int j = 0; j < e->out_count; j++) {
      DImg sv;
      memset(&sv, 0, sizeof(sv));
      sv.data = (uint8_t *)cur + (size_t)j * e->out_w * e->bpp; sv.w = e->out_w; sv.h = e->g_out.h; sv.pitch = e->g_out.pitch;
      sv.fmt = e->dfmt; sv.abt = cfg->abs_black_threshold;
      if (e->out_fmt >= 0) {
        DImg dv = sv;
        dv.data = dst + img_bytes * j; dv.pitch = e->out_row; dv.fmt = e->out_dfmt;
        b200k_convert_out(c.st, sv, dv, n, e->sheet_stride, out_sheet);
      } else {
        b200k_pack_rows(c.st, sv.data, e->g_out.pitch, dst + img_bytes * j, e->out_row, e->out_row, e->g_out.h, n,
                        e->sheet_stride, out_sheet);
      }
      c.launches++;
    }
    if (ln->host_mode)
      CUDA_OK(cudaMemcpyAsync(ln->out_host, ln->out_stage, out_sheet * n, cudaMemcpyDeviceToHost, c.st));
  } else if (ln->host_mode) {
    size_t sheet_bytes = (size_t)osheet_row * e->g_out.h;
    if (e->g_out.pitch == osheet_row) {
      /* rows are tight, sheets are `sheet_stride` apart: one 2-D copy with one "row" per sheet */
      CUDA_OK(cudaMemcpy2DAsync(ln->out_host, sheet_bytes, cur, e->sheet_stride, sheet_bytes, (size_t)n,
                                cudaMemcpyDeviceToHost, c.st));
    } else {
      for (int p = 0; p < n; p++)
        CUDA_OK(cudaMemcpy2DAsync(ln->out_host + sheet_bytes * p, (size_t)osheet_row,
                                  cur + e->sheet_stride * p, (size_t)e->g_out.pitch, (size_t)osheet_row,
                                  (size_t)e->g_out.h, cudaMemcpyDeviceToHost, c.st));
    }
  } else if (!direct_out) {
    b200k_pack_rows(c.st, cur, e->g_out.pitch, ln->out_dev, osheet_row, osheet_row, e->g_out.h, n,
                    e->sheet_stride, (size_t)osheet_row * e->g_out.h);
    c.launches++;
  }
  CUDA_OK(cudaMemcpyAsync(ln->fl[ln->slot].pages_res, ln->pages_dev, sizeof(DPage) * n, cudaMemcpyDeviceToHost, c.st));
  mark(e, ln, STG_COUNT);
  CUDA_OK(cudaEventRecord(ln->fl[ln->slot].done_t, c.st));
  CUDA_OK(cudaEventRecord(ln->fl[ln->slot].done, c.st));
  ln->ran = 1;
  ln->last_done_t = ln->fl[ln->slot].done_t;
  e->launches += c.launches;
}

static void collect(B200Engine *e, Lane *ln) {
  if (!ln->fl[ln->slot].busy) return;
  B200SheetResult *results = ln->fl[ln->slot].res;
  CUDA_OK(cudaEventSynchronize(ln->fl[ln->slot].done));
  if (e->profiling) {
    for (int s = 0; s < STG_COUNT; s++) {
      float ms = 0;
      if ((ln->fl[ln->slot].ev_mask >> s & 1) && (ln->fl[ln->slot].ev_mask >> (s + 1) & 1) &&
          cudaEventElapsedTime(&ms, ln->fl[ln->slot].ev[s], ln->fl[ln->slot].ev[s + 1]) == cudaSuccess) {
        if (e->stage_groups[s] == 0 || ms < e->stage_ms_min[s]) e->stage_ms_min[s] = ms;
        if (e->stage_groups[s] == 0 || ms > e->stage_ms_max[s]) e->stage_ms_max[s] = ms;
        e->stage_ms[s] += ms; e->stage_groups[s] += 1;
      }
    }
  }
  for (int p = 0; p < ln->fl[ln->slot].n; p++)
    if (ln->fl[ln->slot].pages_res[p].error) { e->bad_sheets++; e->bad_flags |= ln->fl[ln->slot].pages_res[p].error; e->bad_first = ln->fl[ln->slot].first + p; }
  if (results || e->done_fn) {
    for (int p = 0; p < ln->fl[ln->slot].n; p++) {
      const DPage *pg = &ln->fl[ln->slot].pages_res[p];
      B200SheetResult local;
      B200SheetResult *r = results ? &results[p] : &local;
      memset(r, 0, sizeof(*r));
      r->status = pg->error ? -(int)pg->error : 0;
      r->sheet_width = e->g_out.w; r->sheet_height = e->g_out.h;
      int nm = pg->mask_count < B200_TRACE_MAX_MASKS ? pg->mask_count : B200_TRACE_MAX_MASKS;
      /* the deskew-stage masks are overwritten by the post-stage detection in
       * the device record; rotation[] belongs to the former, masks[] to the latter */
      unsigned sk = ln->fl[ln->slot].skip;
      bool no_scan = e->cfg.no_mask_scan || (sk >> SW_MASK_SCAN & 1);
      r->center_mask_count = e->cfg.no_mask_center || (sk >> SW_MASK_CENTER & 1) || no_scan ? 0 : pg->mask_count;
      r->deskew_mask_count = e->cfg.no_deskew || (sk >> SW_DESKEW & 1) || no_scan ? 0 : pg->mask_count_deskew;
      for (int i = 0; i < nm; i++) {
        r->center_masks[i] = (Rectangle){{{pg->masks[i].x0, pg->masks[i].y0}, {pg->masks[i].x1, pg->masks[i].y1}}};
        r->centered[i] = pg->centered[i];
      }
      for (int i = 0; i < r->deskew_mask_count && i < B200_TRACE_MAX_MASKS; i++) {
        r->deskew_masks[i] = (Rectangle){{{pg->masks_deskew[i].x0, pg->masks_deskew[i].y0}, {pg->masks_deskew[i].x1, pg->masks_deskew[i].y1}}};
        r->rotation[i] = pg->rotation[i];
      }
      r->border_count = e->cfg.no_border_scan || (sk >> SW_BORDER_SCAN & 1) ? 0 : pg->outside_count;
      for (int i = 0; i < r->border_count && i < MAX_PAGES; i++) {
        r->borders[i] = (Border){pg->border[i].left, pg->border[i].top, pg->border[i].right, pg->border[i].bottom};
        r->border_masks[i] = (Rectangle){{{pg->border_mask[i].x0, pg->border_mask[i].y0}, {pg->border_mask[i].x1, pg->border_mask[i].y1}}};
      }
      r->blackfilter_fills = (int32_t)pg->bf_fills;
      r->noise_clusters = (int32_t)pg->nf_clusters;
      /* the sheet is complete (in host mode: downloaded): hand it to the caller while
       * later groups are still in flight (reference: post_process_fn, batch_worker.c:153-158) */
      if (e->done_fn) {
        const uint8_t *sheet = ln->fl[ln->slot].out + unpaper_b200_engine_sheet_bytes(e) * (size_t)p;
        if (e->done_fn(e->done_user, ln->fl[ln->slot].first + p, sheet, r) != 0) e->done_failed++;
      }
    }
  }
  ln->fl[ln->slot].busy = 0;
}

/* ---- streams: begin, feed any number of batches, end ------------------------------
 * Groups are issued round-robin over the lanes, two flights per lane, and collected in
 * issue order; a process_* call is one stream fed once. */

static void collect_next(B200Engine *e) {
  int k = e->g_done++;
  Lane *ln = &e->lanes[k % e->nlanes];
  ln->slot = (k / e->nlanes) & 1;
  static int trace = -1;
  if (trace < 0) trace = getenv("UNPAPER_B200_ENGINE_TRACE") != NULL;
  if (trace) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    CUDA_OK(cudaEventSynchronize(ln->fl[ln->slot].done));
    clock_gettime(CLOCK_MONOTONIC, &b);
    int ready = 0;
    for (int j = e->g_done; j < e->g; j++) {
      Lane *o = &e->lanes[j % e->nlanes];
      if (cudaEventQuery(o->fl[(j / e->nlanes) & 1].done) == cudaSuccess) ready |= 1 << (j - e->g_done);
    }
    fprintf(stderr, "[engine] group %d lane %d waited %.2f ms; later groups already done: 0x%x\n", k, k % e->nlanes,
            (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) / 1e6, ready);
  }
  collect(e, ln);
}

int unpaper_b200_engine_stream_begin(B200Engine *e, int host_mode) {
  if (!e) { b200_set_error("engine: bad arguments"); return -1; }
  if (e->streaming) { b200_set_error("engine: a stream is already open"); return -1; }
  unpaper_b200_set_device(e->device);
  for (int i = 0; i < e->nlanes; i++) e->lanes[i].ran = 0;
  /* every lane is idle here, so an event on lane 0 marks the start of device work */
  if (!e->ev_begin) CUDA_OK(cudaEventCreate(&e->ev_begin));
  CUDA_OK(cudaEventRecord(e->ev_begin, e->lanes[0].st));
  e->streaming = 1; e->host_mode = host_mode; e->g = 0; e->g_done = 0; e->fed = 0;
  return 0;
}

/* `total`: sheets of the whole stream if known (a process_* call), else -1 */
static int stream_feed(B200Engine *e, const uint8_t *pages, uint8_t *out, int n_sheets, B200SheetResult *results, int total, int first_index) {
  if (!e || !e->streaming || (n_sheets > 0 && (!pages || !out)) || n_sheets < 0) { b200_set_error("engine: bad arguments"); return -1; }
  unpaper_b200_set_device(e->device);
  int P = e->group, ic = e->cfg.input_count, host_mode = e->host_mode;
  size_t out_sheet = unpaper_b200_engine_sheet_bytes(e);
  /* Through host buffers the first groups only start computing once their upload is
   * done and the last download runs after everything else: ramp the group size up at
   * the start and (when the end is known) down at the end — quarter, half, full. */
  int q = P / 4 > 0 ? P / 4 : 1, hf = P / 2 > 0 ? P / 2 : 1;
  int ramp_n = e->nlanes * (q + hf);          /* sheets in each ramp */
  int ramp_up = host_mode && (total < 0 || total >= 4 * ramp_n);
  int ramp_down = host_mode && total >= 4 * ramp_n;
  for (int first = 0, n = 0; first < n_sheets; first += n) {
    while (e->g - e->g_done >= 2 * e->nlanes) collect_next(e);   /* the flight this group reuses */
    Lane *ln = &e->lanes[e->g % e->nlanes];
    ln->slot = (e->g / e->nlanes) & 1;
    int pos = e->fed + first;                 /* position in the stream */
    int left = n_sheets - first;
    n = P;
    if (ramp_up && pos < e->nlanes * q) n = q;
    else if (ramp_up && pos < ramp_n) n = hf;
    else if (ramp_down) {
      int tleft = total - pos;
      if (tleft <= e->nlanes * q) n = q;
      else if (tleft <= ramp_n) n = hf;
      else if (tleft - P < ramp_n) n = tleft - ramp_n;   /* the last full-size group ends where the down-ramp begins */
    }
    if (n > left) n = left;
    if (n > P) n = P;
    if (n < 1) n = 1;
    /* a group shares its kernel sequence: cut it where the per-sheet stage switches change */
    int idx = first_index >= 0 ? first_index + first : pos;   /* job index of the group's first sheet */
    unsigned skip = sheet_skip(e, idx);
    for (int m = 1; m < n; m++) if (sheet_skip(e, idx + m) != skip) { n = m; break; }
    struct Flight *fl = &ln->fl[ln->slot];
    fl->first = idx; fl->n = n; fl->skip = skip; fl->res = results ? results + first : NULL;
    ln->host_mode = host_mode;
    const uint8_t *src = pages + e->page_bytes * ic * (size_t)first;
    fl->out = out + out_sheet * first;
    if (host_mode) {
      ln->out_host = out + out_sheet * first;
      if (e->direct_upload) {
        /* page == sheet geometry: the upload IS the decode stage's centre copy
         * (one 2-D copy, one "row" per sheet slot) */
        CUDA_OK(cudaMemcpy2DAsync(ln->sheets, e->sheet_stride, src, e->page_bytes, e->page_bytes, (size_t)n,
                                  cudaMemcpyHostToDevice, ln->st));
        issue_group(e, ln, NULL, n, skip);
      } else {
        CUDA_OK(cudaMemcpyAsync(ln->page_stage, src, e->page_bytes * ic * (size_t)n, cudaMemcpyHostToDevice, ln->st));
        issue_group(e, ln, ln->page_stage, n, skip);
      }
    } else {
      ln->out_dev = out + out_sheet * first;
      issue_group(e, ln, src, n, skip);
    }
    fl->busy = 1;
    e->g++;
  }
  e->fed += n_sheets;
  return 0;
}

int unpaper_b200_engine_stream_feed(B200Engine *e, const uint8_t *pages, uint8_t *out, int n_sheets, B200SheetResult *results,
                                    int first_index) {
  return stream_feed(e, pages, out, n_sheets, results, -1, first_index);
}

/* collect the oldest group still in flight (blocks until it is done); 0 when nothing is in flight */
int unpaper_b200_engine_stream_poll(B200Engine *e) {
  if (!e || !e->streaming) return 0;
  unpaper_b200_set_device(e->device);
  if (e->g_done >= e->g) return 0;
  collect_next(e);
  return 1;
}
int unpaper_b200_engine_stream_in_flight(const B200Engine *e) { return e && e->streaming ? e->g - e->g_done : 0; }

int unpaper_b200_engine_stream_end(B200Engine *e) {
  if (!e || !e->streaming) { b200_set_error("engine: no open stream"); return -1; }
  unpaper_b200_set_device(e->device);
  while (e->g_done < e->g) collect_next(e);   /* drain in issue order */
  e->streaming = 0;
  CUDA_OK(cudaGetLastError());
  e->last_device_ms = 0.0;
  for (int i = 0; i < e->nlanes; i++) {
    float ms = 0;
    if (e->lanes[i].ran && cudaEventElapsedTime(&ms, e->ev_begin, e->lanes[i].last_done_t) == cudaSuccess && ms > e->last_device_ms)
      e->last_device_ms = ms;
  }
  if (e->issue_failed) {
    e->issue_failed = 0; e->done_failed = 0; e->bad_sheets = 0; e->bad_flags = 0;
    b200_set_error("engine: a stage could not be launched: %s", unpaper_b200_last_error());
    return -5;
  }
  if (e->done_failed) {
    int nf = e->done_failed;
    e->done_failed = 0; e->bad_sheets = 0; e->bad_flags = 0;
    b200_set_error("engine: the sheet callback failed for %d sheet(s)", nf);
    return -3;
  }
  int bad = e->bad_sheets;
  e->bad_sheets = 0;
  if (bad) b200_set_error("engine: %d sheet(s) reported device-side failure flags 0x%x (e.g. sheet %d)", bad, e->bad_flags, e->bad_first);
  e->bad_flags = 0;
  return bad ? -2 : 0;
}

static int process(B200Engine *e, const uint8_t *pages, uint8_t *out, int n_sheets, B200SheetResult *results, int host_mode) {
  if (!e || !pages || !out || n_sheets < 0) { b200_set_error("engine: bad arguments"); return -1; }
  if (unpaper_b200_engine_stream_begin(e, host_mode)) return -1;
  if (stream_feed(e, pages, out, n_sheets, results, n_sheets, -1)) { e->streaming = 0; return -1; }
  return unpaper_b200_engine_stream_end(e);
}

int unpaper_b200_engine_process_device(B200Engine *e, const uint8_t *pages_dev, uint8_t *out_dev, int n_sheets,
                                       B200SheetResult *results) {
  return process(e, pages_dev, out_dev, n_sheets, results, 0);
}
int unpaper_b200_engine_process_host(B200Engine *e, const uint8_t *pages_host, uint8_t *out_host, int n_sheets,
                                     B200SheetResult *results) {
  return process(e, pages_host, out_host, n_sheets, results, 1);
}
