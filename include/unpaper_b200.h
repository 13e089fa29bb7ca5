/* unpaper_b200.h — C-ABI of the B200-native `--device=cuda` backend for
 * unpaper's per-sheet processing hot path.
 *
 * Three layers, all `extern "C"`, plain pointers and sizes only:
 *
 *  (1) the reference's own backend boundary: the vtable symbol `backend_cuda`
 *      (reference imageprocess/backend.c:86-88 picks it up through
 *      `extern const ImageBackend backend_cuda`) plus the image residency
 *      protocol of imageprocess/image.h:32-61 and `image_cuda_release`
 *      (imageprocess/image.c:13,52);
 *  (2) host-buffer entry points (`unpaper_b200_host_*`): one vtable op on an
 *      image that lives in caller memory — upload, op, download.  These are
 *      what the parity tests and any non-C host (ctypes, cgo, JNI) bind;
 *  (3) the sheet engine (`unpaper_b200_engine_*`): the reference's
 *      process_sheet() stage order (src/core/sheet_stages.c:660-672) run for a
 *      GROUP of independent sheets per launch with every data-dependent
 *      decision kept on the device — the throughput path bench.py measures.
 *
 * Error convention follows the reference (lib/logging.c:129-141): vtable
 * entry points do not return errors, an unrecoverable CUDA failure prints and
 * exits.  The layer (2)/(3) entry points return 0 on success and a negative
 * code on failure instead, so that a foreign host can recover.
 */
#pragma once

#include "unpaper_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * (1) reference boundary
 * --------------------------------------------------------------------- */

/* replaces reference imageprocess/backend_cuda.c:586-612 */
extern const ImageBackend backend_cuda;

/* replaces reference imageprocess/image_cuda.c (API: imageprocess/image.h:32-61) */
void image_ensure_cuda(Image *image);       /* image_cuda.c:135-206 */
void image_ensure_cuda_alloc(Image *image); /* image_cuda.c:208-267 */
void image_ensure_cpu(Image *image);        /* image_cuda.c:269-305 */
void image_mark_cpu_dirty(Image *image);
void image_mark_cuda_dirty(Image *image);
Image create_image_from_gpu(void *gpu_ptr, size_t pitch, int width, int height,
                            int pixel_format, Pixel background,
                            uint8_t abs_black_threshold, bool owns_memory);
bool image_is_gpu_resident(Image *image);
void image_set_gpu_resident(Image *image, bool resident);
void *image_get_gpu_ptr(Image *image);
size_t image_get_gpu_pitch(Image *image);
void image_cuda_release(Image *image);      /* image.c:13,52 */

/* replaces reference imageprocess/cuda_runtime.h, cuda_stream_pool.h, cuda_mempool.h: the
 * subset the reference's L3/L4 code links against when built with UNPAPER_WITH_CUDA
 * (lib/perf.c:39,63,86; lib/batch_worker.c:198-255; src/pipeline/image_pipeline.c:303-627).
 * Defined in csrc/compat.c over the per-device caches and stream pool of csrc/rt.c. */
#ifdef UNPAPER_B200_WITH_REFERENCE_HEADERS
#include "imageprocess/cuda_mempool.h"
#include "imageprocess/cuda_runtime.h"
#include "imageprocess/cuda_stream_pool.h"
#else
typedef enum {
  UNPAPER_CUDA_INIT_OK = 0,
  UNPAPER_CUDA_INIT_NO_RUNTIME = 1,
  UNPAPER_CUDA_INIT_NO_DEVICE = 2,
  UNPAPER_CUDA_INIT_ERROR = 3,
} UnpaperCudaInitStatus;
typedef struct UnpaperCudaStream UnpaperCudaStream;   /* cuda_runtime.h:15 */
UnpaperCudaInitStatus unpaper_cuda_try_init(void); /* cuda_runtime.c:137-215 */
const char *unpaper_cuda_init_status_string(UnpaperCudaInitStatus st);
/* cuda_runtime.h:62-67,93: streams; the current stream is per thread */
UnpaperCudaStream *unpaper_cuda_stream_create(void);
UnpaperCudaStream *unpaper_cuda_stream_get_default(void);
void unpaper_cuda_stream_destroy(UnpaperCudaStream *stream);
void unpaper_cuda_set_current_stream(UnpaperCudaStream *stream);
UnpaperCudaStream *unpaper_cuda_get_current_stream(void);
void unpaper_cuda_stream_synchronize_on(UnpaperCudaStream *stream);
void unpaper_cuda_stream_synchronize(void);
void *unpaper_cuda_stream_get_raw_handle(UnpaperCudaStream *stream);
/* cuda_runtime.h:27-47,69-71: raw device memory (e.g. a GPU decoder's output buffer, then
 * create_image_from_gpu) */
uint64_t unpaper_cuda_malloc(size_t bytes);
void unpaper_cuda_free(uint64_t dptr);
void unpaper_cuda_memcpy_h2d(uint64_t dst, const void *src, size_t bytes);
void unpaper_cuda_memcpy_d2h(void *dst, uint64_t src, size_t bytes);
void unpaper_cuda_memcpy_d2d(uint64_t dst, uint64_t src, size_t bytes);
void unpaper_cuda_memcpy_h2d_async(UnpaperCudaStream *stream, uint64_t dst, const void *src, size_t bytes);
void unpaper_cuda_memcpy_d2h_async(UnpaperCudaStream *stream, void *dst, uint64_t src, size_t bytes);
void unpaper_cuda_memcpy_d2d_async(UnpaperCudaStream *stream, uint64_t dst, uint64_t src, size_t bytes);
void unpaper_cuda_memset_d8(uint64_t dst, uint8_t value, size_t bytes);
void unpaper_cuda_memset_async(UnpaperCudaStream *stream, uint64_t dst, uint8_t value, size_t bytes);
/* cuda_runtime.h:86-92: event pairs (perf recorder) */
bool unpaper_cuda_events_supported(void);
bool unpaper_cuda_event_pair_start(void **start, void **stop);
double unpaper_cuda_event_pair_stop_ms(void **start, void **stop);
bool unpaper_cuda_events_supported_on(UnpaperCudaStream *stream);
bool unpaper_cuda_event_pair_start_on(UnpaperCudaStream *stream, void **start, void **stop);
double unpaper_cuda_event_pair_stop_ms_on(UnpaperCudaStream *stream, void **start, void **stop);
/* cuda_stream_pool.h:22-89 */
typedef struct CudaStreamPool CudaStreamPool;
typedef struct {
  size_t stream_count, total_acquisitions, waits, peak_in_use, current_in_use;
} CudaStreamPoolStats;
CudaStreamPool *cuda_stream_pool_create(size_t stream_count);
void cuda_stream_pool_destroy(CudaStreamPool *pool);
UnpaperCudaStream *cuda_stream_pool_acquire(CudaStreamPool *pool);
void cuda_stream_pool_release(CudaStreamPool *pool, UnpaperCudaStream *stream);
CudaStreamPoolStats cuda_stream_pool_get_stats(const CudaStreamPool *pool);
void cuda_stream_pool_print_stats(const CudaStreamPool *pool);
bool cuda_stream_pool_global_init(size_t stream_count);
void cuda_stream_pool_global_cleanup(void);
bool cuda_stream_pool_global_active(void);
UnpaperCudaStream *cuda_stream_pool_global_acquire(void);
void cuda_stream_pool_global_release(UnpaperCudaStream *stream);
CudaStreamPoolStats cuda_stream_pool_global_get_stats(void);
void cuda_stream_pool_global_print_stats(void);
/* cuda_mempool.h:19-141 (three global pools: image, integral, scratch) */
typedef struct CudaMemPool CudaMemPool;
typedef struct {
  size_t total_allocations, pool_hits, pool_misses, size_mismatches, pool_exhaustion;
  size_t current_in_use, peak_in_use, total_bytes_pooled, buffer_count, buffer_size;
} CudaMemPoolStats;
CudaMemPool *cuda_mempool_create(size_t buffer_count, size_t buffer_size);
void cuda_mempool_destroy(CudaMemPool *pool);
uint64_t cuda_mempool_acquire(CudaMemPool *pool, size_t bytes);
void cuda_mempool_release(CudaMemPool *pool, uint64_t dptr);
CudaMemPoolStats cuda_mempool_get_stats(const CudaMemPool *pool);
void cuda_mempool_print_stats(const CudaMemPool *pool);
#define B200_DECL_GLOBAL_POOL(NAME)                                        \
  bool cuda_mempool_##NAME##global_init(size_t buffer_count, size_t buffer_size); \
  void cuda_mempool_##NAME##global_cleanup(void);                          \
  bool cuda_mempool_##NAME##global_active(void);                           \
  uint64_t cuda_mempool_##NAME##global_acquire(size_t bytes);              \
  void cuda_mempool_##NAME##global_release(uint64_t dptr);                 \
  CudaMemPoolStats cuda_mempool_##NAME##global_get_stats(void);            \
  void cuda_mempool_##NAME##global_print_stats(void);
B200_DECL_GLOBAL_POOL()
B200_DECL_GLOBAL_POOL(integral_)
B200_DECL_GLOBAL_POOL(scratch_)
#undef B200_DECL_GLOBAL_POOL
#endif
/* Per-thread current device + stream (reference: thread-local current stream,
 * cuda_runtime.c:70,616-626; extended here to {device, stream}). */
int unpaper_b200_set_device(int device);
int unpaper_b200_get_device(void);
int unpaper_b200_device_count(void);
void unpaper_b200_thread_sync(void);

/* ------------------------------------------------------------------------
 * (2) host-buffer entry points
 * --------------------------------------------------------------------- */

/* An image in caller memory.  `format` is an AVPixelFormat value
 * (GRAY8=8, RGB24=2, Y400A=58, MONOWHITE=9, MONOBLACK=10). */
typedef struct {
  uint8_t *data;
  int32_t width;
  int32_t height;
  int32_t linesize;
  int32_t format;
  Pixel background;
  uint8_t abs_black_threshold;
} B200HostImage;

#define B200_FMT_RGB24 2
#define B200_FMT_GRAY8 8
#define B200_FMT_MONOWHITE 9
#define B200_FMT_MONOBLACK 10
#define B200_FMT_Y400A 58

int unpaper_b200_host_wipe_rectangle(B200HostImage *img, const Rectangle *area, Pixel color);
int unpaper_b200_host_copy_rectangle(const B200HostImage *src, B200HostImage *dst,
                                     const Rectangle *src_area, Point target);
int unpaper_b200_host_center_image(const B200HostImage *src, B200HostImage *dst,
                                   Point target_origin, RectangleSize target_size);
/* The four *_and_replace ops allocate their result: `out` must provide a
 * buffer of out->linesize*out->height bytes; width/height are checked. */
int unpaper_b200_host_stretch(const B200HostImage *img, B200HostImage *out, int32_t interp);
int unpaper_b200_host_resize(const B200HostImage *img, B200HostImage *out, int32_t interp);
int unpaper_b200_host_flip_rotate_90(const B200HostImage *img, B200HostImage *out, int32_t direction);
int unpaper_b200_host_mirror(B200HostImage *img, Direction direction);
int unpaper_b200_host_shift(const B200HostImage *img, B200HostImage *out, Delta d);

int unpaper_b200_host_apply_masks(B200HostImage *img, const Rectangle *masks, size_t n, Pixel color);
int unpaper_b200_host_apply_wipes(B200HostImage *img, const Wipes *wipes, Pixel color);
int unpaper_b200_host_apply_border(B200HostImage *img, const Border *border, Pixel color);
/* returns mask count (>=0) or a negative error */
int unpaper_b200_host_detect_masks(const B200HostImage *img, const MaskDetectionParameters *p,
                                   const Point *points, size_t n, Rectangle *masks_out);
int unpaper_b200_host_center_mask(B200HostImage *img, Point center, const Rectangle *area);
int unpaper_b200_host_align_mask(B200HostImage *img, const Rectangle *inside,
                                 const Rectangle *outside, const MaskAlignmentParameters *p);
int unpaper_b200_host_detect_border(const B200HostImage *img, const BorderScanParameters *p,
                                    const Rectangle *outside, Border *out);

int unpaper_b200_host_blackfilter(B200HostImage *img, const BlackfilterParameters *p);
int unpaper_b200_host_blurfilter(B200HostImage *img, const BlurfilterParameters *p,
                                 uint8_t abs_white_threshold);
int unpaper_b200_host_noisefilter(B200HostImage *img, uint64_t intensity, uint8_t min_white_level);
int unpaper_b200_host_grayfilter(B200HostImage *img, const GrayfilterParameters *p);

int unpaper_b200_host_detect_rotation(const B200HostImage *img, const Rectangle *mask,
                                      const DeskewParameters *p, float *radians_out);
int unpaper_b200_host_deskew(B200HostImage *img, const Rectangle *mask, float radians, int32_t interp);

/* Output side (sheet_stage_output -> saveImage, file.c:134-260).
 * unpaper_b200_output_format: the format saveImage() really writes for a
 *   requested one (Y400A -> GRAY8, MONOBLACK -> MONOWHITE; file.c:201-208).
 * unpaper_b200_host_convert_format: saveImage()'s conversion into `out`
 *   (same size; out->format must already be an output format): MONOWHITE
 *   thresholds gray < in->abs_black_threshold (file.c:215-243), MONOBLACK input
 *   inverts bytes (:244-255), otherwise copy_rectangle semantics (:258).
 * unpaper_b200_pnm_header / _write_pnm: saveImageDirect (file.c:134-176). */
int unpaper_b200_output_format(int av_pix_fmt);
int unpaper_b200_host_convert_format(const B200HostImage *in, B200HostImage *out);
int unpaper_b200_pnm_header(int av_pix_fmt, int width, int height, char *buf, size_t cap);
int unpaper_b200_write_pnm(const char *path, const uint8_t *data, int linesize,
                           int width, int height, int av_pix_fmt);

/* ------------------------------------------------------------------------
 * (3) sheet engine
 * --------------------------------------------------------------------- */

/* struct MultiIndex (parse.h:15-18) */
typedef struct {
  int32_t count;            /* -1: all sheets */
  const int32_t *indexes;
} B200MultiIndex;

/* The part of the reference's `Options` + `SheetProcessConfig`
 * (lib/options.h:29-119, sheet_process.h:22-37) that process_sheet() reads on
 * the hot path, flattened to a POD. */
typedef struct {
  int32_t layout;       /* Layout */
  int32_t input_count;  /* pages per sheet on input (1 or 2) */
  int32_t interpolate_type;
  Pixel sheet_background;
  Pixel mask_color;
  uint8_t abs_black_threshold;
  uint8_t abs_white_threshold;
  /* stage switches: options->no_*_multi_index with count -1 ("disable all") */
  uint8_t no_blackfilter, no_noisefilter, no_blurfilter, no_grayfilter;
  uint8_t no_mask_scan, no_mask_center, no_deskew, no_wipe, no_border;
  uint8_t no_border_scan, no_border_align;
  uint8_t reserved0;
  uint64_t noisefilter_intensity;
  BlackfilterParameters blackfilter; /* exclusions_count==0 -> layout default */
  BlurfilterParameters blurfilter;
  GrayfilterParameters grayfilter;
  DeskewParameters deskew;
  MaskDetectionParameters mask_detection;
  MaskAlignmentParameters mask_alignment;
  BorderScanParameters border_scan;
  Border pre_border, border, post_border;
  int32_t middle_wipe[2];
  int32_t point_count;      /* 0 -> layout default points */
  Point points[8];
  int32_t pre_mask_count;
  Rectangle pre_masks[8];
  int32_t pre_wipe_count, wipe_count, post_wipe_count;
  Rectangle pre_wipes[8], wipes[8], post_wipes[8];
  /* size-preserving geometry options: options->pre_mirror / pre_shift before the
   * pre-masks (sheet_stages.c:200-208), post_mirror / post_shift after the post
   * border (:499-508) */
  Direction pre_mirror, post_mirror;
  Delta pre_shift, post_shift;
  /* sheet_stage_output (sheet_stages.c:606-621): 0 or 1 = one output image per sheet;
   * 2 = the sheet is split into two pages of width sheet_width/2 (--output-pages 2),
   * each converted to the output format on its own like saveImage() does */
  int32_t output_count;
  /* job->sheet_nr of the first sheet of a process_* call (1-based; 0 means 1): the
   * number the per-sheet switches below are tested against */
  int32_t first_sheet_nr;
  /* options->no_*_multi_index / ignore_multi_index (parse.h:15-34, tested with
   * isExcluded() at sheet_stages.c:282-493 and :644-650): sheets for which a stage is
   * skipped.  count -1 = every sheet (same as the uint8 switch above), 0 = none.
   * The index arrays are read during engine_create only. */
  B200MultiIndex no_blackfilter_sheets, no_noisefilter_sheets, no_blurfilter_sheets, no_grayfilter_sheets;
  B200MultiIndex no_mask_scan_sheets, no_mask_center_sheets, no_deskew_sheets, no_wipe_sheets, no_border_sheets;
  B200MultiIndex no_border_scan_sheets, no_border_align_sheets, ignore_sheets;
  /* size-changing options of the decode / pre / post stages (sheet_stages.c:134-145,
   * :216-230, :511-531).  Rotations in degrees (0, 90, -90); sizes -1,-1 = unset;
   * zoom factors 0 or 1.0 = none. */
  int32_t pre_rotate, post_rotate;
  RectangleSize sheet_size, stretch_size, page_size, post_stretch_size, post_page_size;
  float pre_zoom_factor, post_zoom_factor;
} B200SheetConfig;

#define B200_TRACE_MAX_MASKS 8

/* What process_sheet() decided for one sheet (every integer here must be
 * bit-identical with the reference CPU backend). */
typedef struct {
  int32_t status;        /* 0 ok */
  int32_t sheet_width, sheet_height;
  int32_t deskew_mask_count;                 /* sheet_stages.c:401-404 */
  Rectangle deskew_masks[B200_TRACE_MAX_MASKS];
  float rotation[B200_TRACE_MAX_MASKS];      /* sheet_stages.c:408-409 */
  int32_t center_mask_count;                 /* sheet_stages.c:433-436 */
  Rectangle center_masks[B200_TRACE_MAX_MASKS];
  int32_t centered[B200_TRACE_MAX_MASKS];    /* masks.c:232 took the branch */
  int32_t border_count;                      /* sheet_stages.c:467-473 */
  Border borders[MAX_PAGES];
  Rectangle border_masks[MAX_PAGES];
  int32_t blackfilter_fills;                 /* flood-filled scan areas */
  int32_t noise_clusters;                    /* filters.c:342 count */
  int32_t reserved[6];
} B200SheetResult;

/* Defaults = options_init() + options_init_filter_defaults() + the CLI's
 * threshold defaults (lib/options.c:22-170, src/cli/cli_options.c:229-274,
 * :1108-1109). */
void unpaper_b200_sheet_config_defaults(B200SheetConfig *cfg);

typedef struct B200Engine B200Engine;

/* One engine = one GPU, `lanes` groups in flight (each lane owns a stream, a
 * workspace and pinned staging), `group_pages` sheets per group.  All sheets
 * of an engine share input geometry and configuration. */
B200Engine *unpaper_b200_engine_create(const B200SheetConfig *cfg, int device,
                                       int page_width, int page_height,
                                       int page_format, int group_pages,
                                       int lanes);
void unpaper_b200_engine_destroy(B200Engine *e);
int unpaper_b200_engine_sheet_width(const B200Engine *e);
int unpaper_b200_engine_sheet_height(const B200Engine *e);
/* bytes of one sheet's output (format = page format unless an output format was
 * set): output_count tightly packed images of unpaper_b200_engine_output_width()
 * x sheet_height, one after the other */
size_t unpaper_b200_engine_sheet_bytes(const B200Engine *e);
int unpaper_b200_engine_output_width(const B200Engine *e);
int unpaper_b200_engine_output_count(const B200Engine *e);
/* sheet number (job->sheet_nr) of sheet 0 of the next process_* call, for callers that
 * feed one job list through several calls (the per-sheet switches are tested against it) */
void unpaper_b200_engine_set_first_sheet_nr(B200Engine *e, int sheet_nr);
/* sheet_stage_output's format conversion on the device (so the D2H carries
 * 1 bit/px for pbm output): sheets leave process_* in `av_pix_fmt` (mapped by
 * unpaper_b200_output_format), tight rows; -1 restores the page format.
 * Call between process_* calls. */
int unpaper_b200_engine_set_output_format(B200Engine *e, int av_pix_fmt);
int unpaper_b200_engine_output_format(const B200Engine *e);

/* Pages resident in device memory: `pages_dev` holds n_sheets*input_count
 * tightly packed pages (row stride = width*bpp); `out_dev` receives n_sheets
 * tightly packed output sheets.  Asynchronous across lanes; returns after all
 * work has been issued AND completed. `results` may be NULL. */
int unpaper_b200_engine_process_device(B200Engine *e, const uint8_t *pages_dev,
                                       uint8_t *out_dev, int n_sheets,
                                       B200SheetResult *results);
/* Same through host memory: H2D of every page and D2H of every sheet happen
 * inside the call (pinned staging, overlapped across lanes). */
int unpaper_b200_engine_process_host(B200Engine *e, const uint8_t *pages_host,
                                     uint8_t *out_host, int n_sheets,
                                     B200SheetResult *results);
/* The same as a stream: begin, feed any number of batches as they become available
 * (each call issues its groups and returns; sheets are handed to the callback /
 * `results` as their groups complete, in order), end drains.  This is what a page
 * scheduler feeds from a decoded-page queue (reference lib/batch_worker.c:101-149
 * consuming lib/decode_queue.h).  `pages`/`out` of a feed must stay valid until its
 * sheets have been reported.  `first_index`: job index of the feed's first sheet
 * (callback index, per-sheet switches); -1 = position in the stream.
 * stream_poll collects the oldest group in flight (blocking) and returns 1, or 0 if none. */
int unpaper_b200_engine_stream_begin(B200Engine *e, int host_mode);
int unpaper_b200_engine_stream_feed(B200Engine *e, const uint8_t *pages, uint8_t *out, int n_sheets,
                                    B200SheetResult *results, int first_index);
int unpaper_b200_engine_stream_poll(B200Engine *e);
int unpaper_b200_engine_stream_in_flight(const B200Engine *e);
int unpaper_b200_engine_stream_end(B200Engine *e);
/* Per-sheet completion hook (reference: BatchWorkerPostProcessFn,
 * lib/batch_worker.h:22-25, called at batch_worker.c:153-158 after a successful
 * process_sheet()).  Called on the thread that runs process_*, in sheet order,
 * as soon as the sheet's group has finished — in host mode its bytes are
 * already in `sheet` (caller memory), while later groups are still running, so
 * encoding/writing overlaps with the GPU.  In device mode `sheet` is the device
 * pointer.  A non-zero return marks the call as failed (process_* returns -3
 * after draining).  NULL removes the hook. */
typedef int (*B200SheetDoneFn)(void *user, int sheet_index, const uint8_t *sheet,
                               const B200SheetResult *result);
void unpaper_b200_engine_set_sheet_callback(B200Engine *e, B200SheetDoneFn fn, void *user);
/* Kernel launches issued by the engine since creation (for bench.py). */
uint64_t unpaper_b200_engine_launch_count(const B200Engine *e);
/* Device time of the last process_* call: CUDA events from the first enqueue on
 * the first lane's stream to the last lane's completion, in ms. */
double unpaper_b200_engine_last_device_ms(const B200Engine *e);
/* Device-time (ms) per named kernel family accumulated with CUDA events when
 * profiling is enabled; returns number of entries written. */
int unpaper_b200_engine_set_profiling(B200Engine *e, int enabled);
int unpaper_b200_engine_get_profile(const B200Engine *e, int max_entries,
                                    const char **names, double *ms,
                                    uint64_t *launches, double *alg_bytes);
/* fastest / slowest group of each stage since profiling was switched on (ms) */
int unpaper_b200_engine_get_profile_spread(const B200Engine *e, int max_entries, double *min_ms, double *max_ms);

/* ------------------------------------------------------------------------
 * (4) page scheduler across GPUs (replaces lib/batch_worker.c:174-296 +
 *     lib/decode_queue.h for this path)
 * --------------------------------------------------------------------- */

/* Producer hook, shaped like DecodeQueueCustomDecoder (lib/decode_queue.h:54-56): put
 * the input_count decoded pages of job `sheet_index` (tight rows, the engine's page
 * format) into `dst_pinned`.  Return 0, 1 = no more input (the job list ends before
 * this sheet), < 0 = error.  Called concurrently from one producer thread per device. */
typedef int (*B200PageProducerFn)(void *user, int sheet_index, uint8_t *dst_pinned);
/* Sink: sheet `sheet_index` is finished on `device`; `sheet` (pinned host memory, the
 * engine's output layout) is valid during the call only.  Called from that device's
 * feeder thread, in job order per device.  Non-zero fails the run. */
typedef int (*B200PoolSheetFn)(void *user, int sheet_index, int device, const uint8_t *sheet,
                               const B200SheetResult *result);
typedef struct B200Pool B200Pool;

/* One engine, one bounded ring of pinned slots (the decoded-page queue) and two
 * threads (producer, feeder) per device; job indices come from one shared counter.
 * devices == NULL: devices 0..n_devices-1.  slot_sheets <= 0: group_pages;
 * slots_per_device is raised to what the lanes need to stay full. */
B200Pool *unpaper_b200_pool_create(const B200SheetConfig *cfg, const int *devices, int n_devices,
                                   int page_width, int page_height, int page_format,
                                   int group_pages, int lanes, int slot_sheets, int slots_per_device);
void unpaper_b200_pool_destroy(B200Pool *p);
/* Process jobs 0..n_sheets-1; returns when all are done.  `results` may be NULL. */
int unpaper_b200_pool_run(B200Pool *p, int n_sheets, B200PageProducerFn produce, void *produce_user,
                          B200PoolSheetFn sink, void *sink_user, B200SheetResult *results);
int unpaper_b200_pool_device_count(const B200Pool *p);
size_t unpaper_b200_pool_sheet_bytes(const B200Pool *p);
uint64_t unpaper_b200_pool_sheets_done(const B200Pool *p, int device_index);   /* of the last run */
B200Engine *unpaper_b200_pool_engine(B200Pool *p, int device_index);           /* e.g. to set the output format */

const char *unpaper_b200_last_error(void);
const char *unpaper_b200_version(void);

#ifdef __cplusplus
}
#endif
