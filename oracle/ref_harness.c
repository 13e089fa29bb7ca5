/* ref_harness.c — TEST INFRASTRUCTURE (never linked into the product).
 *
 * Thin C-ABI over the UNMODIFIED reference CPU backend, compiled by
 * oracle/Makefile from the sources where they lie under /root/reference into
 * oracle/_ref/libunpaper_ref.so.  It exists so that tests/ and bench.py's
 * cpu_baseline / --impl reference legs can drive the reference's own
 * `*_cpu` functions and its own process_sheet() on in-memory pages.
 *
 * Every entry point mirrors the signature of the matching
 * `unpaper_b200_host_*` / engine call in include/unpaper_b200.h with the
 * prefix `ref_`, so one ctypes binding serves both.
 *
 * Tracing: reference imageprocess/backend.c is compiled with
 * UNPAPER_WITH_CUDA=1 so that image_backend_select(UNPAPER_DEVICE_CUDA) binds
 * the symbol `backend_cuda` — provided HERE as a pass-through that forwards
 * every op to the reference's `*_cpu` function and records what the detectors
 * returned (masks, rotation, border).  No arithmetic is changed.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <libavutil/frame.h>

#include "imageprocess/backend.h"
#include "imageprocess/blit.h"
#include "imageprocess/image.h"
#include "imageprocess/masks.h"
#include "imageprocess/pixel.h"
#include "lib/logging.h"
#include "lib/options.h"
#include "parse.h"
#include "sheet_process.h"

#define UNPAPER_B200_WITH_REFERENCE_HEADERS 1
#include "unpaper_b200.h"

/* ---- symbols the reference expects from files we do not compile ---------- */

/* file.c (loadImage / saveImage / detectPixelFormatFromExtension) is compiled from the
 * reference against oracle/shim_codec; its FFmpeg calls land here.  saveImage() converts
 * and hands GRAY8 / RGB24 / MONOWHITE to saveImageDirect() (file.c:134-176), which covers
 * every format saveImage() can produce, so none of these is reached by the tests. */
#include <libavformat/avformat.h>
#include <libavutil/opt.h>
const AVCodec *avcodec_find_decoder(enum AVCodecID id) { (void)id; return NULL; }
const AVCodec *avcodec_find_encoder(enum AVCodecID id) { (void)id; return NULL; }
AVCodecContext *avcodec_alloc_context3(const AVCodec *c) { (void)c; return NULL; }
int avcodec_parameters_to_context(AVCodecContext *c, const AVCodecParameters *p) { (void)c; (void)p; return -1; }
int avcodec_open2(AVCodecContext *c, const AVCodec *k, void *o) { (void)c; (void)k; (void)o; return -1; }
int avcodec_send_packet(AVCodecContext *c, const AVPacket *p) { (void)c; (void)p; return -1; }
int avcodec_receive_frame(AVCodecContext *c, AVFrame *f) { (void)c; (void)f; return -1; }
int avcodec_send_frame(AVCodecContext *c, const AVFrame *f) { (void)c; (void)f; return -1; }
int avcodec_receive_packet(AVCodecContext *c, AVPacket *p) { (void)c; (void)p; return -1; }
void avcodec_free_context(AVCodecContext **c) { (void)c; }
AVPacket *av_packet_alloc(void) { return NULL; }
void av_packet_free(AVPacket **p) { (void)p; }
int avformat_open_input(AVFormatContext **s, const char *u, void *f, void *o) { (void)s; (void)u; (void)f; (void)o; return -1; }
int avformat_find_stream_info(AVFormatContext *s, void *o) { (void)s; (void)o; return -1; }
void av_dump_format(AVFormatContext *s, int i, const char *u, int o) { (void)s; (void)i; (void)u; (void)o; }
int av_read_frame(AVFormatContext *s, AVPacket *p) { (void)s; (void)p; return -1; }
void avformat_close_input(AVFormatContext **s) { (void)s; }
int avformat_alloc_output_context2(AVFormatContext **c, void *of, const char *n, const char *f) { (void)of; (void)n; (void)f; if (c) *c = NULL; return -1; }
AVStream *avformat_new_stream(AVFormatContext *s, const AVCodec *c) { (void)s; (void)c; return NULL; }
int avio_open(AVIOContext **s, const char *u, int f) { (void)s; (void)u; (void)f; return -1; }
int avformat_write_header(AVFormatContext *s, void *o) { (void)s; (void)o; return -1; }
int av_write_frame(AVFormatContext *s, AVPacket *p) { (void)s; (void)p; return -1; }
int av_write_trailer(AVFormatContext *s) { (void)s; return -1; }
void avformat_free_context(AVFormatContext *s) { (void)s; }
int av_opt_set(void *o, const char *n, const char *v, int f) { (void)o; (void)n; (void)v; (void)f; return -1; }
void saveImage(char *filename, Image image, int fmt);   /* file.c:186 */
struct EncodeQueue;
bool encode_queue_gpu_enabled(struct EncodeQueue *q) { (void)q; return false; }
bool encode_queue_submit_gpu(struct EncodeQueue *q, void *p, size_t pitch, int w,
                             int h, int c, char **files, int n, int fmt, int j) {
  (void)q; (void)p; (void)pitch; (void)w; (void)h; (void)c; (void)files; (void)n; (void)fmt; (void)j;
  return false;
}
bool encode_queue_submit(struct EncodeQueue *q, AVFrame *f, char **files, int n,
                         int fmt, int j, bool pinned) {
  (void)q; (void)files; (void)n; (void)fmt; (void)j; (void)pinned;
  av_frame_free(&f);
  return false;
}

/* ---- the reference's CPU entry points (imageprocess/backend.c:11-45) ------ */

void wipe_rectangle_cpu(Image image, Rectangle input_area, Pixel color);
void copy_rectangle_cpu(Image source, Image target, Rectangle source_area, Point target_coords);
void center_image_cpu(Image source, Image target, Point target_origin, RectangleSize target_size);
void stretch_and_replace_cpu(Image *pImage, RectangleSize size, Interpolation interpolate_type);
void resize_and_replace_cpu(Image *pImage, RectangleSize size, Interpolation interpolate_type);
void flip_rotate_90_cpu(Image *pImage, RotationDirection direction);
void mirror_cpu(Image image, Direction direction);
void shift_image_cpu(Image *pImage, Delta d);
void apply_masks_cpu(Image image, const Rectangle masks[], size_t masks_count, Pixel color);
void apply_wipes_cpu(Image image, Wipes wipes, Pixel color);
void apply_border_cpu(Image image, const Border border, Pixel color);
size_t detect_masks_cpu(Image image, MaskDetectionParameters params, const Point points[],
                        size_t points_count, Rectangle masks[]);
void align_mask_cpu(Image image, const Rectangle inside_area, const Rectangle outside,
                    MaskAlignmentParameters params);
Border detect_border_cpu(Image image, BorderScanParameters params, const Rectangle outside_mask);
void blackfilter_cpu(Image image, BlackfilterParameters params);
void blurfilter_cpu(Image image, BlurfilterParameters params, uint8_t abs_white_threshold);
void noisefilter_cpu(Image image, uint64_t intensity, uint8_t min_white_level);
void grayfilter_cpu(Image image, GrayfilterParameters params);
float detect_rotation_cpu(Image image, Rectangle mask, const DeskewParameters params);
void deskew_cpu(Image source, Rectangle mask, float radians, Interpolation interpolate_type);

/* ---- tracing pass-through backend --------------------------------------- */

typedef struct {
  B200SheetResult *res;
  int detect_masks_calls;
  int no_deskew;   /* the deskew stage (and its detect_masks call) is skipped */
} Trace;
static __thread Trace *tls_trace = NULL;

static size_t tr_detect_masks(Image image, MaskDetectionParameters params,
                              const Point points[], size_t n, Rectangle masks[]) {
  size_t c = detect_masks_cpu(image, params, points, n, masks);
  Trace *t = tls_trace;
  if (t && t->res) {
    t->detect_masks_calls++;
    /* call 1: masks stage (result discarded by the reference, sheet_stages.c:368-372)
     * call 2: deskew stage; call 3: post stage (mask centring). */
    size_t m = c < B200_TRACE_MAX_MASKS ? c : B200_TRACE_MAX_MASKS;
    int deskew_call = t->no_deskew ? -1 : 2, center_call = t->no_deskew ? 2 : 3;
    if (t->detect_masks_calls == deskew_call) {
      t->res->deskew_mask_count = (int32_t)c;
      memcpy(t->res->deskew_masks, masks, m * sizeof(Rectangle));
    } else if (t->detect_masks_calls == center_call) {
      t->res->center_mask_count = (int32_t)c;
      memcpy(t->res->center_masks, masks, m * sizeof(Rectangle));
    }
  }
  return c;
}
static float tr_detect_rotation(Image image, Rectangle mask, const DeskewParameters p) {
  float r = detect_rotation_cpu(image, mask, p);
  Trace *t = tls_trace;
  if (t && t->res) {
    for (int i = 0; i < t->res->deskew_mask_count && i < B200_TRACE_MAX_MASKS; i++) {
      if (memcmp(&t->res->deskew_masks[i], &mask, sizeof(mask)) == 0) {
        t->res->rotation[i] = r;
        break;
      }
    }
  }
  return r;
}
static Border tr_detect_border(Image image, BorderScanParameters p, const Rectangle outside) {
  Border b = detect_border_cpu(image, p, outside);
  Trace *t = tls_trace;
  if (t && t->res && t->res->border_count < MAX_PAGES) {
    int k = t->res->border_count++;
    t->res->borders[k] = b;
    t->res->border_masks[k] = border_to_mask(image, b);
  }
  return b;
}

#ifndef REF_DROPIN
const ImageBackend backend_cuda = {
    .name = "cpu-traced",
    .wipe_rectangle = wipe_rectangle_cpu,
    .copy_rectangle = copy_rectangle_cpu,
    .center_image = center_image_cpu,
    .stretch_and_replace = stretch_and_replace_cpu,
    .resize_and_replace = resize_and_replace_cpu,
    .flip_rotate_90 = flip_rotate_90_cpu,
    .mirror = mirror_cpu,
    .shift_image = shift_image_cpu,
    .apply_masks = apply_masks_cpu,
    .apply_wipes = apply_wipes_cpu,
    .apply_border = apply_border_cpu,
    .detect_masks = tr_detect_masks,
    .align_mask = align_mask_cpu,
    .detect_border = tr_detect_border,
    .blackfilter = blackfilter_cpu,
    .blurfilter = blurfilter_cpu,
    .noisefilter = noisefilter_cpu,
    .grayfilter = grayfilter_cpu,
    .detect_rotation = tr_detect_rotation,
    .deskew = deskew_cpu,
};
#endif

#ifdef REF_DROPIN
/* DROP-IN BUILD (oracle/Makefile target `dropin`): every reference source file is compiled
 * with -DUNPAPER_WITH_CUDA=1 and the shared object is linked against libunpaper_b200.so,
 * which supplies `backend_cuda`, the image residency calls and the cuda_runtime /
 * stream-pool symbols.  ref_select_device() is the reference's own --device switch
 * (image_backend_select, backend.c:79-97); process_sheet() then runs the reference's
 * stage code over whichever backend is selected. */
static int g_device = UNPAPER_DEVICE_CPU;
int ref_select_device(int cuda) {
  verbose = VERBOSE_QUIET;
  g_device = cuda ? UNPAPER_DEVICE_CUDA : UNPAPER_DEVICE_CPU;
  image_backend_select((UnpaperDevice)g_device);
  return 0;
}
const char *ref_backend_name(void) { return image_backend_get()->name; }
static int g_perf = 0;
/* options->perf: the reference's perf recorder then times every stage with
 * unpaper_cuda_event_pair_start/stop_ms (lib/perf.c:39,63,86) */
int ref_set_perf(int on) { g_perf = on; return 0; }
/* what src/pipeline/image_pipeline.c does around a batch: a global stream pool that
 * lib/batch_worker.c:198-255 hands out per job */
#include "imageprocess/cuda_runtime.h"
#include "imageprocess/cuda_stream_pool.h"
int ref_stream_pool(int n) {
  if (n > 0) return cuda_stream_pool_global_init((size_t)n) ? 0 : -1;
  cuda_stream_pool_global_cleanup();
  return 0;
}
int ref_stream_pool_acquisitions(void) { return (int)cuda_stream_pool_global_get_stats().total_acquisitions; }

/* The residency API the reference's GPU decode / encode paths use (image.h:32-61):
 * a device buffer that did not come from image_ensure_cuda() becomes an Image
 * (create_image_from_gpu), is processed in place through the selected backend and read
 * back.  `mode` 0: the buffer is a view (owns_memory=false, freed by the caller);
 * 1: unpaper_cuda_malloc'ed and handed over (owns_memory=true, freed with the image).
 * `pitch` may be any value >= the row bytes.  Ops: detect_masks at the centre, wipe of
 * `wipe`, copy of `wipe`'s area shifted by (100,60).  Returns the mask count, -1 on a protocol error. */
int ref_gpu_image_roundtrip(const B200HostImage *in, int pitch, int mode, const Rectangle *wipe,
                            const MaskDetectionParameters *mp, Rectangle *mask_out, uint8_t *out) {
  int row = av_shim_row_bytes(in->format, in->width);
  size_t bytes = (size_t)pitch * in->height;
  uint8_t *staged = calloc(1, bytes);
  for (int y = 0; y < in->height; y++) memcpy(staged + (size_t)y * pitch, in->data + (size_t)y * in->linesize, (size_t)row);
  uint64_t d = unpaper_cuda_malloc(bytes);
  unpaper_cuda_memcpy_h2d(d, staged, bytes);
  free(staged);
  Image img = create_image_from_gpu((void *)(uintptr_t)d, (size_t)pitch, in->width, in->height, in->format,
                                    in->background, in->abs_black_threshold, mode == 1);
  if (!img.frame) { unpaper_cuda_free(d); return -1; }
  int rc = 0;
  if (!image_is_gpu_resident(&img) || image_get_gpu_ptr(&img) != (void *)(uintptr_t)d ||
      image_get_gpu_pitch(&img) != (size_t)pitch) rc = -1;
  Point c = {in->width / 2, in->height / 2};
  int n = (int)detect_masks(img, *mp, &c, 1, mask_out);
  wipe_rectangle(img, *wipe, (Pixel){10, 20, 30});
  copy_rectangle(img, img, *wipe, (Point){wipe->vertex[0].x + 100, wipe->vertex[0].y + 60});   /* disjoint from the source */
  /* still the caller's buffer: nothing was reallocated or uploaded over it */
  if (image_get_gpu_ptr(&img) != (void *)(uintptr_t)d) rc = -1;
  image_ensure_cpu(&img);
  for (int y = 0; y < in->height; y++) memcpy(out + (size_t)y * row, img.frame->data[0] + (size_t)y * img.frame->linesize[0], (size_t)row);
  image_set_gpu_resident(&img, false);
  if (image_is_gpu_resident(&img) || image_get_gpu_ptr(&img) != NULL) rc = -1;
  image_set_gpu_resident(&img, true);
  if (!image_is_gpu_resident(&img)) rc = -1;
  free_image(&img);
  if (mode == 0) unpaper_cuda_free(d);
  return rc < 0 ? rc : n;
}
static void ensure_init(void) { verbose = VERBOSE_QUIET; }
#else
static pthread_once_t init_once = PTHREAD_ONCE_INIT;
static void do_init(void) {
  verbose = VERBOSE_QUIET;
  image_backend_select(UNPAPER_DEVICE_CUDA); /* = the pass-through above */
}
static void ensure_init(void) { pthread_once(&init_once, do_init); }
#endif

/* ---- host image <-> reference Image ------------------------------------- */

/* Borrow the caller's buffer: no copy, nothing owned by the frame. */
static Image wrap(const B200HostImage *h, AVFrame *storage) {
  memset(storage, 0, sizeof(*storage));
  storage->data[0] = h->data;
  storage->linesize[0] = h->linesize;
  storage->width = h->width;
  storage->height = h->height;
  storage->format = h->format;
  return (Image){.frame = storage, .background = h->background,
                 .abs_black_threshold = h->abs_black_threshold};
}

static Image clone_owned(const B200HostImage *h) {
  Image img = create_image((RectangleSize){h->width, h->height}, h->format, false,
                           h->background, h->abs_black_threshold);
  int row = av_shim_row_bytes(h->format, h->width);
  for (int y = 0; y < h->height; y++)
    memcpy(img.frame->data[0] + (size_t)y * img.frame->linesize[0],
           h->data + (size_t)y * h->linesize, (size_t)row);
  return img;
}

static int export_owned(Image img, B200HostImage *out) {
  if (out->width != img.frame->width || out->height != img.frame->height ||
      out->format != img.frame->format)
    return -2;
  int row = av_shim_row_bytes(out->format, out->width);
  for (int y = 0; y < out->height; y++)
    memcpy(out->data + (size_t)y * out->linesize,
           img.frame->data[0] + (size_t)y * img.frame->linesize[0], (size_t)row);
  return 0;
}

/* ---- per-op entry points -------------------------------------------------- */

int ref_host_wipe_rectangle(B200HostImage *img, const Rectangle *area, Pixel color) {
  ensure_init();
  AVFrame f; wipe_rectangle(wrap(img, &f), *area, color); return 0;
}
int ref_host_copy_rectangle(const B200HostImage *src, B200HostImage *dst,
                            const Rectangle *src_area, Point target) {
  ensure_init();
  AVFrame a, b; copy_rectangle(wrap(src, &a), wrap(dst, &b), *src_area, target); return 0;
}
int ref_host_center_image(const B200HostImage *src, B200HostImage *dst,
                          Point target_origin, RectangleSize target_size) {
  ensure_init();
  AVFrame a, b; center_image(wrap(src, &a), wrap(dst, &b), target_origin, target_size); return 0;
}
int ref_host_stretch(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  ensure_init();
  Image w = clone_owned(img);
  stretch_and_replace(&w, (RectangleSize){out->width, out->height}, (Interpolation)interp);
  int rc = export_owned(w, out); free_image(&w); return rc;
}
int ref_host_resize(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  ensure_init();
  Image w = clone_owned(img);
  resize_and_replace(&w, (RectangleSize){out->width, out->height}, (Interpolation)interp);
  int rc = export_owned(w, out); free_image(&w); return rc;
}
int ref_host_flip_rotate_90(const B200HostImage *img, B200HostImage *out, int32_t direction) {
  ensure_init();
  Image w = clone_owned(img);
  flip_rotate_90(&w, (RotationDirection)direction);
  int rc = export_owned(w, out); free_image(&w); return rc;
}
int ref_host_mirror(B200HostImage *img, Direction direction) {
  ensure_init();
  AVFrame f; mirror(wrap(img, &f), direction); return 0;
}
int ref_host_shift(const B200HostImage *img, B200HostImage *out, Delta d) {
  ensure_init();
  Image w = clone_owned(img);
  shift_image(&w, d);
  int rc = export_owned(w, out); free_image(&w); return rc;
}
int ref_host_apply_masks(B200HostImage *img, const Rectangle *masks, size_t n, Pixel color) {
  ensure_init();
  AVFrame f; apply_masks(wrap(img, &f), masks, n, color); return 0;
}
int ref_host_apply_wipes(B200HostImage *img, const Wipes *wipes, Pixel color) {
  ensure_init();
  AVFrame f; apply_wipes(wrap(img, &f), *wipes, color); return 0;
}
int ref_host_apply_border(B200HostImage *img, const Border *border, Pixel color) {
  ensure_init();
  AVFrame f; apply_border(wrap(img, &f), *border, color); return 0;
}
int ref_host_detect_masks(const B200HostImage *img, const MaskDetectionParameters *p,
                          const Point *points, size_t n, Rectangle *masks_out) {
  ensure_init();
  AVFrame f; return (int)detect_masks(wrap(img, &f), *p, points, n, masks_out);
}
int ref_host_center_mask(B200HostImage *img, Point center, const Rectangle *area) {
  ensure_init();
  AVFrame f; center_mask(wrap(img, &f), center, *area); return 0;
}
int ref_host_align_mask(B200HostImage *img, const Rectangle *inside, const Rectangle *outside,
                        const MaskAlignmentParameters *p) {
  ensure_init();
  AVFrame f; align_mask(wrap(img, &f), *inside, *outside, *p); return 0;
}
int ref_host_detect_border(const B200HostImage *img, const BorderScanParameters *p,
                           const Rectangle *outside, Border *out) {
  ensure_init();
  AVFrame f; *out = detect_border(wrap(img, &f), *p, *outside); return 0;
}
int ref_host_blackfilter(B200HostImage *img, const BlackfilterParameters *p) {
  ensure_init();
  AVFrame f; blackfilter(wrap(img, &f), *p); return 0;
}
int ref_host_blurfilter(B200HostImage *img, const BlurfilterParameters *p, uint8_t abs_white) {
  ensure_init();
  AVFrame f; blurfilter(wrap(img, &f), *p, abs_white); return 0;
}
int ref_host_noisefilter(B200HostImage *img, uint64_t intensity, uint8_t min_white_level) {
  ensure_init();
  AVFrame f; noisefilter(wrap(img, &f), intensity, min_white_level); return 0;
}
int ref_host_grayfilter(B200HostImage *img, const GrayfilterParameters *p) {
  ensure_init();
  AVFrame f; grayfilter(wrap(img, &f), *p); return 0;
}
int ref_host_detect_rotation(const B200HostImage *img, const Rectangle *mask,
                             const DeskewParameters *p, float *radians_out) {
  ensure_init();
  AVFrame f; *radians_out = detect_rotation(wrap(img, &f), *mask, *p); return 0;
}
int ref_host_deskew(B200HostImage *img, const Rectangle *mask, float radians, int32_t interp) {
  ensure_init();
  AVFrame f; deskew(wrap(img, &f), *mask, radians, (Interpolation)interp); return 0;
}

/* ---- whole sheets through the reference's own process_sheet() ------------- */

static const struct MultiIndex MI_NONE = {.count = 0, .indexes = NULL};
static const struct MultiIndex MI_ALL = {.count = -1, .indexes = NULL};

static void options_from_cfg(Options *o, Rectangle *bf_excl, const B200SheetConfig *c) {
  options_init(o);
  options_init_filter_defaults(o, bf_excl);
#ifdef REF_DROPIN
  o->device = (UnpaperDevice)g_device;
#else
  o->device = UNPAPER_DEVICE_CPU;
#endif
  o->write_output = false;
#ifdef REF_DROPIN
  o->perf = g_perf != 0;
#endif
  o->layout = (Layout)c->layout;
  o->input_count = c->input_count;
  o->interpolate_type = (Interpolation)c->interpolate_type;
  o->sheet_background = c->sheet_background;
  o->mask_color = c->mask_color;
  o->abs_black_threshold = c->abs_black_threshold;
  o->abs_white_threshold = c->abs_white_threshold;
#define SW(flag, field) o->field = (c->flag || c->flag##_sheets.count == -1) ? MI_ALL : \
    (struct MultiIndex){.count = c->flag##_sheets.count, .indexes = (int *)c->flag##_sheets.indexes}
  SW(no_blackfilter, no_blackfilter_multi_index);
  SW(no_noisefilter, no_noisefilter_multi_index);
  SW(no_blurfilter, no_blurfilter_multi_index);
  SW(no_grayfilter, no_grayfilter_multi_index);
  SW(no_mask_scan, no_mask_scan_multi_index);
  SW(no_mask_center, no_mask_center_multi_index);
  SW(no_deskew, no_deskew_multi_index);
  SW(no_wipe, no_wipe_multi_index);
  SW(no_border, no_border_multi_index);
  SW(no_border_scan, no_border_scan_multi_index);
  SW(no_border_align, no_border_align_multi_index);
#undef SW
  o->ignore_multi_index = (struct MultiIndex){.count = c->ignore_sheets.count, .indexes = (int *)c->ignore_sheets.indexes};
  o->noisefilter_intensity = c->noisefilter_intensity;
  o->blackfilter_parameters = c->blackfilter;
  o->blackfilter_parameters.exclusions = bf_excl;
  o->blackfilter_parameters.exclusions_count = 0;
  for (size_t i = 0; i < c->blackfilter.exclusions_count && i < MAX_MASKS; i++)
    bf_excl[o->blackfilter_parameters.exclusions_count++] = c->blackfilter.exclusions[i];
  o->blurfilter_parameters = c->blurfilter;
  o->grayfilter_parameters = c->grayfilter;
  o->deskew_parameters = c->deskew;
  o->mask_detection_parameters = c->mask_detection;
  o->mask_alignment_parameters = c->mask_alignment;
  o->border_scan_parameters = c->border_scan;
  o->pre_border = c->pre_border;
  o->border = c->border;
  o->post_border = c->post_border;
  o->pre_wipes.count = (size_t)c->pre_wipe_count;
  o->wipes.count = (size_t)c->wipe_count;
  o->post_wipes.count = (size_t)c->post_wipe_count;
  /* size-changing options (sheet_stages.c:134-145, :216-230, :511-531) */
  o->pre_rotate = c->pre_rotate; o->post_rotate = c->post_rotate;
  o->sheet_size = c->sheet_size; o->stretch_size = c->stretch_size; o->page_size = c->page_size;
  o->post_stretch_size = c->post_stretch_size; o->post_page_size = c->post_page_size;
  if (c->pre_zoom_factor != 0.0f) o->pre_zoom_factor = c->pre_zoom_factor;
  if (c->post_zoom_factor != 0.0f) o->post_zoom_factor = c->post_zoom_factor;
  o->pre_mirror = c->pre_mirror; o->post_mirror = c->post_mirror;
  o->pre_shift = c->pre_shift; o->post_shift = c->post_shift;
  for (int i = 0; i < 8; i++) {
    o->pre_wipes.areas[i] = c->pre_wipes[i];
    o->wipes.areas[i] = c->wipes[i];
    o->post_wipes.areas[i] = c->post_wipes[i];
  }
}

/* One sheet: pages -> reference process_sheet() -> sheet in the page format. */
static int run_one_sheet(const B200SheetConfig *cfg, const Options *opt,
                         const uint8_t *pages, int page_w, int page_h, int page_fmt,
                         uint8_t *out, B200SheetResult *res, char **out_files, int out_count, int sheet_nr,
                         int out_w, int out_h) {
  int row = av_shim_row_bytes(page_fmt, page_w);
  if (row < 0) return -1;
  size_t page_bytes = (size_t)row * page_h;

  SheetProcessConfig spc;
  /* The batch path passes blackfilter_exclude_count = 0 so that layout
   * defaults apply (image_pipeline.c:377-379); user exclusions already sit in
   * opt->blackfilter_parameters. */
  sheet_process_config_init(&spc, opt, cfg->pre_masks, (size_t)cfg->pre_mask_count,
                            cfg->points, (size_t)cfg->point_count, cfg->middle_wipe,
                            NULL, opt->blackfilter_parameters.exclusions_count);

  BatchJob job;
  memset(&job, 0, sizeof(job));
  job.sheet_nr = sheet_nr;
  job.input_count = cfg->input_count;
  job.output_count = out_files ? out_count : 1;
  for (int j = 0; out_files && j < out_count; j++) job.output_files[j] = out_files[j];
  job.layout_override = -1;

  SheetProcessState st;
  sheet_process_state_init(&st, &spc, &job);
  for (int j = 0; j < cfg->input_count; j++) {
    AVFrame *fr = av_frame_alloc();
    fr->width = page_w; fr->height = page_h; fr->format = page_fmt;
    if (av_frame_get_buffer(fr, 8) < 0) return -3;
    for (int y = 0; y < page_h; y++)
      memcpy(fr->data[0] + (size_t)y * fr->linesize[0],
             pages + (size_t)j * page_bytes + (size_t)y * row, (size_t)row);
    sheet_process_state_set_decoded(&st, fr, j);
  }

  B200SheetResult local;
  if (!res) res = &local;
  memset(res, 0, sizeof(*res));
  Trace tr = {.res = res, .detect_masks_calls = 0,
              .no_deskew = isExcluded(sheet_nr, opt->no_deskew_multi_index, opt->ignore_multi_index)};
  tls_trace = &tr;
  bool ok = process_sheet(&st, &spc);
  tls_trace = NULL;
  res->status = ok ? 0 : -1;
  if (ok) {
    res->sheet_width = st.sheet.frame->width;
    res->sheet_height = st.sheet.frame->height;
    if (out && (res->sheet_width != out_w || res->sheet_height != out_h)) {
      ok = false;                        /* the caller's buffer was sized for another geometry */
      res->status = -5;
    } else if (out) {
      /* what saveImage() would hand to the writer (file.c:211-262) */
      B200HostImage o = {.data = out, .width = res->sheet_width, .height = res->sheet_height,
                         .linesize = av_shim_row_bytes(page_fmt, res->sheet_width),
                         .format = page_fmt, .background = st.sheet.background,
                         .abs_black_threshold = st.sheet.abs_black_threshold};
      AVFrame f;
#ifdef REF_DROPIN
      image_ensure_cpu(&st.sheet);   /* what sheet_stage_output does before saving (sheet_stages.c:586) */
#endif
      copy_rectangle_cpu(st.sheet, wrap(&o, &f), full_image(st.sheet), POINT_ORIGIN);
    }
  }
  sheet_process_state_cleanup(&st);
  return ok ? 0 : -1;
}

typedef struct {
  const B200SheetConfig *cfg;
  const Options *opt;
  const uint8_t *pages;
  uint8_t *out;
  B200SheetResult *results;
  const char *out_dir; int out_count;   /* write_output runs: files <dir>/s<sheet>_<page>.pnm */
  int page_w, page_h, page_fmt, n_sheets;
  int out_w, out_h;                      /* size of the buffers behind `out` */
  size_t sheet_in_bytes, sheet_out_bytes;
  atomic_int next;
  atomic_int failed;
} Job;

static void *worker(void *arg) {
  Job *jb = (Job *)arg;
  for (;;) {
    int i = atomic_fetch_add(&jb->next, 1);
    if (i >= jb->n_sheets) break;
#ifdef REF_DROPIN
    /* lib/batch_worker.c:198-203: one pooled stream per job, made the thread's current stream */
    UnpaperCudaStream *stream = NULL;
    if (g_device == UNPAPER_DEVICE_CUDA && cuda_stream_pool_global_active()) {
      stream = cuda_stream_pool_global_acquire();
      if (stream) unpaper_cuda_set_current_stream(stream);
    }
#endif
    char names[2][512];
    char *files[2] = {names[0], names[1]};
    for (int j = 0; jb->out_dir && j < jb->out_count && j < 2; j++)
      snprintf(names[j], sizeof(names[j]), "%s/s%06d_%d.pnm", jb->out_dir, i, j);
    int rc = run_one_sheet(jb->cfg, jb->opt, jb->pages + (size_t)i * jb->sheet_in_bytes,
                           jb->page_w, jb->page_h, jb->page_fmt,
                           jb->out ? jb->out + (size_t)i * jb->sheet_out_bytes : NULL,
                           jb->results ? &jb->results[i] : NULL, jb->out_dir ? files : NULL, jb->out_count,
                           (jb->cfg->first_sheet_nr > 0 ? jb->cfg->first_sheet_nr : 1) + i, jb->out_w, jb->out_h);
    if (rc != 0) atomic_fetch_add(&jb->failed, 1);
#ifdef REF_DROPIN
    if (stream) {   /* lib/batch_worker.c:253-255 */
      unpaper_cuda_stream_synchronize_on(stream);
      cuda_stream_pool_global_release(stream);
      unpaper_cuda_set_current_stream(NULL);
    }
#endif
  }
  return NULL;
}

/* Reference process_sheet() over n_sheets sheets with `threads` pthreads —
 * functionally batch_process_parallel() (lib/batch_worker.c:273-296) minus the
 * codecs.  `pages` holds n_sheets*input_count tightly packed pages; `out`
 * (may be NULL) receives tightly packed sheets in the page format.
 * Returns 0 or the number of failed sheets (negative). */
int ref_process_sheets(const B200SheetConfig *cfg, const uint8_t *pages, int page_w,
                       int page_h, int page_fmt, int n_sheets, uint8_t *out,
                       B200SheetResult *results, int threads, int *sheet_w, int *sheet_h) {
  ensure_init();
  Options opt;
  static __thread Rectangle bf_excl[MAX_MASKS];
  options_from_cfg(&opt, bf_excl, cfg);
  int row = av_shim_row_bytes(page_fmt, page_w);
  if (row < 0) return -1;
  /* in: the expected size of the finished sheets when the options change it (the caller sizes
   * `out` with it; every sheet's actual size comes back in results[].sheet_width/height) */
  int sw = (sheet_w && *sheet_w > 0) ? *sheet_w : page_w * cfg->input_count;
  int sh = (sheet_h && *sheet_h > 0) ? *sheet_h : page_h;
  if (sheet_w) *sheet_w = sw;
  if (sheet_h) *sheet_h = sh;
  Job jb = {.cfg = cfg, .opt = &opt, .pages = pages, .out = out, .results = results,
            .page_w = page_w, .page_h = page_h, .page_fmt = page_fmt, .n_sheets = n_sheets,
            .out_w = sw, .out_h = sh,
            .sheet_in_bytes = (size_t)row * page_h * cfg->input_count,
            .sheet_out_bytes = (size_t)av_shim_row_bytes(page_fmt, sw) * sh};
  atomic_init(&jb.next, 0);
  atomic_init(&jb.failed, 0);
  if (threads <= 1) {
    worker(&jb);
  } else {
    pthread_t *th = calloc((size_t)threads, sizeof(*th));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &jb);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  return -atomic_load(&jb.failed);
}

/* The same with the reference's own output stage (sheet_stages.c:536-631): write_output
 * on, `output_count` files per sheet (1, or 2 = the sheet split of :606-621) written by the
 * reference's saveImage() (file.c:186-262: format conversion + saveImageDirect) as
 * <out_dir>/s<sheet>_<page>.pnm.  out_fmt: AVPixelFormat or -1 (= the page format,
 * sheet_stages.c:130-131). */
int ref_process_sheets_files(const B200SheetConfig *cfg, const uint8_t *pages, int page_w,
                             int page_h, int page_fmt, int n_sheets, int out_fmt, int output_count,
                             const char *out_dir, B200SheetResult *results, int threads) {
  ensure_init();
  if (output_count < 1 || output_count > 2 || !out_dir) return -1;
  Options opt;
  static __thread Rectangle bf_excl[MAX_MASKS];
  options_from_cfg(&opt, bf_excl, cfg);
  opt.write_output = true;
  opt.output_count = output_count;
  opt.output_pixel_format = out_fmt < 0 ? AV_PIX_FMT_NONE : out_fmt;
  int row = av_shim_row_bytes(page_fmt, page_w);
  if (row < 0) return -1;
  Job jb = {.cfg = cfg, .opt = &opt, .pages = pages, .out = NULL, .results = results,
            .out_dir = out_dir, .out_count = output_count,
            .page_w = page_w, .page_h = page_h, .page_fmt = page_fmt, .n_sheets = n_sheets,
            .sheet_in_bytes = (size_t)row * page_h * cfg->input_count, .sheet_out_bytes = 0};
  atomic_init(&jb.next, 0);
  atomic_init(&jb.failed, 0);
  if (threads <= 1) {
    worker(&jb);
  } else {
    pthread_t *th = calloc((size_t)threads, sizeof(*th));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &jb);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  return -atomic_load(&jb.failed);
}

/* The reference's saveImage() (file.c:186-262) on an image in caller memory. */
int ref_save_image(const B200HostImage *in, int out_fmt, const char *path) {
  ensure_init();
  Image img = clone_owned(in);
  char *p = strdup(path);
  saveImage(p, img, out_fmt);
  free(p);
  free_image(&img);
  return 0;
}

int ref_online_cpus(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
const char *ref_version(void) { return "unpaper reference CPU backend (oracle/_ref)"; }
