/* host_api.c — layer (2) of include/unpaper_b200.h: one backend op on an image
 * that lives in caller memory.  Upload, run the `backend_cuda` entry point,
 * download.  This is what the parity tests and foreign-language hosts bind. */
#define _GNU_SOURCE
#include <libavutil/frame.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "host.h"

void b200_center_mask(Image image, const Point center, const Rectangle area);

typedef struct { AVFrame f; Image img; } Borrow;

/* Wrap caller memory without copying; the device copy hangs on f.opaque_ref. */
static int borrow(Borrow *b, const B200HostImage *h) {
  if (!h || !h->data || h->width <= 0 || h->height <= 0) { b200_set_error("bad image"); return -1; }
  int row = b200_fmt_row_bytes(h->format, h->width);
  if (row < 0 || h->linesize < row) { b200_set_error("bad format/linesize"); return -1; }
  memset(&b->f, 0, sizeof(b->f));
  b->f.data[0] = h->data; b->f.linesize[0] = h->linesize;
  b->f.width = h->width; b->f.height = h->height; b->f.format = h->format;
  b->img = (Image){.frame = &b->f, .background = h->background, .abs_black_threshold = h->abs_black_threshold};
  return 0;
}
static void give_back(Borrow *b) {
  image_ensure_cpu(&b->img);
  image_cuda_release(&b->img);
}
static void drop(Borrow *b) { image_cuda_release(&b->img); }

static Image clone_owned(const B200HostImage *h) {
  Image img = {.frame = av_frame_alloc(), .background = h->background, .abs_black_threshold = h->abs_black_threshold};
  img.frame->width = h->width; img.frame->height = h->height; img.frame->format = h->format;
  if (av_frame_get_buffer(img.frame, 8) < 0) b200_fatal("unable to allocate image buffer");
  int row = b200_fmt_row_bytes(h->format, h->width);
  for (int y = 0; y < h->height; y++)
    memcpy(img.frame->data[0] + (size_t)y * img.frame->linesize[0], h->data + (size_t)y * h->linesize, (size_t)row);
  return img;
}
static int export_owned(Image *img, B200HostImage *out) {
  image_ensure_cpu(img);
  int rc = 0;
  if (out->width != img->frame->width || out->height != img->frame->height || out->format != img->frame->format) {
    b200_set_error("result is %dx%d, caller expected %dx%d", img->frame->width, img->frame->height, out->width, out->height);
    rc = -2;
  } else {
    int row = b200_fmt_row_bytes(out->format, out->width);
    for (int y = 0; y < out->height; y++)
      memcpy(out->data + (size_t)y * out->linesize, img->frame->data[0] + (size_t)y * img->frame->linesize[0], (size_t)row);
  }
  image_cuda_release(img);
  av_frame_free(&img->frame);
  return rc;
}

#define INPLACE(h, call)                      \
  Borrow b;                                   \
  if (borrow(&b, (h))) return -1;             \
  call;                                       \
  give_back(&b);                              \
  return 0

int unpaper_b200_host_wipe_rectangle(B200HostImage *img, const Rectangle *area, Pixel color) {
  INPLACE(img, backend_cuda.wipe_rectangle(b.img, *area, color));
}
int unpaper_b200_host_copy_rectangle(const B200HostImage *src, B200HostImage *dst, const Rectangle *area, Point target) {
  Borrow s, d;
  if (borrow(&s, src) || borrow(&d, dst)) return -1;
  backend_cuda.copy_rectangle(s.img, d.img, *area, target);
  give_back(&d); drop(&s);
  return 0;
}
int unpaper_b200_host_center_image(const B200HostImage *src, B200HostImage *dst, Point origin, RectangleSize size) {
  Borrow s, d;
  if (borrow(&s, src) || borrow(&d, dst)) return -1;
  backend_cuda.center_image(s.img, d.img, origin, size);
  give_back(&d); drop(&s);
  return 0;
}
int unpaper_b200_host_stretch(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  Image w = clone_owned(img);
  backend_cuda.stretch_and_replace(&w, (RectangleSize){out->width, out->height}, (Interpolation)interp);
  return export_owned(&w, out);
}
int unpaper_b200_host_resize(const B200HostImage *img, B200HostImage *out, int32_t interp) {
  Image w = clone_owned(img);
  backend_cuda.resize_and_replace(&w, (RectangleSize){out->width, out->height}, (Interpolation)interp);
  return export_owned(&w, out);
}
int unpaper_b200_host_flip_rotate_90(const B200HostImage *img, B200HostImage *out, int32_t direction) {
  Image w = clone_owned(img);
  backend_cuda.flip_rotate_90(&w, (RotationDirection)direction);
  return export_owned(&w, out);
}
int unpaper_b200_host_mirror(B200HostImage *img, Direction direction) {
  INPLACE(img, backend_cuda.mirror(b.img, direction));
}
int unpaper_b200_host_shift(const B200HostImage *img, B200HostImage *out, Delta d) {
  Image w = clone_owned(img);
  backend_cuda.shift_image(&w, d);
  return export_owned(&w, out);
}
int unpaper_b200_host_apply_masks(B200HostImage *img, const Rectangle *masks, size_t n, Pixel color) {
  INPLACE(img, backend_cuda.apply_masks(b.img, masks, n, color));
}
int unpaper_b200_host_apply_wipes(B200HostImage *img, const Wipes *wipes, Pixel color) {
  INPLACE(img, backend_cuda.apply_wipes(b.img, *wipes, color));
}
int unpaper_b200_host_apply_border(B200HostImage *img, const Border *border, Pixel color) {
  INPLACE(img, backend_cuda.apply_border(b.img, *border, color));
}
int unpaper_b200_host_detect_masks(const B200HostImage *img, const MaskDetectionParameters *p,
                                   const Point *points, size_t n, Rectangle *masks_out) {
  Borrow b;
  if (borrow(&b, img)) return -1;
  size_t c = backend_cuda.detect_masks(b.img, *p, points, n, masks_out);
  drop(&b);
  return (int)c;
}
int unpaper_b200_host_center_mask(B200HostImage *img, Point center, const Rectangle *area) {
  INPLACE(img, b200_center_mask(b.img, center, *area));
}
int unpaper_b200_host_align_mask(B200HostImage *img, const Rectangle *inside, const Rectangle *outside,
                                 const MaskAlignmentParameters *p) {
  INPLACE(img, backend_cuda.align_mask(b.img, *inside, *outside, *p));
}
int unpaper_b200_host_detect_border(const B200HostImage *img, const BorderScanParameters *p,
                                    const Rectangle *outside, Border *out) {
  Borrow b;
  if (borrow(&b, img)) return -1;
  *out = backend_cuda.detect_border(b.img, *p, *outside);
  drop(&b);
  return 0;
}
int unpaper_b200_host_blackfilter(B200HostImage *img, const BlackfilterParameters *p) {
  INPLACE(img, backend_cuda.blackfilter(b.img, *p));
}
int unpaper_b200_host_blurfilter(B200HostImage *img, const BlurfilterParameters *p, uint8_t abs_white) {
  INPLACE(img, backend_cuda.blurfilter(b.img, *p, abs_white));
}
int unpaper_b200_host_noisefilter(B200HostImage *img, uint64_t intensity, uint8_t min_white_level) {
  INPLACE(img, backend_cuda.noisefilter(b.img, intensity, min_white_level));
}
int unpaper_b200_host_grayfilter(B200HostImage *img, const GrayfilterParameters *p) {
  INPLACE(img, backend_cuda.grayfilter(b.img, *p));
}
int unpaper_b200_host_detect_rotation(const B200HostImage *img, const Rectangle *mask,
                                      const DeskewParameters *p, float *radians_out) {
  Borrow b;
  if (borrow(&b, img)) return -1;
  *radians_out = backend_cuda.detect_rotation(b.img, *mask, *p);
  drop(&b);
  return 0;
}
int unpaper_b200_host_deskew(B200HostImage *img, const Rectangle *mask, float radians, int32_t interp) {
  INPLACE(img, backend_cuda.deskew(b.img, *mask, radians, (Interpolation)interp));
}

/* options_init() + options_init_filter_defaults() + CLI threshold defaults
 * (reference lib/options.c:22-170, src/cli/cli_options.c:229-274,:1108-1109) */
void unpaper_b200_sheet_config_defaults(B200SheetConfig *c) {
  memset(c, 0, sizeof(*c));
  const float degrees = 3.14159265358979323846 / 180.0;
  c->layout = LAYOUT_SINGLE;
  c->input_count = 1;
  c->interpolate_type = INTERP_CUBIC;
  c->sheet_background = (Pixel){255, 255, 255};
  c->mask_color = (Pixel){255, 255, 255};
  float whiteThreshold = 0.9, blackThreshold = 0.33;
  c->abs_black_threshold = 0xFF * (1.0 - blackThreshold);
  c->abs_white_threshold = 0xFF * (whiteThreshold);
  c->noisefilter_intensity = 4;
  c->blackfilter = (BlackfilterParameters){
      .scan_size = {20, 20}, .scan_step = {5, 5}, .scan_depth = {500, 500},
      .scan_direction = {true, true}, .abs_threshold = UINT8_MAX * 0.95f, .intensity = 20,
      .exclusions_count = 0, .exclusions = NULL};
  c->blurfilter = (BlurfilterParameters){.scan_size = {100, 100}, .scan_step = {50, 50}, .intensity = 0.01f};
  c->grayfilter = (GrayfilterParameters){.scan_size = {50, 50}, .scan_step = {20, 20}, .abs_threshold = UINT8_MAX * 0.5f};
  (void)degrees;
  c->deskew = (DeskewParameters){
      .deskewScanRangeRad = 5.0f * M_PI / 180.0, .deskewScanStepRad = 0.1f * M_PI / 180.0,
      .deskewScanDeviationRad = 1.0f * M_PI / 180.0, .deskewScanSize = 1500, .deskewScanDepth = 0.5f,
      .scan_edges = {.left = true, .top = false, .right = true, .bottom = false}};
  c->mask_detection = (MaskDetectionParameters){
      .scan_size = {50, 50}, .scan_step = {5, 5}, .scan_depth = {-1, -1},
      .scan_direction = {true, false}, .scan_threshold = {0.1f, 0.1f},
      .minimum_width = 100, .maximum_width = -1, .minimum_height = 100, .maximum_height = -1};
  c->mask_alignment = (MaskAlignmentParameters){.alignment = {false, false, false, false}, .margin = {0, 0}};
  c->border_scan = (BorderScanParameters){
      .scan_size = {5, 5}, .scan_step = {5, 5}, .scan_threshold = {5, 5}, .scan_direction = {false, true}};
  c->output_count = 1;
  c->first_sheet_nr = 1;
  c->sheet_size = c->stretch_size = c->page_size = c->post_stretch_size = c->post_page_size = (RectangleSize){-1, -1};
  c->pre_zoom_factor = c->post_zoom_factor = 1.0f;
}

/* ---- output side (sheet_stage_output, sheet_stages.c:536-631 -> saveImage) ---- */

/* the format saveImage() really writes for a requested one (file.c:201-208) */
int unpaper_b200_output_format(int av_pix_fmt) {
  if (av_pix_fmt == AV_PIX_FMT_Y400A) return AV_PIX_FMT_GRAY8;
  if (av_pix_fmt == AV_PIX_FMT_MONOBLACK) return AV_PIX_FMT_MONOWHITE;
  return av_pix_fmt;
}

int unpaper_b200_host_convert_format(const B200HostImage *in, B200HostImage *out) {
  if (!in || !out || !out->data) { b200_set_error("bad image"); return -1; }
  if (out->width != in->width || out->height != in->height) { b200_set_error("convert: size mismatch"); return -1; }
  int ofmt = unpaper_b200_output_format(out->format);
  if (ofmt != out->format) { b200_set_error("convert: format %d is written as %d", out->format, ofmt); return -1; }
  int orow = b200_fmt_row_bytes(out->format, out->width);
  if (orow < 0 || b200_fmt_to_dev(out->format) < 0 || out->linesize < orow) { b200_set_error("bad format/linesize"); return -1; }
  Borrow b;
  if (borrow(&b, in)) return -1;
  if (in->format == out->format) {   /* file.c:210: written as it is */
    for (int y = 0; y < in->height; y++)
      memcpy(out->data + (size_t)y * out->linesize, in->data + (size_t)y * in->linesize, (size_t)orow);
    return 0;
  }
  image_ensure_cuda(&b.img);
  DImg sv, dv;
  if (!b200_image_view(&b.img, &sv)) { drop(&b); b200_set_error("unsupported pixel format"); return -1; }
  size_t bytes = (size_t)orow * out->height;
  uint8_t *d = (uint8_t *)b200_dev_alloc(bytes);
  dv = sv; dv.data = d; dv.pitch = orow; dv.fmt = b200_fmt_to_dev(out->format);
  cudaStream_t st = b200_rt_stream();
  b200k_convert_out(st, sv, dv, 1, 0, 0);
  CUDA_OK(cudaMemcpy2DAsync(out->data, (size_t)out->linesize, d, (size_t)orow, (size_t)orow, (size_t)out->height,
                            cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  b200_dev_free(d);
  drop(&b);
  return 0;
}

/* saveImageDirect (file.c:134-176): PNM header + tight rows */
int unpaper_b200_pnm_header(int av_pix_fmt, int width, int height, char *buf, size_t cap) {
  int n;
  switch (av_pix_fmt) {
  case AV_PIX_FMT_GRAY8: n = snprintf(buf, cap, "P5\n%d %d\n255\n", width, height); break;
  case AV_PIX_FMT_RGB24: n = snprintf(buf, cap, "P6\n%d %d\n255\n", width, height); break;
  case AV_PIX_FMT_MONOWHITE: n = snprintf(buf, cap, "P4\n%d %d\n", width, height); break;
  default: b200_set_error("pnm: unsupported pixel format %d", av_pix_fmt); return -1;
  }
  if (n < 0 || (size_t)n >= cap) { b200_set_error("pnm: header buffer too small"); return -1; }
  return n;
}

int unpaper_b200_write_pnm(const char *path, const uint8_t *data, int linesize, int width, int height, int av_pix_fmt) {
  char hdr[64];
  int n = unpaper_b200_pnm_header(av_pix_fmt, width, height, hdr, sizeof(hdr));
  int row = b200_fmt_row_bytes(av_pix_fmt, width);
  if (n < 0 || row < 0 || !data || linesize < row) { if (n >= 0) b200_set_error("pnm: bad arguments"); return -1; }
  FILE *f = fopen(path, "wb");
  if (!f) { b200_set_error("pnm: unable to open %s", path); return -1; }
  bool ok = fwrite(hdr, 1, (size_t)n, f) == (size_t)n;
  if (linesize == row) ok = ok && fwrite(data, 1, (size_t)row * height, f) == (size_t)row * height;
  else
    for (int y = 0; ok && y < height; y++) ok = fwrite(data + (size_t)y * linesize, 1, (size_t)row, f) == (size_t)row;
  ok = (fclose(f) == 0) && ok;
  if (!ok) { b200_set_error("pnm: short write to %s", path); return -1; }
  return 0;
}

