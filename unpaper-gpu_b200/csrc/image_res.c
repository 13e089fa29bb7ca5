/* image_res.c — host/device residency of an Image.
 *
 * Same protocol as reference imageprocess/image_cuda.c: the device copy is
 * owned by a small state object hung on frame->opaque_ref (an AVBufferRef whose
 * free callback returns the buffer to the cache), host pixels stay in
 * frame->data[0], validity is tracked with cpu_dirty / cuda_dirty, and the
 * device layout equals the host layout (same pitch) so that upload and
 * download are single copies.  Differences: no upload of an uninitialised host
 * buffer when the first touch is a full-image wipe (the reference's
 * create_image(fill) wart, image.c:38-40 -> image_cuda.c:186-196), and device
 * memory comes from the bucketed cache in rt.c.
 */
#include <libavutil/frame.h>
#include <string.h>

#include "host.h"

typedef struct {
  void *dptr;
  size_t bytes;
  int width, height, format, pitch;
  bool cpu_dirty;   /* host copy is newer */
  bool cuda_dirty;  /* device copy is newer */
  bool owns;        /* dptr came from b200_dev_alloc */
  bool foreign;     /* dptr was handed in by create_image_from_gpu() */
  bool owns_foreign; /* ... with owns_memory: cudaFree it on release */
  bool host_valid;  /* host buffer holds defined data */
  int device;
} ImageRes;

static void res_free(void *opaque, uint8_t *data) {
  (void)opaque;
  ImageRes *st = (ImageRes *)data;
  if (!st) return;
  if (st->dptr && st->owns) b200_dev_free(st->dptr);
  if (st->dptr && st->owns_foreign) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != st->device) cudaSetDevice(st->device);
    cudaFree(st->dptr);
    if (cur != st->device) cudaSetDevice(cur);
  }
  free(st);
}

static ImageRes *res_get(Image *image, bool create) {
  if (!image || !image->frame) return NULL;
  AVFrame *f = image->frame;
  if (f->opaque_ref) return (ImageRes *)f->opaque_ref->data;
  if (!create) return NULL;
  ImageRes *st = (ImageRes *)calloc(1, sizeof(*st));
  if (!st) b200_fatal("out of memory");
  st->cpu_dirty = true;
  st->host_valid = true;
  f->opaque_ref = av_buffer_create((uint8_t *)st, sizeof(*st), res_free, NULL, 0);
  if (!f->opaque_ref) b200_fatal("out of memory");
  return st;
}

static size_t frame_bytes(const AVFrame *f) { return (size_t)f->linesize[0] * (size_t)f->height; }

static void res_ensure_buffer(ImageRes *st, const AVFrame *f) {
  size_t need = frame_bytes(f) + 16;   /* +16: mono word atomics may touch the tail word */
  /* a buffer handed in by create_image_from_gpu() (GRAY8 / RGB24 only: byte stores, no
   * tail word) is exactly pitch*height bytes and is never replaced while the frame keeps
   * its geometry */
  size_t have_need = st->foreign ? frame_bytes(f) : need;
  bool same = st->dptr && st->width == f->width && st->height == f->height &&
              st->format == f->format && st->pitch == f->linesize[0] && st->bytes >= have_need;
  if (same) return;
  if (st->dptr && st->owns) b200_dev_free(st->dptr);
  if (st->dptr && st->owns_foreign) cudaFree(st->dptr);
  st->foreign = false; st->owns_foreign = false;
  st->dptr = b200_dev_alloc(need);
  st->owns = true;
  st->bytes = need;
  st->width = f->width; st->height = f->height; st->format = f->format; st->pitch = f->linesize[0];
  st->device = b200_rt_device();
  st->cpu_dirty = true;
  st->cuda_dirty = false;
}

void image_ensure_cuda_alloc(Image *image) {
  ImageRes *st = res_get(image, true);
  if (!st) return;
  res_ensure_buffer(st, image->frame);
}

void image_ensure_cuda(Image *image) {
  ImageRes *st = res_get(image, true);
  if (!st) return;
  res_ensure_buffer(st, image->frame);
  if (st->cpu_dirty) {
    cudaStream_t s = b200_rt_stream();
    CUDA_OK(cudaMemcpyAsync(st->dptr, image->frame->data[0], frame_bytes(image->frame),
                            cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaStreamSynchronize(s));
    st->cpu_dirty = false;
    st->cuda_dirty = false;
  }
}

void image_ensure_cpu(Image *image) {
  ImageRes *st = res_get(image, false);
  if (!st || !st->dptr || !st->cuda_dirty) return;
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemcpyAsync(image->frame->data[0], st->dptr, frame_bytes(image->frame),
                          cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  st->cuda_dirty = false;
  st->cpu_dirty = false;
}

void image_mark_cpu_dirty(Image *image) {
  ImageRes *st = res_get(image, true);
  if (st) { st->cpu_dirty = true; st->cuda_dirty = false; }
}
void image_mark_cuda_dirty(Image *image) {
  ImageRes *st = res_get(image, true);
  if (st) { st->cuda_dirty = true; st->cpu_dirty = false; }
}

void image_cuda_release(Image *image) {
  if (!image || !image->frame) return;
  av_buffer_unref(&image->frame->opaque_ref);
}

bool image_is_gpu_resident(Image *image) {
  ImageRes *st = res_get(image, false);
  return st && st->dptr && !st->cpu_dirty;
}
void image_set_gpu_resident(Image *image, bool resident) {
  ImageRes *st = res_get(image, true);
  if (!st) return;
  if (resident) { st->cpu_dirty = false; st->cuda_dirty = true; }
  else { st->cpu_dirty = true; st->cuda_dirty = false; }
}
void *image_get_gpu_ptr(Image *image) {
  ImageRes *st = res_get(image, false);
  return (st && st->dptr && !st->cpu_dirty) ? st->dptr : NULL;
}
size_t image_get_gpu_pitch(Image *image) {
  ImageRes *st = res_get(image, false);
  return (st && st->dptr && !st->cpu_dirty) ? (size_t)st->pitch : 0;
}

Image create_image_from_gpu(void *gpu_ptr, size_t pitch, int width, int height, int pixel_format,
                            Pixel background, uint8_t abs_black_threshold, bool owns_memory) {
  Image img = {.frame = NULL, .background = background, .abs_black_threshold = abs_black_threshold};
  if (!gpu_ptr || width <= 0 || height <= 0) return img;
  if (pixel_format != AV_PIX_FMT_GRAY8 && pixel_format != AV_PIX_FMT_RGB24) return img;
  AVFrame *f = av_frame_alloc();
  if (!f) return img;
  f->width = width; f->height = height; f->format = pixel_format;
  /* host buffer with the GPU pitch as linesize (reference image_cuda.c:307-362) */
  size_t bytes = pitch * (size_t)height;
  uint8_t *buf = (uint8_t *)av_malloc(bytes + 64);
  if (!buf) { av_frame_free(&f); return img; }
  f->buf[0] = av_buffer_create(buf, bytes + 64, NULL, NULL, 0);
  f->data[0] = buf;
  f->linesize[0] = (int)pitch;
  img.frame = f;
  ImageRes *st = res_get(&img, true);
  st->dptr = gpu_ptr; st->owns = false; st->foreign = true; st->bytes = bytes;
  st->width = width; st->height = height; st->format = pixel_format; st->pitch = (int)pitch;
  st->device = b200_rt_device();
  st->cpu_dirty = false; st->cuda_dirty = true; st->host_valid = false;
  /* owns_memory (image_cuda.c:307-362): the block was cudaMalloc'ed by the decoder and
   * now belongs to the image; it is cudaFree'd when the image is released */
  st->owns_foreign = owns_memory;
  return img;
}

/* internal: device view of a resident image */
bool b200_image_view(Image *image, DImg *out) {
  ImageRes *st = res_get(image, false);
  if (!st || !st->dptr) return false;
  int df = b200_fmt_to_dev(image->frame->format);
  if (df < 0) return false;
  out->data = (uint8_t *)st->dptr;
  out->w = st->width; out->h = st->height; out->pitch = st->pitch; out->fmt = df;
  out->abt = image->abs_black_threshold;
  out->bg[0] = image->background.r; out->bg[1] = image->background.g; out->bg[2] = image->background.b;
  return true;
}
