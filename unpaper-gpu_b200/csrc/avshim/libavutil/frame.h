/* Minimal stand-in for <libavutil/frame.h>: the AVFrame fields and the four
 * allocation calls unpaper's imageprocess/ layer uses.  Build shim for hosts
 * without FFmpeg headers; see pixfmt.h. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "libavutil/buffer.h"
#include "libavutil/pixfmt.h"
#ifdef __cplusplus
extern "C" {
#endif
#define AV_NUM_DATA_POINTERS 8
typedef struct AVFrame {
  uint8_t *data[AV_NUM_DATA_POINTERS];
  int linesize[AV_NUM_DATA_POINTERS];
  int width, height;
  int format;
  AVBufferRef *buf[AV_NUM_DATA_POINTERS];
  AVBufferRef *opaque_ref;
} AVFrame;

static inline int av_strerror(int errnum, char *buf, size_t n) {
  snprintf(buf, n, "error %d", errnum);
  return 0;
}
static inline AVFrame *av_frame_alloc(void) {
  AVFrame *f = (AVFrame *)calloc(1, sizeof(AVFrame));
  if (f) f->format = AV_PIX_FMT_NONE;
  return f;
}
static inline int av_shim_row_bytes(int format, int width) {
  switch (format) {
  case AV_PIX_FMT_GRAY8: case AV_PIX_FMT_PAL8: return width;
  case AV_PIX_FMT_Y400A: return width * 2;
  case AV_PIX_FMT_RGB24: return width * 3;
  case AV_PIX_FMT_MONOWHITE: case AV_PIX_FMT_MONOBLACK: return (width + 7) / 8;
  default: return -1;
  }
}
static inline int av_frame_get_buffer(AVFrame *f, int align) {
  int row = av_shim_row_bytes(f->format, f->width);
  if (row < 0 || f->width <= 0 || f->height <= 0) return -22;
  if (align <= 0) align = 32;
  int ls = (row + align - 1) / align * align;
  size_t size = (size_t)ls * (size_t)f->height + 64;
  uint8_t *mem = (uint8_t *)calloc(1, size);
  if (!mem) return -12;
  f->buf[0] = av_buffer_create(mem, size, NULL, NULL, 0);
  if (!f->buf[0]) { free(mem); return -12; }
  f->data[0] = mem;
  f->linesize[0] = ls;
  return 0;
}
static inline void av_frame_free(AVFrame **pf) {
  if (!pf || !*pf) return;
  AVFrame *f = *pf;
  for (int i = 0; i < AV_NUM_DATA_POINTERS; i++) av_buffer_unref(&f->buf[i]);
  av_buffer_unref(&f->opaque_ref);
  free(f);
  *pf = NULL;
}
static inline AVFrame *av_frame_clone(const AVFrame *src) {
  AVFrame *f = av_frame_alloc();
  if (!f) return NULL;
  *f = *src;
  for (int i = 0; i < AV_NUM_DATA_POINTERS; i++)
    f->buf[i] = src->buf[i] ? av_buffer_ref(src->buf[i]) : NULL;
  f->opaque_ref = src->opaque_ref ? av_buffer_ref(src->opaque_ref) : NULL;
  return f;
}
#ifdef __cplusplus
}
#endif
