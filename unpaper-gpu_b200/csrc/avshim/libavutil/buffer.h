/* Minimal stand-in for <libavutil/buffer.h> (reference-counted byte buffers).
 * Build shim for hosts without FFmpeg headers; see pixfmt.h. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct AVBuffer {
  uint8_t *data;
  size_t size;
  int refcount;
  void (*free_cb)(void *opaque, uint8_t *data);
  void *opaque;
} AVBuffer;
typedef struct AVBufferRef {
  AVBuffer *buffer;
  uint8_t *data;
  size_t size;
} AVBufferRef;

static inline void *av_malloc(size_t n) { return malloc(n ? n : 1); }
static inline void *av_mallocz(size_t n) { return calloc(1, n ? n : 1); }
static inline void av_free(void *p) { free(p); }

static inline AVBufferRef *av_buffer_create(uint8_t *data, size_t size,
                                            void (*free_cb)(void *, uint8_t *),
                                            void *opaque, int flags) {
  (void)flags;
  AVBuffer *b = (AVBuffer *)calloc(1, sizeof(*b));
  AVBufferRef *r = (AVBufferRef *)calloc(1, sizeof(*r));
  if (!b || !r) { free(b); free(r); return NULL; }
  b->data = data; b->size = size; b->refcount = 1;
  b->free_cb = free_cb; b->opaque = opaque;
  r->buffer = b; r->data = data; r->size = size;
  return r;
}
static inline AVBufferRef *av_buffer_ref(AVBufferRef *src) {
  AVBufferRef *r = (AVBufferRef *)calloc(1, sizeof(*r));
  if (!r) return NULL;
  *r = *src;
  __atomic_add_fetch(&src->buffer->refcount, 1, __ATOMIC_SEQ_CST);
  return r;
}
static inline void av_buffer_unref(AVBufferRef **pref) {
  if (!pref || !*pref) return;
  AVBuffer *b = (*pref)->buffer;
  free(*pref);
  *pref = NULL;
  if (__atomic_sub_fetch(&b->refcount, 1, __ATOMIC_SEQ_CST) == 0) {
    if (b->free_cb) b->free_cb(b->opaque, b->data);
    else free(b->data);
    free(b);
  }
}
#ifdef __cplusplus
}
#endif
