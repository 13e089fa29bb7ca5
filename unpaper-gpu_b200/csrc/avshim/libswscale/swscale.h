/* stand-in for <libswscale/swscale.h>: the YUV->RGB conversion branch of the
 * decode stage is outside the hot path; the context is never created here.
 * Build shim for hosts without FFmpeg headers. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#define SWS_BILINEAR 2
struct SwsContext;
static inline struct SwsContext *sws_getContext(int a, int b, int c, int d,
                                                int e, int f, int g, void *h,
                                                void *i, void *j) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; (void)i; (void)j;
  return NULL;
}
static inline int sws_scale(struct SwsContext *c, const uint8_t *const *s,
                            const int *ss, int y, int h, uint8_t *const *d,
                            const int *ds) {
  (void)c; (void)s; (void)ss; (void)y; (void)h; (void)d; (void)ds;
  return 0;
}
static inline void sws_freeContext(struct SwsContext *c) { (void)c; }
