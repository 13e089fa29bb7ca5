/* stand-in for <libavutil/common.h>; Build shim for hosts without FFmpeg headers. */
#pragma once
#include <stdint.h>
static inline uint8_t av_clip_uint8(int a) {
  if (a & (~0xFF)) return (uint8_t)((~a) >> 31);
  return (uint8_t)a;
}
