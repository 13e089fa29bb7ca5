/* Minimal stand-in for <libavutil/pixfmt.h>: only the pixel formats the
 * unpaper sheet path touches.  Used when FFmpeg's own headers are absent (this
 * image has none): by the B200 backend build and by oracle/Makefile.  Numeric
 * values follow FFmpeg's enum so a binary built against the real header agrees. */
#pragma once
enum AVPixelFormat {
  AV_PIX_FMT_NONE = -1,
  AV_PIX_FMT_YUV420P = 0,
  AV_PIX_FMT_RGB24 = 2,
  AV_PIX_FMT_GRAY8 = 8,
  AV_PIX_FMT_MONOWHITE = 9,
  AV_PIX_FMT_MONOBLACK = 10,
  AV_PIX_FMT_PAL8 = 11,
  AV_PIX_FMT_YA8 = 58,
  AV_PIX_FMT_Y400A = AV_PIX_FMT_YA8,
};
