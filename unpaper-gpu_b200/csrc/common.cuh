// common.cuh — pixel access with the reference's get_pixel/set_pixel semantics
// (reference imageprocess/pixel.c:20-173) and small warp helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev.h"

struct Px { int r, g, b; };

__device__ __forceinline__ bool in_img(const DImg &im, int x, int y) {
  return (unsigned)x < (unsigned)im.w && (unsigned)y < (unsigned)im.h;
}

// pixel.c:20-63 for in-image coordinates
__device__ __forceinline__ Px px_load(const DImg &im, int x, int y) {
  const uint8_t *row = im.data + (size_t)y * (size_t)im.pitch;
  int v;
  switch (im.fmt) {
  case DF_GRAY8: v = row[x]; return Px{v, v, v};
  case DF_Y400A: v = row[2 * x]; return Px{v, v, v};
  case DF_RGB24: { const uint8_t *p = row + 3 * x; return Px{p[0], p[1], p[2]}; }
  case DF_MONOWHITE: v = (row[x >> 3] & (128 >> (x & 7))) ? 0 : 255; return Px{v, v, v};
  default: v = (row[x >> 3] & (128 >> (x & 7))) ? 255 : 0; return Px{v, v, v};
  }
}

// pixel.c:23-25: outside the image reads as white
__device__ __forceinline__ Px px_get(const DImg &im, int x, int y) {
  if (!in_img(im, x, y)) return Px{255, 255, 255};
  return px_load(im, x, y);
}

__device__ __forceinline__ int px_gray(Px p) { return (p.r + p.g + p.b) / 3; }      // pixel.c:16-18
__device__ __forceinline__ int px_light(Px p) { return min(p.r, min(p.g, p.b)); }    // pixel.c:109-112
__device__ __forceinline__ int px_darkinv(Px p) { return max(p.r, max(p.g, p.b)); }  // pixel.c:128-131

// pixel.c:136-173 for in-image coordinates.  Mono formats go through word
// atomics so that threads sharing a byte cannot lose each other's bits.
__device__ __forceinline__ void px_store(const DImg &im, int x, int y, int r, int g, int b) {
  uint8_t *row = im.data + (size_t)y * (size_t)im.pitch;
  switch (im.fmt) {
  case DF_GRAY8: row[x] = (uint8_t)((r + g + b) / 3); return;
  case DF_Y400A: row[2 * x] = (uint8_t)((r + g + b) / 3); row[2 * x + 1] = 0xFF; return;
  case DF_RGB24: { uint8_t *p = row + 3 * x; p[0] = (uint8_t)r; p[1] = (uint8_t)g; p[2] = (uint8_t)b; return; }
  default: {
    bool black = ((r + g + b) / 3) < (int)im.abt;
    if (im.fmt == DF_MONOWHITE) black = !black;
    uintptr_t a = (uintptr_t)(row + (x >> 3));
    unsigned *w = (unsigned *)(a & ~(uintptr_t)3);
    unsigned m = (unsigned)(128 >> (x & 7)) << (8 * (unsigned)(a & 3));
    if (!black) atomicOr(w, m); else atomicAnd(w, ~m);
    return;
  }
  }
}

__device__ __forceinline__ void px_set(const DImg &im, int x, int y, int r, int g, int b) {
  if (in_img(im, x, y)) px_store(im, x, y, r, g, b);
}

// primitives.c:48-61 + :95-100 (point_in_rectangle normalises first)
__device__ __forceinline__ bool pt_in_rect(int x, int y, const DRect &r) {
  int ax = min(r.x0, r.x1), bx = max(r.x0, r.x1);
  int ay = min(r.y0, r.y1), by = max(r.y0, r.y1);
  return x >= ax && x <= bx && y >= ay && y <= by;
}

__device__ __forceinline__ int bytes_pp(int fmt) { return fmt == DF_GRAY8 ? 1 : fmt == DF_Y400A ? 2 : fmt == DF_RGB24 ? 3 : 0; }

__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
