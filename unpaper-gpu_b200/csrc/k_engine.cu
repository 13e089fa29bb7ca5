// k_engine.cu — small helpers of the sheet engine.
#include "common.cuh"
#include "launch.h"

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

// zero the per-sheet counters and results of every page of a group
__global__ void k_page_reset(DPage *pages, int npages) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  pg.list_n = 0; pg.nf_clusters = 0; pg.bf_fills = 0; pg.error = 0;
  if (pg.buf[0]) { pg.img.data = pg.buf[0]; pg.other = pg.buf[1]; }
  pg.move.enabled = 0; pg.move.use_masks = 0; pg.move.nseg = 0;
  pg.mask_count = 0; pg.mask_count_deskew = 0; pg.ink_ok = 0;
  for (int i = 0; i < D_MAX_MASKS; i++) {
    pg.rotation[i] = 0.0f; pg.rot_sin[i] = 0.0f; pg.rot_cos[i] = 1.0f; pg.rot_apply[i] = 0; pg.centered[i] = 0; pg.rot_more[i] = 0;
    pg.mask_valid[i] = 0;
    for (int e = 0; e < 4; e++) { pg.edge_count[i][e] = 0; pg.rot_angle_idx[i][e] = -1; }
  }
}

// after a pass img -> other: the rendered buffer becomes the working image
__global__ void k_swap_sheets(DPage *pages, int npages) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  DPage &pg = pages[p];
  uint8_t *t = pg.img.data; pg.img.data = pg.other; pg.other = t;
}

// the next pass img -> other renders sheet p straight into base + p * stride (the caller's output)
__global__ void k_set_other(DPage *pages, int npages, uint8_t *base, size_t stride) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  pages[p].other = base + (size_t)p * stride;
}

// the working sheet of every page now has this geometry (after a size-changing pass, or back to
// the decoded size at the start of a sheet)
__global__ void k_set_geometry(DPage *pages, int npages, int w, int h, int pitch) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npages) return;
  pages[p].img.w = w; pages[p].img.h = h; pages[p].img.pitch = pitch;
}

// static job tables (wipes, borders, pre-masks: geometry known at engine creation) carry an
// image descriptor per page; point them at the page's CURRENT working buffer
__global__ void k_retarget_jobs(const DPage *pages, int npages, DFillJob *fills, int nfill, DMaskJob *masks, int nmask, int stride) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nfill) { int p = i % stride; if (p < npages) fills[i].img = pages[p].img; }
  if (i < nmask) { int p = i % stride; if (p < npages) masks[i].img = pages[p].img; }
}

// Strided row copy between packed host-layout images and pitched device
// images, 16 bytes per thread where alignment allows.
__global__ void k_pack_rows(const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                            int row_bytes, int rows, size_t src_stride, size_t dst_stride) {
  const uint8_t *s = src + (size_t)blockIdx.z * src_stride;
  uint8_t *d = dst + (size_t)blockIdx.z * dst_stride;
  for (int y = blockIdx.y; y < rows; y += gridDim.y) {
    const uint8_t *sr = s + (size_t)y * src_pitch;
    uint8_t *dr = d + (size_t)y * dst_pitch;
    bool vec = ((((uintptr_t)sr) | ((uintptr_t)dr)) & 15) == 0;
    if (vec) {
      int n16 = row_bytes >> 4;
      const uint4 *s4 = (const uint4 *)sr; uint4 *d4 = (uint4 *)dr;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) d4[i] = s4[i];
      for (int i = (n16 << 4) + blockIdx.x * blockDim.x + threadIdx.x; i < row_bytes; i += gridDim.x * blockDim.x) dr[i] = sr[i];
    } else {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_bytes; i += gridDim.x * blockDim.x) dr[i] = sr[i];
    }
  }
}

// saveImage()'s pixel-format conversion (file.c:197-260) for `nimages` images of
// equal geometry: MONOWHITE output thresholds gray < abs_black_threshold with a
// cleared tail in the last byte (:211-243), MONOBLACK -> MONOWHITE inverts whole
// bytes (:244-255), everything else is copy_rectangle()'s get_pixel/set_pixel.
// One thread per output byte (mono) or pixel.
__global__ void k_convert_out(DImg src, DImg dst, size_t src_stride, size_t dst_stride) {
  src.data += (size_t)blockIdx.z * src_stride;
  dst.data += (size_t)blockIdx.z * dst_stride;
  int y = blockIdx.y;
  if (dst.fmt == DF_MONOWHITE) {
    int row_bytes = (src.w + 7) / 8;
    for (int bx = blockIdx.x * blockDim.x + threadIdx.x; bx < row_bytes; bx += gridDim.x * blockDim.x) {
      unsigned v = 0;
      if (src.fmt == DF_MONOBLACK) v = src.data[(size_t)y * src.pitch + bx] ^ 0xFFu;
      else if (src.fmt == DF_GRAY8 && bx * 8 + 8 <= src.w && ((src.pitch | (unsigned)(uintptr_t)src.data) & 7) == 0) {
        uint2 q = *(const uint2 *)(src.data + (size_t)y * src.pitch + (size_t)bx * 8);
        unsigned t4 = (unsigned)src.abt * 0x01010101u;
        unsigned m0 = __vcmpltu4(q.x, t4), m1 = __vcmpltu4(q.y, t4);   // 0xFF per dark byte
#pragma unroll
        for (int k = 0; k < 4; k++) {
          v |= ((m0 >> (8 * k)) & 1u) << (7 - k);
          v |= ((m1 >> (8 * k)) & 1u) << (3 - k);
        }
      } else {
        for (int k = 0; k < 8; k++) {
          int x = bx * 8 + k;
          if (x < src.w && px_gray(px_load(src, x, y)) < src.abt) v |= 0x80u >> k;
        }
      }
      dst.data[(size_t)y * dst.pitch + bx] = (uint8_t)v;
    }
  } else {
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < src.w; x += gridDim.x * blockDim.x) {
      Px p = px_load(src, x, y);
      px_store(dst, x, y, p.r, p.g, p.b);
    }
  }
}

extern "C" {
void b200k_convert_out(cudaStream_t st, DImg src, DImg dst, int nimages, size_t src_stride, size_t dst_stride) {
  if (nimages <= 0 || src.w <= 0 || src.h <= 0) return;
  unsigned per_row = dst.fmt == DF_MONOWHITE ? (unsigned)(src.w + 7) / 8 : (unsigned)src.w;
  dim3 g(min(cdiv(per_row, 256), 16u), src.h, nimages);
  k_convert_out<<<g, 256, 0, st>>>(src, dst, src_stride, dst_stride);
}
void b200k_swap_sheets(cudaStream_t st, DPage *pages, int npages) {
  if (npages <= 0) return;
  k_swap_sheets<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages);
}
void b200k_set_other(cudaStream_t st, DPage *pages, int npages, uint8_t *base, size_t stride) {
  if (npages <= 0) return;
  k_set_other<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, base, stride);
}
void b200k_set_geometry(cudaStream_t st, DPage *pages, int npages, int w, int h, int pitch) {
  if (npages <= 0) return;
  k_set_geometry<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages, w, h, pitch);
}
void b200k_retarget_jobs(cudaStream_t st, const DPage *pages, int npages, DFillJob *fills, int nfill,
                         DMaskJob *masks, int nmask, int stride) {
  int n = nfill > nmask ? nfill : nmask;
  if (npages <= 0 || n <= 0) return;
  k_retarget_jobs<<<cdiv(n, 128), 128, 0, st>>>(pages, npages, fills, nfill, masks, nmask, stride);
}
void b200k_page_reset(cudaStream_t st, DPage *pages, int npages) {
  if (npages <= 0) return;
  k_page_reset<<<cdiv(npages, 64), 64, 0, st>>>(pages, npages);
}
void b200k_pack_rows(cudaStream_t st, const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                     int row_bytes, int rows, int nimages, size_t src_stride, size_t dst_stride) {
  if (nimages <= 0 || rows <= 0 || row_bytes <= 0) return;
  dim3 g(min(cdiv(row_bytes, 16 * 256), 8u), min((unsigned)rows, 1024u), nimages);
  k_pack_rows<<<g, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, row_bytes, rows, src_stride, dst_stride);
}
}
