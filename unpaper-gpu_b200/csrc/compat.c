/* compat.c — the reference's CUDA runtime symbol names over the B200 runtime (rt.c).
 *
 * When the reference tree is built with -DUNPAPER_WITH_CUDA=1, its L3/L4 code
 * calls a handful of runtime helpers besides the vtable:
 *   lib/perf.c:39,63,86              unpaper_cuda_events_supported / event_pair_start /
 *                                    event_pair_stop_ms        (cuda_runtime.h:86-88)
 *   lib/batch_worker.c:198-255       cuda_stream_pool_global_active/_acquire/_release,
 *                                    unpaper_cuda_set_current_stream, event_pair_*_on
 *   src/pipeline/image_pipeline.c    cuda_stream_pool_global_init/_cleanup/_print_stats,
 *                                    cuda_mempool_*global_init/_cleanup/_print_stats
 * They are exported here with the reference's signatures (imageprocess/cuda_runtime.h,
 * cuda_stream_pool.h, cuda_mempool.h) so that the reference links against
 * libunpaper_b200.so unchanged.  The memory "pools" are views of rt.c's bucketed
 * cache (nothing is pre-carved: a cache hit is what the reference calls a pool hit).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt.h"
#include "unpaper_b200.h"

struct UnpaperCudaStream { cudaStream_t s; int device; int pooled; };

static __thread UnpaperCudaStream *tls_current = NULL;

/* ---- streams (cuda_runtime.h:62-67,93) ------------------------------------ */

UnpaperCudaStream *unpaper_cuda_stream_create(void) {
  if (!b200_rt_init()) return NULL;
  UnpaperCudaStream *st = (UnpaperCudaStream *)calloc(1, sizeof(*st));
  if (!st) return NULL;
  st->device = b200_rt_device();
  st->s = b200_stream_acquire();
  return st;
}
void unpaper_cuda_stream_destroy(UnpaperCudaStream *st) {
  if (!st) return;
  if (tls_current == st) { tls_current = NULL; b200_rt_set_stream(NULL); }
  if (st->s) { cudaStreamSynchronize(st->s); b200_stream_release(st->s); }
  free(st);
}
UnpaperCudaStream *unpaper_cuda_stream_get_default(void) {
  static __thread UnpaperCudaStream dflt;
  b200_rt_set_stream(NULL);
  dflt.s = b200_rt_stream(); dflt.device = b200_rt_device();
  if (tls_current) b200_rt_set_stream(tls_current->s);
  return &dflt;
}
/* per-thread current stream (cuda_runtime.c:70,616-626): every vtable op of this
 * thread is enqueued on it; NULL restores the thread's own stream */
void unpaper_cuda_set_current_stream(UnpaperCudaStream *st) {
  tls_current = st;
  b200_rt_set_stream(st ? st->s : NULL);
}
UnpaperCudaStream *unpaper_cuda_get_current_stream(void) { return tls_current; }
void unpaper_cuda_stream_synchronize(void) { CUDA_OK(cudaStreamSynchronize(b200_rt_stream())); }
void unpaper_cuda_stream_synchronize_on(UnpaperCudaStream *st) {
  CUDA_OK(cudaStreamSynchronize(st ? st->s : b200_rt_stream()));
}
void *unpaper_cuda_stream_get_raw_handle(UnpaperCudaStream *st) { return st ? (void *)st->s : (void *)b200_rt_stream(); }

/* ---- raw device memory (cuda_runtime.h:27-47,69-71): what a GPU decoder uses to
 * produce the buffer it hands to create_image_from_gpu() ----------------------- */

uint64_t unpaper_cuda_malloc(size_t bytes) {
  b200_rt_device();
  void *p = NULL;
  CUDA_OK(cudaMalloc(&p, bytes ? bytes : 1));
  return (uint64_t)(uintptr_t)p;
}
void unpaper_cuda_free(uint64_t dptr) {
  if (dptr) CUDA_OK(cudaFree((void *)(uintptr_t)dptr));
}
void unpaper_cuda_memcpy_h2d(uint64_t dst, const void *src, size_t bytes) {
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemcpyAsync((void *)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaStreamSynchronize(s));
}
void unpaper_cuda_memcpy_d2h(void *dst, uint64_t src, size_t bytes) {
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemcpyAsync(dst, (const void *)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
}
void unpaper_cuda_memcpy_d2d(uint64_t dst, uint64_t src, size_t bytes) {
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemcpyAsync((void *)(uintptr_t)dst, (const void *)(uintptr_t)src, bytes, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaStreamSynchronize(s));
}
void unpaper_cuda_memcpy_h2d_async(UnpaperCudaStream *st, uint64_t dst, const void *src, size_t bytes) {
  CUDA_OK(cudaMemcpyAsync((void *)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, st ? st->s : b200_rt_stream()));
}
void unpaper_cuda_memcpy_d2h_async(UnpaperCudaStream *st, void *dst, uint64_t src, size_t bytes) {
  CUDA_OK(cudaMemcpyAsync(dst, (const void *)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, st ? st->s : b200_rt_stream()));
}
void unpaper_cuda_memcpy_d2d_async(UnpaperCudaStream *st, uint64_t dst, uint64_t src, size_t bytes) {
  CUDA_OK(cudaMemcpyAsync((void *)(uintptr_t)dst, (const void *)(uintptr_t)src, bytes, cudaMemcpyDeviceToDevice,
                          st ? st->s : b200_rt_stream()));
}
void unpaper_cuda_memset_d8(uint64_t dst, uint8_t value, size_t bytes) {
  cudaStream_t s = b200_rt_stream();
  CUDA_OK(cudaMemsetAsync((void *)(uintptr_t)dst, value, bytes, s));
  CUDA_OK(cudaStreamSynchronize(s));
}
void unpaper_cuda_memset_async(UnpaperCudaStream *st, uint64_t dst, uint8_t value, size_t bytes) {
  CUDA_OK(cudaMemsetAsync((void *)(uintptr_t)dst, value, bytes, st ? st->s : b200_rt_stream()));
}

/* ---- event pairs (cuda_runtime.h:86-92; used by lib/perf.c) ----------------- */

bool unpaper_cuda_events_supported(void) { return b200_rt_init(); }
bool unpaper_cuda_events_supported_on(UnpaperCudaStream *st) { (void)st; return b200_rt_init(); }

static bool pair_start(cudaStream_t s, void **start, void **stop) {
  if (!start || !stop || !b200_rt_init()) return false;
  cudaEvent_t a = NULL, b = NULL;
  if (cudaEventCreate(&a) != cudaSuccess) return false;
  if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return false; }
  if (cudaEventRecord(a, s) != cudaSuccess) { cudaEventDestroy(a); cudaEventDestroy(b); return false; }
  *start = a; *stop = b;
  return true;
}
static double pair_stop(cudaStream_t s, void **start, void **stop) {
  if (!start || !stop || !*start || !*stop) return 0.0;
  cudaEvent_t a = (cudaEvent_t)*start, b = (cudaEvent_t)*stop;
  float ms = 0.0f;
  if (cudaEventRecord(b, s) == cudaSuccess && cudaEventSynchronize(b) == cudaSuccess) cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  *start = NULL; *stop = NULL;
  return (double)ms;
}
bool unpaper_cuda_event_pair_start(void **start, void **stop) { return pair_start(b200_rt_stream(), start, stop); }
double unpaper_cuda_event_pair_stop_ms(void **start, void **stop) { return pair_stop(b200_rt_stream(), start, stop); }
bool unpaper_cuda_event_pair_start_on(UnpaperCudaStream *st, void **start, void **stop) {
  return pair_start(st ? st->s : b200_rt_stream(), start, stop);
}
double unpaper_cuda_event_pair_stop_ms_on(UnpaperCudaStream *st, void **start, void **stop) {
  return pair_stop(st ? st->s : b200_rt_stream(), start, stop);
}

/* ---- stream pool (cuda_stream_pool.h) --------------------------------------- */

struct CudaStreamPool {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  UnpaperCudaStream **all;
  int *busy;
  size_t n;
  CudaStreamPoolStats stats;
};

CudaStreamPool *cuda_stream_pool_create(size_t stream_count) {
  if (stream_count == 0 || !b200_rt_init()) return NULL;
  CudaStreamPool *p = (CudaStreamPool *)calloc(1, sizeof(*p));
  if (!p) return NULL;
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv, NULL);
  p->all = (UnpaperCudaStream **)calloc(stream_count, sizeof(*p->all));
  p->busy = (int *)calloc(stream_count, sizeof(int));
  for (size_t i = 0; i < stream_count; i++) { p->all[i] = unpaper_cuda_stream_create(); if (p->all[i]) p->all[i]->pooled = 1; }
  p->n = stream_count;
  p->stats.stream_count = stream_count;
  return p;
}
void cuda_stream_pool_destroy(CudaStreamPool *p) {
  if (!p) return;
  for (size_t i = 0; i < p->n; i++) unpaper_cuda_stream_destroy(p->all[i]);
  free(p->all); free(p->busy);
  pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv);
  free(p);
}
UnpaperCudaStream *cuda_stream_pool_acquire(CudaStreamPool *p) {
  if (!p) return NULL;
  pthread_mutex_lock(&p->mu);
  p->stats.total_acquisitions++;
  bool waited = false;
  for (;;) {
    for (size_t i = 0; i < p->n; i++)
      if (!p->busy[i] && p->all[i]) {
        p->busy[i] = 1;
        if (++p->stats.current_in_use > p->stats.peak_in_use) p->stats.peak_in_use = p->stats.current_in_use;
        pthread_mutex_unlock(&p->mu);
        return p->all[i];
      }
    if (!waited) { p->stats.waits++; waited = true; }
    pthread_cond_wait(&p->cv, &p->mu);
  }
}
void cuda_stream_pool_release(CudaStreamPool *p, UnpaperCudaStream *st) {
  if (!p || !st) return;
  pthread_mutex_lock(&p->mu);
  for (size_t i = 0; i < p->n; i++)
    if (p->all[i] == st && p->busy[i]) { p->busy[i] = 0; p->stats.current_in_use--; break; }
  pthread_cond_signal(&p->cv);
  pthread_mutex_unlock(&p->mu);
}
CudaStreamPoolStats cuda_stream_pool_get_stats(const CudaStreamPool *p) {
  CudaStreamPoolStats z;
  memset(&z, 0, sizeof(z));
  return p ? p->stats : z;
}
void cuda_stream_pool_print_stats(const CudaStreamPool *p) {
  if (!p) return;
  fprintf(stderr, "CUDA stream pool: %zu streams, %zu acquisitions, %zu waits, peak %zu in use\n",
          p->stats.stream_count, p->stats.total_acquisitions, p->stats.waits, p->stats.peak_in_use);
}

static CudaStreamPool *g_pool = NULL;
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
bool cuda_stream_pool_global_init(size_t stream_count) {
  pthread_mutex_lock(&g_pool_mu);
  if (!g_pool) g_pool = cuda_stream_pool_create(stream_count);
  bool ok = g_pool != NULL;
  pthread_mutex_unlock(&g_pool_mu);
  return ok;
}
void cuda_stream_pool_global_cleanup(void) {
  pthread_mutex_lock(&g_pool_mu);
  CudaStreamPool *p = g_pool; g_pool = NULL;
  pthread_mutex_unlock(&g_pool_mu);
  cuda_stream_pool_destroy(p);
}
bool cuda_stream_pool_global_active(void) { return g_pool != NULL; }
UnpaperCudaStream *cuda_stream_pool_global_acquire(void) { return g_pool ? cuda_stream_pool_acquire(g_pool) : NULL; }
void cuda_stream_pool_global_release(UnpaperCudaStream *st) { if (g_pool) cuda_stream_pool_release(g_pool, st); }
CudaStreamPoolStats cuda_stream_pool_global_get_stats(void) { return cuda_stream_pool_get_stats(g_pool); }
void cuda_stream_pool_global_print_stats(void) { cuda_stream_pool_print_stats(g_pool); }

/* ---- memory pools (cuda_mempool.h) ------------------------------------------- */

struct CudaMemPool { pthread_mutex_t mu; CudaMemPoolStats stats; };

CudaMemPool *cuda_mempool_create(size_t buffer_count, size_t buffer_size) {
  if (!b200_rt_init()) return NULL;
  CudaMemPool *p = (CudaMemPool *)calloc(1, sizeof(*p));
  if (!p) return NULL;
  pthread_mutex_init(&p->mu, NULL);
  p->stats.buffer_count = buffer_count; p->stats.buffer_size = buffer_size;
  /* warm the cache: the first `buffer_count` acquisitions of this size are hits */
  void **tmp = (void **)calloc(buffer_count ? buffer_count : 1, sizeof(void *));
  for (size_t i = 0; i < buffer_count; i++) tmp[i] = b200_dev_alloc(buffer_size);
  for (size_t i = 0; i < buffer_count; i++) b200_dev_free(tmp[i]);
  free(tmp);
  p->stats.total_bytes_pooled = buffer_count * buffer_size;
  return p;
}
void cuda_mempool_destroy(CudaMemPool *p) {
  if (!p) return;
  pthread_mutex_destroy(&p->mu);
  free(p);
}
uint64_t cuda_mempool_acquire(CudaMemPool *p, size_t bytes) {
  void *d = b200_dev_alloc(bytes);
  if (p) {
    pthread_mutex_lock(&p->mu);
    p->stats.total_allocations++;
    if (bytes <= p->stats.buffer_size) p->stats.pool_hits++; else { p->stats.pool_misses++; p->stats.size_mismatches++; }
    if (++p->stats.current_in_use > p->stats.peak_in_use) p->stats.peak_in_use = p->stats.current_in_use;
    pthread_mutex_unlock(&p->mu);
  }
  return (uint64_t)(uintptr_t)d;
}
void cuda_mempool_release(CudaMemPool *p, uint64_t dptr) {
  if (!dptr) return;
  b200_dev_free((void *)(uintptr_t)dptr);
  if (p) { pthread_mutex_lock(&p->mu); if (p->stats.current_in_use) p->stats.current_in_use--; pthread_mutex_unlock(&p->mu); }
}
CudaMemPoolStats cuda_mempool_get_stats(const CudaMemPool *p) {
  CudaMemPoolStats z;
  memset(&z, 0, sizeof(z));
  return p ? p->stats : z;
}
void cuda_mempool_print_stats(const CudaMemPool *p) {
  if (!p) return;
  fprintf(stderr, "GPU memory pool: %zu x %zu bytes, %zu acquisitions, %zu hits, %zu misses, peak %zu in use\n",
          p->stats.buffer_count, p->stats.buffer_size, p->stats.total_allocations, p->stats.pool_hits,
          p->stats.pool_misses, p->stats.peak_in_use);
}

#define GLOBAL_POOL(NAME, VAR)                                                               \
  static CudaMemPool *VAR = NULL;                                                            \
  bool cuda_mempool_##NAME##global_init(size_t n, size_t sz) {                               \
    if (!VAR) VAR = cuda_mempool_create(n, sz);                                              \
    return VAR != NULL;                                                                      \
  }                                                                                          \
  void cuda_mempool_##NAME##global_cleanup(void) { cuda_mempool_destroy(VAR); VAR = NULL; } \
  bool cuda_mempool_##NAME##global_active(void) { return VAR != NULL; }                      \
  uint64_t cuda_mempool_##NAME##global_acquire(size_t bytes) { return VAR ? cuda_mempool_acquire(VAR, bytes) : 0; } \
  void cuda_mempool_##NAME##global_release(uint64_t d) { if (VAR) cuda_mempool_release(VAR, d); } \
  CudaMemPoolStats cuda_mempool_##NAME##global_get_stats(void) { return cuda_mempool_get_stats(VAR); } \
  void cuda_mempool_##NAME##global_print_stats(void) { cuda_mempool_print_stats(VAR); }
GLOBAL_POOL(, g_mp_image)
GLOBAL_POOL(integral_, g_mp_integral)
GLOBAL_POOL(scratch_, g_mp_scratch)
