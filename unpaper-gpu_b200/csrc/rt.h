/* rt.h — device runtime of the B200 backend (internal).
 *
 * Replaces reference imageprocess/cuda_runtime.c (driver-API PTX loader,
 * device 0 hard-wired), cuda_mempool.c (three fixed-slot global pools) and
 * cuda_stream_pool.c.  Here: kernels are linked AOT through the runtime API;
 * the current {device, stream} is per thread; device and pinned memory come
 * from size-bucketed caches per device, so the steady state does no
 * cudaMalloc/cudaFree (the serialisation the reference's own post-mortem
 * blames for its 1.8x stream scaling, doc/CUDA_BACKEND_HISTORY.md:837-848).
 */
#pragma once
#include <cuda_runtime_api.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MAX_DEVICES 16

/* fatal error, reference convention (lib/logging.c:129-141): print + exit(1) */
void b200_fatal(const char *fmt, ...) __attribute__((format(printf, 1, 2), noreturn));
void b200_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

#define CUDA_OK(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess)                                                    \
      b200_fatal("CUDA failure %s at %s:%d: %s", #call, __FILE__, __LINE__,    \
                 cudaGetErrorString(e__));                                     \
  } while (0)

bool b200_rt_init(void);                 /* idempotent; false if no usable GPU */
int b200_rt_device(void);                /* calling thread's device */
cudaStream_t b200_rt_stream(void);       /* calling thread's stream on that device */
void b200_rt_set_stream(cudaStream_t s); /* override (engine lanes); NULL = default per-thread */

/* cached allocations; sizes are rounded up to a bucket */
void *b200_dev_alloc(size_t bytes);
void b200_dev_free(void *p);
void *b200_pinned_alloc(size_t bytes);
void b200_pinned_free(void *p);
void b200_rt_trim(void);                 /* release every cached block */

/* stream pool (reference cuda_stream_pool.h): n non-blocking streams per device */
cudaStream_t b200_stream_acquire(void);
void b200_stream_release(cudaStream_t s);

#ifdef __cplusplus
}
#endif
