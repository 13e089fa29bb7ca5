/* rt.c — see rt.h */
#define _GNU_SOURCE
#include "rt.h"

#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "unpaper_b200.h"

static __thread char tls_error[512];

void b200_fatal(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "unpaper-b200: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  exit(1);
}

void b200_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

const char *unpaper_b200_last_error(void) { return tls_error; }
const char *unpaper_b200_version(void) { return "unpaper-b200 0.1 (sm_100a)"; }

/* ---- init ---------------------------------------------------------------- */

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_ndev = 0;
static UnpaperCudaInitStatus g_status = UNPAPER_CUDA_INIT_ERROR;

static void do_init(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e == cudaErrorNoDevice || (e == cudaSuccess && n == 0)) { g_status = UNPAPER_CUDA_INIT_NO_DEVICE; return; }
  if (e == cudaErrorInsufficientDriver) { g_status = UNPAPER_CUDA_INIT_NO_RUNTIME; return; }
  if (e != cudaSuccess) { g_status = UNPAPER_CUDA_INIT_ERROR; return; }
  g_ndev = n > B200_MAX_DEVICES ? B200_MAX_DEVICES : n;
  g_status = UNPAPER_CUDA_INIT_OK;
}

bool b200_rt_init(void) {
  pthread_once(&g_once, do_init);
  return g_status == UNPAPER_CUDA_INIT_OK;
}

UnpaperCudaInitStatus unpaper_cuda_try_init(void) {
  b200_rt_init();
  return g_status;
}

const char *unpaper_cuda_init_status_string(UnpaperCudaInitStatus st) {
  switch (st) {
  case UNPAPER_CUDA_INIT_OK: return "ok";
  case UNPAPER_CUDA_INIT_NO_RUNTIME: return "no CUDA driver/runtime";
  case UNPAPER_CUDA_INIT_NO_DEVICE: return "no CUDA device";
  default: return "CUDA initialisation error";
  }
}

/* ---- per-thread {device, stream} ---------------------------------------- */

static __thread int tls_device = -1;
static __thread cudaStream_t tls_own_stream[B200_MAX_DEVICES];
static __thread cudaStream_t tls_override = NULL;

int unpaper_b200_device_count(void) { return b200_rt_init() ? g_ndev : 0; }

int unpaper_b200_set_device(int device) {
  if (!b200_rt_init() || device < 0 || device >= g_ndev) { b200_set_error("invalid device %d", device); return -1; }
  CUDA_OK(cudaSetDevice(device));
  tls_device = device;
  return 0;
}

int b200_rt_device(void) {
  if (tls_device < 0) {
    if (!b200_rt_init()) b200_fatal("CUDA backend requested, but no usable GPU: %s", unpaper_cuda_init_status_string(g_status));
    CUDA_OK(cudaSetDevice(0));
    tls_device = 0;
  }
  return tls_device;
}
int unpaper_b200_get_device(void) { return b200_rt_device(); }

cudaStream_t b200_rt_stream(void) {
  if (tls_override) return tls_override;
  int d = b200_rt_device();
  if (!tls_own_stream[d]) CUDA_OK(cudaStreamCreateWithFlags(&tls_own_stream[d], cudaStreamNonBlocking));
  return tls_own_stream[d];
}
void b200_rt_set_stream(cudaStream_t s) { tls_override = s; }
void unpaper_b200_thread_sync(void) { CUDA_OK(cudaStreamSynchronize(b200_rt_stream())); }

/* ---- bucketed caches ------------------------------------------------------ */

#define NBUCKET 112
#define LIVE_SLOTS 4096   /* open-addressing table of the blocks handed out (power of two) */
typedef struct Block { struct Block *next; void *ptr; size_t bytes; int dev; int bucket; } Block;
typedef struct {
  pthread_mutex_t mu;
  Block *free_list[B200_MAX_DEVICES + 1][NBUCKET];
  Block *live[LIVE_SLOTS];   /* pointer -> block, linear probing; overflow falls back to the chain below */
  Block *live_overflow;
} Cache;
static Cache g_dev_cache = {PTHREAD_MUTEX_INITIALIZER};
static Cache g_pin_cache = {PTHREAD_MUTEX_INITIALIZER};

static int bucket_of(size_t bytes, size_t *rounded) {
  /* 4 buckets per power of two from 4 KiB up; beyond the table: exact size, not cached */
  size_t b = 4096;
  int k = 0;
  for (int p = 0; p < NBUCKET / 4; p++, b <<= 1)
    for (int q = 0; q < 4; q++, k++) {
      size_t sz = b + (b / 4) * (size_t)q;
      if (bytes <= sz) { *rounded = sz; return k; }
    }
  *rounded = bytes;
  return -1;
}

static unsigned live_hash(const void *p) { return (unsigned)(((uintptr_t)p >> 9) * 2654435761u) & (LIVE_SLOTS - 1); }
/* both with c->mu held */
static void live_put(Cache *c, Block *b) {
  unsigned h = live_hash(b->ptr);
  for (int i = 0; i < 64; i++, h = (h + 1) & (LIVE_SLOTS - 1))
    if (!c->live[h]) { c->live[h] = b; return; }
  b->next = c->live_overflow; c->live_overflow = b;
}
static Block *live_take(Cache *c, const void *p) {
  unsigned h = live_hash(p);
  for (int i = 0; i < 64; i++, h = (h + 1) & (LIVE_SLOTS - 1)) {
    Block *b = c->live[h];
    if (b && b->ptr == p) {
      /* keep probe chains intact: re-insert the run that follows */
      c->live[h] = NULL;
      unsigned j = (h + 1) & (LIVE_SLOTS - 1);
      while (c->live[j]) { Block *m = c->live[j]; c->live[j] = NULL; live_put(c, m); j = (j + 1) & (LIVE_SLOTS - 1); }
      return b;
    }
    if (!b) break;
  }
  for (Block **pp = &c->live_overflow; *pp; pp = &(*pp)->next)
    if ((*pp)->ptr == p) { Block *b = *pp; *pp = b->next; return b; }
  return NULL;
}

static void *cache_alloc(Cache *c, size_t bytes, int dev, bool pinned) {
  size_t rounded;
  int k = bucket_of(bytes, &rounded);
  pthread_mutex_lock(&c->mu);
  if (k >= 0) {
    Block *b = c->free_list[dev][k];
    if (b) {
      c->free_list[dev][k] = b->next;
      b->next = NULL;
      live_put(c, b);
      pthread_mutex_unlock(&c->mu);
      return b->ptr;
    }
  }
  pthread_mutex_unlock(&c->mu);
  void *p = NULL;
  if (pinned) CUDA_OK(cudaHostAlloc(&p, rounded, cudaHostAllocPortable));
  else CUDA_OK(cudaMalloc(&p, rounded));
  Block *b = (Block *)calloc(1, sizeof(*b));
  b->ptr = p; b->bytes = rounded; b->dev = dev; b->bucket = k;
  pthread_mutex_lock(&c->mu);
  live_put(c, b);
  pthread_mutex_unlock(&c->mu);
  return p;
}

static void cache_free(Cache *c, void *p, bool pinned) {
  if (!p) return;
  pthread_mutex_lock(&c->mu);
  Block *b = live_take(c, p);
  if (!b) { pthread_mutex_unlock(&c->mu); b200_fatal("free of unknown block %p", p); }
  if (b->bucket >= 0) {
    b->next = c->free_list[b->dev][b->bucket];
    c->free_list[b->dev][b->bucket] = b;
    pthread_mutex_unlock(&c->mu);
    return;
  }
  pthread_mutex_unlock(&c->mu);
  if (pinned) cudaFreeHost(b->ptr); else { int cur; cudaGetDevice(&cur); cudaSetDevice(b->dev); cudaFree(b->ptr); cudaSetDevice(cur); }
  free(b);
}

void *b200_dev_alloc(size_t bytes) { return cache_alloc(&g_dev_cache, bytes ? bytes : 1, b200_rt_device(), false); }
void b200_dev_free(void *p) { cache_free(&g_dev_cache, p, false); }
void *b200_pinned_alloc(size_t bytes) { b200_rt_device(); return cache_alloc(&g_pin_cache, bytes ? bytes : 1, B200_MAX_DEVICES, true); }
void b200_pinned_free(void *p) { cache_free(&g_pin_cache, p, true); }

void b200_rt_trim(void) {
  Cache *cs[2] = {&g_dev_cache, &g_pin_cache};
  for (int ci = 0; ci < 2; ci++) {
    Cache *c = cs[ci];
    pthread_mutex_lock(&c->mu);
    for (int d = 0; d <= B200_MAX_DEVICES; d++)
      for (int k = 0; k < NBUCKET; k++) {
        Block *b = c->free_list[d][k];
        c->free_list[d][k] = NULL;
        while (b) {
          Block *n = b->next;
          if (ci == 1) cudaFreeHost(b->ptr);
          else { int cur; cudaGetDevice(&cur); cudaSetDevice(b->dev); cudaFree(b->ptr); cudaSetDevice(cur); }
          free(b);
          b = n;
        }
      }
    pthread_mutex_unlock(&c->mu);
  }
}

/* ---- stream pool --------------------------------------------------------- */

#define POOL_STREAMS 32
static pthread_mutex_t g_sp_mu = PTHREAD_MUTEX_INITIALIZER;
static cudaStream_t g_sp[B200_MAX_DEVICES][POOL_STREAMS];
static int g_sp_n[B200_MAX_DEVICES];

cudaStream_t b200_stream_acquire(void) {
  int d = b200_rt_device();
  pthread_mutex_lock(&g_sp_mu);
  if (g_sp_n[d] > 0) {
    cudaStream_t s = g_sp[d][--g_sp_n[d]];
    pthread_mutex_unlock(&g_sp_mu);
    return s;
  }
  pthread_mutex_unlock(&g_sp_mu);
  cudaStream_t s;
  CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return s;
}

void b200_stream_release(cudaStream_t s) {
  if (!s) return;
  int d = b200_rt_device();
  pthread_mutex_lock(&g_sp_mu);
  if (g_sp_n[d] < POOL_STREAMS) { g_sp[d][g_sp_n[d]++] = s; pthread_mutex_unlock(&g_sp_mu); return; }
  pthread_mutex_unlock(&g_sp_mu);
  cudaStreamDestroy(s);
}
