/* pool.c — the page scheduler across GPUs: what reference lib/batch_worker.c:174-296
 * (batch_process_parallel: a pool of pthreads pulling job indices from a shared
 * counter, one stream per job) plus lib/decode_queue.h (a bounded queue of decoded
 * pages filled by producer threads, :54-56 custom decoder hook, :97-99 get) do for one
 * device, for any number of devices in one process.
 *
 * Per device: one sheet engine, one bounded ring of pinned slots (a slot = the pages
 * of up to `slot_sheets` sheets + room for their output), one producer thread that
 * claims the next run of job indices from the pool's shared counter and has the
 * caller's producer hook fill a free slot (decode), and one feeder thread that streams
 * ready slots into the engine (unpaper_b200_engine_stream_*) and hands finished sheets
 * to the caller's sink as their groups complete.  Sheets are independent, so nothing is
 * exchanged between devices; the shared counter balances the load.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>
#include <time.h>

#include "host.h"

static int g_trace = -1;
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec / 1e6; }
#define TRACE(...) do { if (g_trace < 0) g_trace = getenv("UNPAPER_B200_POOL_TRACE") != NULL; if (g_trace) { fprintf(stderr, "[pool %9.3f] ", now_ms()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)

typedef struct {
  uint8_t *in, *out;     /* pinned */
  int first, n;          /* job indices [first, first + n) */
  int reported;          /* sheets already handed to the sink */
  int state;             /* 0 free, 1 being filled, 2 ready, 3 in the engine */
} Slot;

typedef struct PoolDev {
  struct B200Pool *pool;
  int device, index;
  B200Engine *eng;
  Slot *slots;
  int nslots;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int *ready;            /* FIFO of ready slot indices */
  int ready_head, ready_tail, producer_done;
  pthread_t producer, feeder;
  int rc;
  uint64_t sheets_done;
} PoolDev;

struct B200Pool {
  int ndev, slot_sheets;
  size_t sheet_in_bytes, sheet_out_bytes;
  PoolDev *dev;
  /* one run */
  atomic_int next;
  int n_sheets;
  atomic_int failed;
  B200PageProducerFn produce; void *produce_user;
  B200PoolSheetFn sink; void *sink_user;
  B200SheetResult *results;
};

int unpaper_b200_pool_device_count(const B200Pool *p) { return p ? p->ndev : 0; }
size_t unpaper_b200_pool_sheet_bytes(const B200Pool *p) { return p ? p->sheet_out_bytes : 0; }
uint64_t unpaper_b200_pool_sheets_done(const B200Pool *p, int i) { return p && i >= 0 && i < p->ndev ? p->dev[i].sheets_done : 0; }
B200Engine *unpaper_b200_pool_engine(B200Pool *p, int i) { return p && i >= 0 && i < p->ndev ? p->dev[i].eng : NULL; }

void unpaper_b200_pool_destroy(B200Pool *p) {
  if (!p) return;
  for (int i = 0; i < p->ndev; i++) {
    PoolDev *d = &p->dev[i];
    unpaper_b200_set_device(d->device);
    if (d->eng) unpaper_b200_engine_destroy(d->eng);
    for (int k = 0; d->slots && k < d->nslots; k++) {
      if (d->slots[k].in) b200_pinned_free(d->slots[k].in);
      if (d->slots[k].out) b200_pinned_free(d->slots[k].out);
    }
    free(d->slots); free(d->ready);
    pthread_mutex_destroy(&d->mu); pthread_cond_destroy(&d->cv);
  }
  free(p->dev);
  free(p);
}

B200Pool *unpaper_b200_pool_create(const B200SheetConfig *cfg, const int *devices, int n_devices, int page_w, int page_h,
                                   int page_format, int group_pages, int lanes, int slot_sheets, int slots_per_device) {
  if (!cfg || n_devices <= 0 || n_devices > B200_MAX_DEVICES || group_pages <= 0 || lanes <= 0) { b200_set_error("pool: bad arguments"); return NULL; }
  if (slot_sheets <= 0) slot_sheets = group_pages;
  /* every group in flight pins its slot until it is collected: two flights per lane, plus
   * one slot being filled and one waiting */
  int min_slots = (2 * lanes * group_pages + slot_sheets - 1) / slot_sheets + 2;
  if (slots_per_device < min_slots) slots_per_device = min_slots;
  B200Pool *p = (B200Pool *)calloc(1, sizeof(*p));
  p->ndev = n_devices; p->slot_sheets = slot_sheets;
  p->dev = (PoolDev *)calloc((size_t)n_devices, sizeof(PoolDev));
  for (int i = 0; i < n_devices; i++) {
    PoolDev *d = &p->dev[i];
    d->pool = p; d->index = i; d->device = devices ? devices[i] : i;
    pthread_mutex_init(&d->mu, NULL); pthread_cond_init(&d->cv, NULL);
    if (unpaper_b200_set_device(d->device)) { unpaper_b200_pool_destroy(p); return NULL; }
    d->eng = unpaper_b200_engine_create(cfg, d->device, page_w, page_h, page_format, group_pages, lanes);
    if (!d->eng) { unpaper_b200_pool_destroy(p); return NULL; }
    p->sheet_in_bytes = (size_t)b200_fmt_row_bytes(page_format, page_w) * page_h * cfg->input_count;
    p->sheet_out_bytes = unpaper_b200_engine_sheet_bytes(d->eng);
    d->nslots = slots_per_device;
    d->slots = (Slot *)calloc((size_t)d->nslots, sizeof(Slot));
    d->ready = (int *)calloc((size_t)d->nslots + 1, sizeof(int));
    for (int k = 0; k < d->nslots; k++) {
      d->slots[k].in = (uint8_t *)b200_pinned_alloc(p->sheet_in_bytes * slot_sheets);
      d->slots[k].out = (uint8_t *)b200_pinned_alloc(p->sheet_out_bytes * slot_sheets);
    }
  }
  return p;
}

/* producer: claim job indices, let the caller decode them into a free pinned slot */
static void *producer_main(void *arg) {
  PoolDev *d = (PoolDev *)arg;
  B200Pool *p = d->pool;
  for (;;) {
    pthread_mutex_lock(&d->mu);
    int k = -1;
    for (;;) {
      for (int i = 0; i < d->nslots; i++) if (d->slots[i].state == 0) { k = i; break; }
      if (k >= 0) break;
      pthread_cond_wait(&d->cv, &d->mu);
    }
    d->slots[k].state = 1;
    pthread_mutex_unlock(&d->mu);
    int first = atomic_fetch_add(&p->next, p->slot_sheets);
    int n = first < p->n_sheets ? (p->n_sheets - first < p->slot_sheets ? p->n_sheets - first : p->slot_sheets) : 0;
    int got = 0;
    for (; got < n; got++) {
      int rc = p->produce(p->produce_user, first + got, d->slots[k].in + p->sheet_in_bytes * (size_t)got);
      if (rc != 0) {                    /* decode failure or end of input: nothing after it is processed */
        if (rc < 0) atomic_fetch_add(&p->failed, 1);
        atomic_store(&p->next, p->n_sheets);
        break;
      }
    }
    TRACE("dev %d produced slot %d first %d got %d", d->device, k, first, got);
    pthread_mutex_lock(&d->mu);
    if (got > 0) {
      d->slots[k].first = first; d->slots[k].n = got; d->slots[k].reported = 0; d->slots[k].state = 2;
      d->ready[d->ready_tail] = k; d->ready_tail = (d->ready_tail + 1) % (d->nslots + 1);
    } else d->slots[k].state = 0;
    bool done = got < p->slot_sheets || first + got >= p->n_sheets;
    if (done) d->producer_done = 1;
    pthread_cond_broadcast(&d->cv);
    pthread_mutex_unlock(&d->mu);
    if (done) return NULL;
  }
}

/* engine callback: a sheet of this device is complete (runs on the feeder thread) */
static int on_sheet(void *user, int sheet_index, const uint8_t *sheet, const B200SheetResult *res) {
  PoolDev *d = (PoolDev *)user;
  B200Pool *p = d->pool;
  int rc = p->sink ? p->sink(p->sink_user, sheet_index, d->device, sheet, res) : 0;
  d->sheets_done++;
  pthread_mutex_lock(&d->mu);
  for (int k = 0; k < d->nslots; k++) {
    Slot *s = &d->slots[k];
    if (s->state == 3 && sheet_index >= s->first && sheet_index < s->first + s->n) {
      if (++s->reported == s->n) { s->state = 0; pthread_cond_broadcast(&d->cv); }   /* the slot's buffers are free again */
      break;
    }
  }
  pthread_mutex_unlock(&d->mu);
  return rc;
}

static void *feeder_main(void *arg) {
  PoolDev *d = (PoolDev *)arg;
  B200Pool *p = d->pool;
  unpaper_b200_set_device(d->device);
  unpaper_b200_engine_set_sheet_callback(d->eng, on_sheet, d);
  if (unpaper_b200_engine_stream_begin(d->eng, 1)) { d->rc = -1; return NULL; }
  for (;;) {
    int k = -1, finished = 0;
    pthread_mutex_lock(&d->mu);
    for (;;) {
      if (d->ready_head != d->ready_tail) { k = d->ready[d->ready_head]; d->ready_head = (d->ready_head + 1) % (d->nslots + 1); break; }
      if (d->producer_done) { finished = 1; break; }
      /* nothing ready: finish the oldest group in flight instead of sleeping — that also
       * frees the slot the producer may be waiting for */
      if (unpaper_b200_engine_stream_in_flight(d->eng) > 0) break;
      pthread_cond_wait(&d->cv, &d->mu);
    }
    pthread_mutex_unlock(&d->mu);
    if (finished) break;
    if (k < 0) { TRACE("dev %d poll (nothing ready)", d->device); unpaper_b200_engine_stream_poll(d->eng); TRACE("dev %d poll done", d->device); continue; }
    Slot *s = &d->slots[k];
    pthread_mutex_lock(&d->mu);
    s->state = 3;
    pthread_mutex_unlock(&d->mu);
    TRACE("dev %d feed slot %d first %d n %d in flight %d", d->device, k, s->first, s->n, unpaper_b200_engine_stream_in_flight(d->eng));
    if (unpaper_b200_engine_stream_feed(d->eng, s->in, s->out, s->n, p->results ? p->results + s->first : NULL, s->first)) { d->rc = -1; break; }
    TRACE("dev %d fed", d->device);
  }
  int rc = unpaper_b200_engine_stream_end(d->eng);
  if (rc && !d->rc) d->rc = rc;
  unpaper_b200_engine_set_sheet_callback(d->eng, NULL, NULL);
  return NULL;
}

int unpaper_b200_pool_run(B200Pool *p, int n_sheets, B200PageProducerFn produce, void *produce_user,
                          B200PoolSheetFn sink, void *sink_user, B200SheetResult *results) {
  if (!p || n_sheets < 0 || !produce) { b200_set_error("pool: bad arguments"); return -1; }
  p->n_sheets = n_sheets; p->produce = produce; p->produce_user = produce_user;
  p->sink = sink; p->sink_user = sink_user; p->results = results;
  atomic_store(&p->next, 0); atomic_store(&p->failed, 0);
  for (int i = 0; i < p->ndev; i++) {
    PoolDev *d = &p->dev[i];
    d->ready_head = d->ready_tail = 0; d->producer_done = 0; d->rc = 0; d->sheets_done = 0;
    for (int k = 0; k < d->nslots; k++) d->slots[k].state = 0;
    pthread_create(&d->producer, NULL, producer_main, d);
    pthread_create(&d->feeder, NULL, feeder_main, d);
  }
  int rc = 0;
  for (int i = 0; i < p->ndev; i++) {
    pthread_join(p->dev[i].producer, NULL);
    pthread_join(p->dev[i].feeder, NULL);
    if (p->dev[i].rc && !rc) { rc = p->dev[i].rc; b200_set_error("pool: device %d failed (%d)", p->dev[i].device, rc); }
  }
  if (!rc && atomic_load(&p->failed)) { b200_set_error("pool: the page producer failed"); rc = -4; }
  return rc;
}
