/*
 * schro_host_core.c -- host-side plumbing of the drop-in layer: logging, memory
 * domains, frame allocation and the per-thread staging context.
 *
 * Mirrors (own implementation, same behaviour):
 *   schro_memory_domain_*             schroedinger/schrodomain.c:17-136
 *   schro_memory_domain_new_cuda      schroedinger/schrocuda.c:60-71
 *   schro_frame_new_and_alloc_full    schroedinger/schroframe.c:60-191
 *   schro_frame_ref / unref           schroedinger/schroframe.c:748-813
 *   schro_frame_to_gpu / gpuframe_to_cpu  schroedinger/schrogpuframe.c:480-609
 */
#include "schro_host.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

void
sb2h_fatal (const char *func, const char *fmt, ...)
{
  va_list ap;
  fprintf (stderr, "SCHRO-B200 ERROR: %s: ", func);
  va_start (ap, fmt);
  vfprintf (stderr, fmt, ap);
  va_end (ap);
  fprintf (stderr, "\n");
  abort ();
}

void
schro_init (void)
{
  /* nothing to JIT: the kernels are compiled for sm_100a ahead of time.  Touch the
   * runtime so that a missing GPU is reported here and not in the first picture. */
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess || n == 0)
    sb2h_fatal (__func__, "no CUDA device: the B200 picture core has no CPU fallback");
}

int
sb2h_mem_kind (const void *ptr)
{
  struct cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes (&at, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError ();
    return SB2H_MEM_PAGEABLE;
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)
    return SB2H_MEM_DEVICE;
  if (at.type == cudaMemoryTypeHost)
    return SB2H_MEM_PINNED;
  return SB2H_MEM_PAGEABLE;
}

/* ---- per-thread context -------------------------------------------------- */
static __thread Sb2hContext *tl_cx;
/* how sb2h_sync waits: 0 = cudaEventSynchronize on a blocking-sync event, N > 0 = poll with N-microsecond naps
 * (SB2_HOST_WAIT_US; default 0: with as many workers as cores the naps cost more than the wake-ups) */
static int g_wait_mode = 0;
/* The device of the first thread that enters the library becomes the process's device:
 * worker threads created later start on device 0 by CUDA's rules, which is wrong for
 * rank > 0 of a one-process-per-GPU job.  schro_b200_set_device overrides it. */
static int g_device = -1;
static pthread_mutex_t g_device_mutex = PTHREAD_MUTEX_INITIALIZER;

void
schro_b200_set_device (int device)
{
  pthread_mutex_lock (&g_device_mutex);
  g_device = device;
  pthread_mutex_unlock (&g_device_mutex);
  SB2H_CUDA (cudaSetDevice (device));
}

static void sb2h_pool_release_all (void);

#define SB2H_MAX_CTX 128
static Sb2hContext *g_ctx[SB2H_MAX_CTX];        /* live per-thread contexts, under g_device_mutex */
/* per-thread spare region of the in-place transforms (see sb2h_spare_take below) */
static struct { SchroMemoryDomain *domain; void *ptr; int size; } g_spare[SB2H_MAX_CTX];
static pthread_mutex_t g_spare_mutex = PTHREAD_MUTEX_INITIALIZER;

/* Give the calling thread's stream, staging buffers and pooled device blocks back.  Worker
 * threads are expected to be long-lived (as SchroAsync's are); a thread that does exit
 * should call this first.  There is deliberately no automatic thread-exit hook: it would
 * also run during process teardown, after the CUDA runtime has been unloaded. */
void
schro_b200_thread_sync (void)
{
  if (tl_cx) sb2h_sync (tl_cx);
}

void
schro_b200_thread_release (void)
{
  Sb2hContext *cx = tl_cx;
  int i, last;
  if (!cx) return;
  cudaStreamSynchronize (cx->stream);
  {
    /* the thread's spare transform region goes back to its domain (everything on the stream has finished) */
    SchroMemoryDomain *d = NULL;
    void *p = NULL;
    pthread_mutex_lock (&g_spare_mutex);
    if (g_spare[cx->slot].ptr) { d = g_spare[cx->slot].domain; p = g_spare[cx->slot].ptr; g_spare[cx->slot].ptr = NULL; }
    pthread_mutex_unlock (&g_spare_mutex);
    if (p) schro_memory_domain_memfree (d, p);
  }
  pthread_mutex_lock (&g_device_mutex);
  g_ctx[cx->slot] = NULL;
  for (i = 0, last = 1; i < SB2H_MAX_CTX; i++)
    if (g_ctx[i]) last = 0;
  pthread_mutex_unlock (&g_device_mutex);
  for (i = 0; i < SB2H_NBUF; i++)
    if (cx->dev[i]) cudaFree (cx->dev[i]);
  if (cx->pin) cudaFreeHost (cx->pin);
  if (cx->pin_ev) cudaEventDestroy (cx->pin_ev);
  if (last) sb2h_pool_release_all ();
  cudaEventDestroy (cx->sync_ev);
  cudaEventDestroy (cx->ev_fork);
  cudaEventDestroy (cx->ev_join);
  if (cx->stream_hi != cx->stream) cudaStreamDestroy (cx->stream_hi);
  cudaStreamDestroy (cx->stream);
  free (cx);
  tl_cx = NULL;
}

Sb2hContext *
sb2h_context (void)
{
  if (!tl_cx) {
    int i;
    pthread_mutex_lock (&g_device_mutex);
    if (g_device < 0) SB2H_CUDA (cudaGetDevice (&g_device));
    if (getenv ("SB2_HOST_WAIT_US")) g_wait_mode = atoi (getenv ("SB2_HOST_WAIT_US"));
    pthread_mutex_unlock (&g_device_mutex);
    SB2H_CUDA (cudaSetDevice (g_device));
    tl_cx = calloc (1, sizeof (Sb2hContext));
    SB2H_CUDA (cudaStreamCreateWithFlags (&tl_cx->stream, cudaStreamNonBlocking));
    SB2H_CUDA (cudaEventCreateWithFlags (&tl_cx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    {
      int lo = 0, hi = 0;
      SB2H_CUDA (cudaDeviceGetStreamPriorityRange (&lo, &hi));
      /* SB2_HOST_ONE_STREAM=1: the wavefront kernels share the thread's ordering stream (an experiment knob: with
       * more streams than the device has hardware queues, streams that share a queue wait on each other) */
      if (getenv ("SB2_HOST_ONE_STREAM") && atoi (getenv ("SB2_HOST_ONE_STREAM"))) tl_cx->stream_hi = tl_cx->stream;
      else SB2H_CUDA (cudaStreamCreateWithPriority (&tl_cx->stream_hi, cudaStreamNonBlocking, hi));
      SB2H_CUDA (cudaEventCreateWithFlags (&tl_cx->ev_fork, cudaEventDisableTiming));
      SB2H_CUDA (cudaEventCreateWithFlags (&tl_cx->ev_join, cudaEventDisableTiming));
    }
    pthread_mutex_lock (&g_device_mutex);
    for (i = 0; i < SB2H_MAX_CTX && g_ctx[i]; i++) ;
    if (i == SB2H_MAX_CTX) sb2h_fatal (__func__, "more than %d threads use the library", SB2H_MAX_CTX);
    g_ctx[i] = tl_cx;
    tl_cx->slot = i;
    pthread_mutex_unlock (&g_device_mutex);
  }
  return tl_cx;
}

void
sb2h_sync (Sb2hContext *cx)
{
  /* a worker that waits gives its core away (many workers share the host with the
   * application's own threads); the price is a few tens of microseconds of wake-up latency */
  SB2H_CUDA (cudaEventRecord (cx->sync_ev, cx->stream));
  if (g_wait_mode == 0) {
    SB2H_CUDA (cudaEventSynchronize (cx->sync_ev));
  } else {
    /* poll: a few tens of microseconds of spinning for the short waits, then short sleeps.  The
     * driver's interrupt-driven wait costs ~0.4 ms per wake-up once a wait is longer than its own
     * spin phase, which every picture-sized transfer is. */
    struct timespec t0, t, nap = { 0, g_wait_mode * 1000L };
    clock_gettime (CLOCK_MONOTONIC, &t0);
    for (;;) {
      const cudaError_t e = cudaEventQuery (cx->sync_ev);
      if (e == cudaSuccess) break;
      if (e != cudaErrorNotReady) SB2H_CUDA (e);
      clock_gettime (CLOCK_MONOTONIC, &t);
      if ((t.tv_sec - t0.tv_sec) * 1000000000L + (t.tv_nsec - t0.tv_nsec) > 40000L) nanosleep (&nap, NULL);
    }
  }
  cx->dirty = 0;
}

/* ---- last writer of a device region ------------------------------------------ */
#define SB2H_NWRITE 8192
static struct {
  const void *key;
  cudaEvent_t ev;
  cudaStream_t stream;                 /* last writer */
  unsigned long long users[SB2H_MAX_CTX / 64];   /* contexts that enqueued work on the region since it was allocated */
} g_writes[SB2H_NWRITE];
static pthread_mutex_t g_write_mutex = PTHREAD_MUTEX_INITIALIZER;

static int
write_slot (const void *key, int create)
{
  size_t h = (size_t) (((uintptr_t) key >> 8) * 2654435761u) % SB2H_NWRITE;
  int i;
  for (i = 0; i < SB2H_NWRITE; i++) {
    const int k = (int) ((h + i) % SB2H_NWRITE);
    if (g_writes[k].key == key) return k;
    if (!g_writes[k].key) {
      if (!create) return -1;
      g_writes[k].key = key;                 /* keys are never removed: a region that is freed and */
      return k;                              /* handed out again keeps its (by then complete) event */
    }
  }
  return -1;
}

void
sb2h_frame_use (Sb2hContext *cx, const void *region)
{
  int k;
  pthread_mutex_lock (&g_write_mutex);
  k = write_slot (region, 1);
  if (k >= 0) {
    if (g_writes[k].stream && g_writes[k].stream != cx->stream)
      SB2H_CUDA (cudaStreamWaitEvent (cx->stream, g_writes[k].ev, 0));
    g_writes[k].users[cx->slot / 64] |= 1ull << (cx->slot % 64);
  }
  pthread_mutex_unlock (&g_write_mutex);
  if (k < 0) sb2h_sync (cx);     /* table full: at least order this thread */
}

/* which contexts may still have work in flight on `region`; forgets them (the region is being freed) */
static int
frame_take_users (const void *region, unsigned long long *users)
{
  int k, i, known;
  pthread_mutex_lock (&g_write_mutex);
  k = write_slot (region, 0);
  known = k >= 0;
  for (i = 0; i < SB2H_MAX_CTX / 64; i++) {
    users[i] = known ? g_writes[k].users[i] : 0;
    if (known) g_writes[k].users[i] = 0;
  }
  pthread_mutex_unlock (&g_write_mutex);
  return known;
}

/* ---- per-thread spare region for the in-place transforms ------------------------------------
 * An in-place wavelet transform of a CUDA-domain frame writes into a second region and swaps (host
 * wavelet layer).  Going through the domain for that -- free the old region, allocate the next call's new one
 * -- makes the workers wait on each other: a freed region with work in flight is parked, and an allocation
 * of that size waits (on the host, holding the domain's lock) for the parked blocks of EVERY worker.  So a
 * thread keeps the region it has just transformed OUT OF as its spare for the next call, provided no other
 * thread ever touched that region (its reuse is then ordered by the thread's own stream). */
void *
sb2h_spare_take (Sb2hContext *cx, SchroMemoryDomain *domain, int size)
{
  void *p = NULL;
  pthread_mutex_lock (&g_spare_mutex);
  if (g_spare[cx->slot].ptr && g_spare[cx->slot].domain == domain && g_spare[cx->slot].size == size) {
    p = g_spare[cx->slot].ptr;
    g_spare[cx->slot].ptr = NULL;
  }
  pthread_mutex_unlock (&g_spare_mutex);
  return p;
}

int
sb2h_spare_put (Sb2hContext *cx, SchroMemoryDomain *domain, void *region, int size)
{
  unsigned long long users[SB2H_MAX_CTX / 64];
  cudaStream_t last = NULL;
  int k, i, c, only_me = 1, kept = 0;
  pthread_mutex_lock (&g_write_mutex);
  k = write_slot (region, 0);
  for (i = 0; i < SB2H_MAX_CTX / 64; i++) users[i] = k >= 0 ? g_writes[k].users[i] : 0;
  if (k >= 0) last = g_writes[k].stream;
  pthread_mutex_unlock (&g_write_mutex);
  /* other threads count only while they are alive and have un-waited work (a thread that has waited, or has
   * gone -- it synchronises its stream on the way out -- cannot have anything in flight on the region) */
  pthread_mutex_lock (&g_device_mutex);
  for (c = 0; c < SB2H_MAX_CTX; c++) {
    if (c == cx->slot || !g_ctx[c] || !g_ctx[c]->dirty) continue;
    if (((users[c / 64] >> (c % 64)) & 1) || (last && g_ctx[c]->stream == last)) only_me = 0;
  }
  pthread_mutex_unlock (&g_device_mutex);
  if (!only_me) return 0;
  pthread_mutex_lock (&g_write_mutex);
  if (k >= 0) {
    /* a new life for the region: only this thread's stream matters from here on */
    for (i = 0; i < SB2H_MAX_CTX / 64; i++) g_writes[k].users[i] = 0;
    g_writes[k].users[cx->slot / 64] = 1ull << (cx->slot % 64);
    if (g_writes[k].stream != cx->stream) g_writes[k].stream = NULL;
  }
  pthread_mutex_unlock (&g_write_mutex);
  pthread_mutex_lock (&g_spare_mutex);
  if (!g_spare[cx->slot].ptr) {
    g_spare[cx->slot].domain = domain;
    g_spare[cx->slot].ptr = region;
    g_spare[cx->slot].size = size;
    kept = 1;
  }
  pthread_mutex_unlock (&g_spare_mutex);
  return kept;
}

/* a domain is going away (its regions are released with it), or a thread is: forget / hand back the spares */
static void
spare_forget_domain (SchroMemoryDomain *domain)
{
  int c;
  pthread_mutex_lock (&g_spare_mutex);
  for (c = 0; c < SB2H_MAX_CTX; c++)
    if (g_spare[c].domain == domain) g_spare[c].ptr = NULL;
  pthread_mutex_unlock (&g_spare_mutex);
}

void
sb2h_frame_wrote (Sb2hContext *cx, const void *region)
{
  int k;
  pthread_mutex_lock (&g_write_mutex);
  k = write_slot (region, 1);
  if (k < 0) {
    /* table full: fall back to waiting */
    pthread_mutex_unlock (&g_write_mutex);
    sb2h_sync (cx);
    return;
  }
  if (!g_writes[k].ev) SB2H_CUDA (cudaEventCreateWithFlags (&g_writes[k].ev, cudaEventDisableTiming));
  SB2H_CUDA (cudaEventRecord (g_writes[k].ev, cx->stream));
  g_writes[k].stream = cx->stream;
  g_writes[k].users[cx->slot / 64] |= 1ull << (cx->slot % 64);
  pthread_mutex_unlock (&g_write_mutex);
  cx->dirty = 1;
}

void *
sb2h_dev_buffer (Sb2hContext *cx, int which, size_t bytes)
{
  if (bytes == 0) return NULL;
  if (cx->dev_size[which] < bytes) {
    if (cx->dev[which]) {
      sb2h_sync (cx);
      SB2H_CUDA (cudaFree (cx->dev[which]));
    }
    bytes = (bytes + 0xfffff) & ~(size_t) 0xfffff;
    SB2H_CUDA (cudaMalloc (&cx->dev[which], bytes));
    cx->dev_size[which] = bytes;
  }
  return cx->dev[which];
}

void
sb2h_upload_staged (Sb2hContext *cx, void *dev_dst, const void *host_src, size_t bytes)
{
  if (bytes == 0) return;
  if (sb2h_mem_kind (host_src) != SB2H_MEM_PAGEABLE) {
    SB2H_CUDA (cudaMemcpyAsync (dev_dst, host_src, bytes, cudaMemcpyDefault, cx->stream));
    return;
  }
  if (cx->pin_busy) {                      /* the previous DMA out of the block (long finished in practice) */
    SB2H_CUDA (cudaEventSynchronize (cx->pin_ev));
    cx->pin_busy = 0;
  }
  if (cx->pin_size < bytes) {
    if (cx->pin) SB2H_CUDA (cudaFreeHost (cx->pin));
    cx->pin_size = (bytes + 0xfffff) & ~(size_t) 0xfffff;
    SB2H_CUDA (cudaHostAlloc (&cx->pin, cx->pin_size, cudaHostAllocDefault));
    if (!cx->pin_ev) SB2H_CUDA (cudaEventCreateWithFlags (&cx->pin_ev, cudaEventDisableTiming));
  }
  memcpy (cx->pin, host_src, bytes);
  SB2H_CUDA (cudaMemcpyAsync (dev_dst, cx->pin, bytes, cudaMemcpyHostToDevice, cx->stream));
  SB2H_CUDA (cudaEventRecord (cx->pin_ev, cx->stream));
  cx->pin_busy = 1;
}

/* Process-wide pool of device blocks, reused by exact size (no cudaMalloc / cudaFree -- and so no
 * device-wide synchronisation -- in steady state).  A block may be handed back by another thread
 * than the one that allocated it (SchroAsync moves the stages of a picture between workers), and
 * while work that uses it is still in flight: the free records an event on the freeing thread's
 * stream, and the next owner's streams wait for that event before they touch the block. */
#define SB2H_POOL_SLOTS 2048
static struct { void *ptr; size_t bytes; int in_use; cudaEvent_t ev; int ev_valid; } g_pool[SB2H_POOL_SLOTS];
static pthread_mutex_t g_pool_mutex = PTHREAD_MUTEX_INITIALIZER;

void *
sb2h_pool_alloc (size_t bytes)
{
  Sb2hContext *cx = sb2h_context ();
  int i, free_slot = -1, idle_slot = -1;
  void *p, *old = NULL;
  cudaEvent_t old_ev = NULL;
  /* No CUDA call is made with the pool's lock held: a call that has to wait (cudaMalloc behind a busy device, an
   * enqueue behind a full channel) would otherwise stall every other worker's alloc and free with it.  A slot that
   * is marked in use belongs to one thread, so its event and pointer can be touched outside the lock. */
  pthread_mutex_lock (&g_pool_mutex);
  for (i = 0; i < SB2H_POOL_SLOTS; i++) {
    if (g_pool[i].ptr && !g_pool[i].in_use) {
      if (g_pool[i].bytes == bytes) {
        const int wait = g_pool[i].ev_valid;
        cudaEvent_t ev = g_pool[i].ev;
        g_pool[i].in_use = 1;
        p = g_pool[i].ptr;
        pthread_mutex_unlock (&g_pool_mutex);
        if (wait) {
          /* (the event is only re-recorded by the next free of this slot, which is this caller's) */
          SB2H_CUDA (cudaStreamWaitEvent (cx->stream, ev, 0));
          if (cx->stream_hi != cx->stream) SB2H_CUDA (cudaStreamWaitEvent (cx->stream_hi, ev, 0));
        }
        return p;
      }
      if (idle_slot < 0) idle_slot = i;
    }
    if (!g_pool[i].ptr && free_slot < 0) free_slot = i;
  }
  if (free_slot < 0) {
    /* table full: recycle an idle block of another size */
    if (idle_slot < 0) sb2h_fatal (__func__, "device block pool exhausted (%d blocks in use)", SB2H_POOL_SLOTS);
    old = g_pool[idle_slot].ptr;
    if (g_pool[idle_slot].ev_valid) old_ev = g_pool[idle_slot].ev;
    free_slot = idle_slot;
  }
  /* reserve the slot: in use, with a pointer no caller can hold, until the block exists */
  g_pool[free_slot].ptr = (void *) &g_pool[free_slot];
  g_pool[free_slot].bytes = 0;
  g_pool[free_slot].in_use = 1;
  g_pool[free_slot].ev_valid = 0;
  pthread_mutex_unlock (&g_pool_mutex);
  if (old) {
    if (old_ev) SB2H_CUDA (cudaEventSynchronize (old_ev));
    SB2H_CUDA (cudaFree (old));
  }
  SB2H_CUDA (cudaMalloc (&p, bytes + 256));
  pthread_mutex_lock (&g_pool_mutex);
  g_pool[free_slot].ptr = p;
  g_pool[free_slot].bytes = bytes;
  pthread_mutex_unlock (&g_pool_mutex);
  return p;
}

void
sb2h_pool_free (void *ptr)
{
  Sb2hContext *cx;
  cudaEvent_t ev = NULL;
  int i, slot = -1;
  if (!ptr) return;
  cx = sb2h_context ();
  pthread_mutex_lock (&g_pool_mutex);
  for (i = 0; i < SB2H_POOL_SLOTS; i++)
    if (g_pool[i].ptr == ptr && g_pool[i].in_use) { slot = i; ev = g_pool[i].ev; break; }
  pthread_mutex_unlock (&g_pool_mutex);
  if (slot < 0) sb2h_fatal (__func__, "%p is not a pooled device block in use", ptr);
  if (!ev) SB2H_CUDA (cudaEventCreateWithFlags (&ev, cudaEventDisableTiming));
  SB2H_CUDA (cudaEventRecord (ev, cx->stream));
  pthread_mutex_lock (&g_pool_mutex);
  g_pool[slot].ev = ev;
  g_pool[slot].ev_valid = 1;
  g_pool[slot].in_use = 0;
  pthread_mutex_unlock (&g_pool_mutex);
}

/* idle blocks go back to the driver when the last thread context is released */
static void
sb2h_pool_release_all (void)
{
  int i;
  pthread_mutex_lock (&g_pool_mutex);
  for (i = 0; i < SB2H_POOL_SLOTS; i++)
    if (g_pool[i].ptr && !g_pool[i].in_use) {
      cudaFree (g_pool[i].ptr);
      g_pool[i].ptr = NULL;
      if (g_pool[i].ev) { cudaEventDestroy (g_pool[i].ev); g_pool[i].ev = NULL; }
      g_pool[i].ev_valid = 0;
    }
  pthread_mutex_unlock (&g_pool_mutex);
}

/* process-wide pool of page-locked host blocks (motion fields coming back from the GPU) */
#define SB2H_PINNED_SLOTS 256
static struct { void *ptr; size_t bytes; int in_use; } g_pinned[SB2H_PINNED_SLOTS];
static pthread_mutex_t g_pinned_mutex = PTHREAD_MUTEX_INITIALIZER;

void *
sb2h_pinned_pool_alloc (size_t bytes)
{
  int i, free_slot = -1;
  void *p = NULL, *old = NULL;
  pthread_mutex_lock (&g_pinned_mutex);
  for (i = 0; i < SB2H_PINNED_SLOTS; i++) {
    if (g_pinned[i].ptr && !g_pinned[i].in_use && g_pinned[i].bytes == bytes) {
      g_pinned[i].in_use = 1;
      p = g_pinned[i].ptr;
      break;
    }
    if (!g_pinned[i].ptr && free_slot < 0) free_slot = i;
  }
  if (p) {
    pthread_mutex_unlock (&g_pinned_mutex);
    return p;
  }
  if (free_slot < 0) {
    for (i = 0; i < SB2H_PINNED_SLOTS; i++)
      if (!g_pinned[i].in_use) {
        old = g_pinned[i].ptr;
        free_slot = i;
        break;
      }
    if (free_slot < 0) sb2h_fatal (__func__, "pinned block pool exhausted");
  }
  /* the slot is reserved (in use, pointer nobody holds) while the driver call runs outside the lock */
  g_pinned[free_slot].ptr = (void *) &g_pinned[free_slot];
  g_pinned[free_slot].bytes = 0;
  g_pinned[free_slot].in_use = 1;
  pthread_mutex_unlock (&g_pinned_mutex);
  if (old) SB2H_CUDA (cudaFreeHost (old));
  SB2H_CUDA (cudaHostAlloc (&p, bytes, cudaHostAllocPortable));
  pthread_mutex_lock (&g_pinned_mutex);
  g_pinned[free_slot].ptr = p;
  g_pinned[free_slot].bytes = bytes;
  pthread_mutex_unlock (&g_pinned_mutex);
  return p;
}

int
sb2h_pinned_pool_free (void *ptr)
{
  int i, found = 0;
  if (!ptr) return 0;
  pthread_mutex_lock (&g_pinned_mutex);
  for (i = 0; i < SB2H_PINNED_SLOTS; i++)
    if (g_pinned[i].ptr == ptr) {
      g_pinned[i].in_use = 0;
      found = 1;
      break;
    }
  pthread_mutex_unlock (&g_pinned_mutex);
  return found;
}

void
sb2h_copy_rect (Sb2hContext *cx, void *dst, size_t dst_stride, const void *src,
    size_t src_stride, size_t row_bytes, int rows)
{
  if (rows <= 0 || row_bytes == 0) return;
  SB2H_CUDA (cudaMemcpy2DAsync (dst, dst_stride, src, src_stride, row_bytes, rows,
          cudaMemcpyDefault, cx->stream));
}

/* ---- memory domains -------------------------------------------------------- */
/* Blocks of a CUDA domain that were handed back while a thread that used them still had
 * un-waited work on its stream: the slot stays marked in use until an event recorded on each
 * such stream at the time of the free has completed, so nothing enqueued before the free can
 * still touch the block when it is handed out again.  Reaped on every later alloc / free of
 * the domain. */
#define SB2H_LIMBO 256
typedef struct {
  SchroMemoryDomain *domain;
  int slot, nev;
  unsigned long long seq;          /* order of parking: the oldest parked block is the likeliest to be free */
  cudaEvent_t ev[SB2H_MAX_CTX];
} Sb2hLimbo;
static Sb2hLimbo *g_limbo[SB2H_LIMBO];
static unsigned long long g_limbo_seq;
static cudaEvent_t g_evpool[SB2H_LIMBO * 4];
static int g_nevpool;
static pthread_mutex_t g_limbo_mutex = PTHREAD_MUTEX_INITIALIZER;

static cudaEvent_t
evpool_get (void)
{
  cudaEvent_t e;
  if (g_nevpool > 0) return g_evpool[--g_nevpool];
  SB2H_CUDA (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
  return e;
}

static void
evpool_put (cudaEvent_t e)
{
  if (g_nevpool < (int) (sizeof (g_evpool) / sizeof (g_evpool[0]))) g_evpool[g_nevpool++] = e;
  else cudaEventDestroy (e);
}

/* domain->mutex held.  wait != 0: block until every parked block of the domain is free. */
static void
limbo_reap (SchroMemoryDomain *domain, int wait)
{
  int i, k;
  pthread_mutex_lock (&g_limbo_mutex);
  for (i = 0; i < SB2H_LIMBO; i++) {
    Sb2hLimbo *l = g_limbo[i];
    if (!l || l->domain != domain) continue;
    for (k = 0; k < l->nev; ) {
      cudaError_t e = wait ? cudaEventSynchronize (l->ev[k]) : cudaEventQuery (l->ev[k]);
      if (e == cudaSuccess) {
        evpool_put (l->ev[k]);
        l->ev[k] = l->ev[--l->nev];
      } else if (e == cudaErrorNotReady) {
        k++;
      } else {
        sb2h_fatal (__func__, "event query: %s", cudaGetErrorString (e));
      }
    }
    if (l->nev == 0) {
      domain->slots[l->slot].flags &= ~SCHRO_MEMORY_DOMAIN_SLOT_IN_USE;
      free (l);
      g_limbo[i] = NULL;
    }
  }
  pthread_mutex_unlock (&g_limbo_mutex);
}


/* Every region a CUDA domain owns, by address range: entry points that only see a plane pointer
 * (SchroFrameData carries no frame, schroframe.h:58-68) find the region -- and with it the
 * region's last-writer record -- from the pointer. */
#define SB2H_NREGION 4096
static struct { char *ptr; size_t size; } g_regions[SB2H_NREGION];
static pthread_mutex_t g_region_mutex = PTHREAD_MUTEX_INITIALIZER;

const void *
sb2h_region_of (const void *ptr)
{
  const char *p = ptr, *base = NULL;
  int i;
  pthread_mutex_lock (&g_region_mutex);
  for (i = 0; i < SB2H_NREGION; i++)
    if (g_regions[i].ptr && p >= g_regions[i].ptr && p < g_regions[i].ptr + g_regions[i].size) {
      base = g_regions[i].ptr;
      break;
    }
  pthread_mutex_unlock (&g_region_mutex);
  return base;
}

/* before a stream-ordered access through a bare device pointer: order it behind the last write of
 * the containing region; memory this library does not know is ordered by a device-wide wait */
void
sb2h_ptr_use (Sb2hContext *cx, const void *ptr)
{
  const void *region = sb2h_region_of (ptr);
  if (region) sb2h_frame_use (cx, region);
  else SB2H_CUDA (cudaDeviceSynchronize ());
}

void
sb2h_ptr_wrote (Sb2hContext *cx, const void *ptr)
{
  const void *region = sb2h_region_of (ptr);
  if (region) sb2h_frame_wrote (cx, region);
}

static void *
cuda_alloc (int size)
{
  void *p = NULL;
  int i;
  /* 256 spare bytes: the byte-SIMD SAD kernels read whole aligned words */
  SB2H_CUDA (cudaMalloc (&p, (size_t) size + 256));
  pthread_mutex_lock (&g_region_mutex);
  for (i = 0; i < SB2H_NREGION && g_regions[i].ptr; i++) ;
  if (i < SB2H_NREGION) {
    g_regions[i].ptr = p;
    g_regions[i].size = (size_t) size + 256;
  }
  pthread_mutex_unlock (&g_region_mutex);
  return p;
}

static void
cuda_free (void *ptr, int size)
{
  int i;
  (void) size;
  pthread_mutex_lock (&g_region_mutex);
  for (i = 0; i < SB2H_NREGION; i++)
    if (g_regions[i].ptr == (char *) ptr) g_regions[i].ptr = NULL;
  pthread_mutex_unlock (&g_region_mutex);
  SB2H_CUDA (cudaFree (ptr));
}

static void *
pinned_alloc (int size)
{
  void *p = NULL;
  SB2H_CUDA (cudaHostAlloc (&p, (size_t) size, cudaHostAllocPortable));
  return p;
}

static void
pinned_free (void *ptr, int size)
{
  (void) size;
  SB2H_CUDA (cudaFreeHost (ptr));
}

static SchroMemoryDomain *
domain_new (unsigned int flags, void *(*alloc) (int), void (*free_fn) (void *, int))
{
  SchroMemoryDomain *d = calloc (1, sizeof (SchroMemoryDomain));
  pthread_mutex_t *mu = malloc (sizeof (pthread_mutex_t));
  pthread_mutex_init (mu, NULL);
  d->mutex = mu;
  d->flags = flags;
  d->alloc = alloc;
  d->free = free_fn;
  return d;
}

SchroMemoryDomain *
schro_memory_domain_new_cuda (void)
{
  return domain_new (SCHRO_MEMORY_DOMAIN_CUDA, cuda_alloc, cuda_free);
}

SchroMemoryDomain *
schro_memory_domain_new_pinned (void)
{
  return domain_new (SCHRO_MEMORY_DOMAIN_CPU | SCHRO_MEMORY_DOMAIN_PINNED, pinned_alloc,
      pinned_free);
}

void
schro_memory_domain_free (SchroMemoryDomain *domain)
{
  int i;
  SB2H_ASSERT (domain != NULL);
  spare_forget_domain (domain);
  if (domain->flags & SCHRO_MEMORY_DOMAIN_CUDA) {
    pthread_mutex_lock (domain->mutex);
    limbo_reap (domain, 1);
    pthread_mutex_unlock (domain->mutex);
  }
  for (i = 0; i < SCHRO_MEMORY_DOMAIN_SLOTS; i++) {
    if (domain->slots[i].flags & SCHRO_MEMORY_DOMAIN_SLOT_ALLOCATED)
      domain->free (domain->slots[i].ptr, domain->slots[i].size);
  }
  pthread_mutex_destroy (domain->mutex);
  free (domain->mutex);
  free (domain);
}

/* exact-size slot reuse, as the reference's buffer pool does */
void *
schro_memory_domain_alloc (SchroMemoryDomain *domain, int size)
{
  int i;
  void *ptr = NULL;
  SB2H_ASSERT (domain != NULL);
  pthread_mutex_lock (domain->mutex);
  if (domain->flags & SCHRO_MEMORY_DOMAIN_CUDA) limbo_reap (domain, 0);
  for (i = 0; i < SCHRO_MEMORY_DOMAIN_SLOTS; i++) {
    unsigned int f = domain->slots[i].flags;
    if ((f & SCHRO_MEMORY_DOMAIN_SLOT_ALLOCATED) && !(f & SCHRO_MEMORY_DOMAIN_SLOT_IN_USE) &&
        domain->slots[i].size == size) {
      domain->slots[i].flags |= SCHRO_MEMORY_DOMAIN_SLOT_IN_USE;
      ptr = domain->slots[i].ptr;
      break;
    }
  }
  if (!ptr && (domain->flags & SCHRO_MEMORY_DOMAIN_CUDA)) {
    /* a parked block of this size beats a new cudaMalloc (which would stall the whole device,
     * and a caller that never waits would otherwise grow the domain without bound) */
    Sb2hLimbo *l = NULL;
    int at = -1;
    pthread_mutex_lock (&g_limbo_mutex);
    for (i = 0; i < SB2H_LIMBO; i++)
      if (g_limbo[i] && g_limbo[i]->domain == domain && domain->slots[g_limbo[i]->slot].size == size &&
          (at < 0 || g_limbo[i]->seq < g_limbo[at]->seq)) at = i;
    if (at >= 0) { l = g_limbo[at]; g_limbo[at] = NULL; }
    pthread_mutex_unlock (&g_limbo_mutex);
    if (l) {
      /* the block is this caller's now (its slot stays marked in use throughout); the wait for the work that
       * was in flight when it was handed back happens with no lock held, so other threads' allocs and frees
       * of the domain go on meanwhile */
      int k;
      pthread_mutex_unlock (domain->mutex);
      for (k = 0; k < l->nev; k++) SB2H_CUDA (cudaEventSynchronize (l->ev[k]));
      pthread_mutex_lock (&g_limbo_mutex);
      for (k = 0; k < l->nev; k++) evpool_put (l->ev[k]);
      pthread_mutex_unlock (&g_limbo_mutex);
      ptr = domain->slots[l->slot].ptr;
      free (l);
      return ptr;
    }
  }
  if (!ptr) {
    for (i = 0; i < SCHRO_MEMORY_DOMAIN_SLOTS; i++) {
      if (!(domain->slots[i].flags & SCHRO_MEMORY_DOMAIN_SLOT_ALLOCATED)) {
        domain->slots[i].flags = SCHRO_MEMORY_DOMAIN_SLOT_ALLOCATED | SCHRO_MEMORY_DOMAIN_SLOT_IN_USE;
        domain->slots[i].size = size;
        domain->slots[i].ptr = domain->alloc (size);
        ptr = domain->slots[i].ptr;
        break;
      }
    }
  }
  pthread_mutex_unlock (domain->mutex);
  if (!ptr) sb2h_fatal (__func__, "memory domain out of slots");
  return ptr;
}

void
schro_memory_domain_memfree (SchroMemoryDomain *domain, void *ptr)
{
  int i;
  SB2H_ASSERT (domain != NULL);
  pthread_mutex_lock (domain->mutex);
  for (i = 0; i < SCHRO_MEMORY_DOMAIN_SLOTS; i++) {
    if ((domain->slots[i].flags & SCHRO_MEMORY_DOMAIN_SLOT_IN_USE) && domain->slots[i].ptr == ptr) {
      Sb2hLimbo *l = NULL;
      if (domain->flags & SCHRO_MEMORY_DOMAIN_CUDA) {
        unsigned long long users[SB2H_MAX_CTX / 64];
        int c, k;
        limbo_reap (domain, 0);
        l = calloc (1, sizeof (Sb2hLimbo));
        l->domain = domain;
        l->slot = i;
        frame_take_users (ptr, users);
        pthread_mutex_lock (&g_device_mutex);
        for (c = 0; c < SB2H_MAX_CTX; c++)
          if (g_ctx[c] && g_ctx[c]->dirty && ((users[c / 64] >> (c % 64)) & 1)) {
            pthread_mutex_lock (&g_limbo_mutex);
            l->ev[l->nev] = evpool_get ();
            pthread_mutex_unlock (&g_limbo_mutex);
            SB2H_CUDA (cudaEventRecord (l->ev[l->nev], g_ctx[c]->stream));
            l->nev++;
          }
        pthread_mutex_unlock (&g_device_mutex);
        if (l->nev) {
          pthread_mutex_lock (&g_limbo_mutex);
          for (k = 0; k < SB2H_LIMBO && g_limbo[k]; k++) ;
          if (k < SB2H_LIMBO) { l->seq = g_limbo_seq++; g_limbo[k] = l; }
          pthread_mutex_unlock (&g_limbo_mutex);
          if (k == SB2H_LIMBO) {
            /* no room to park it: wait here */
            for (c = 0; c < l->nev; c++) {
              SB2H_CUDA (cudaEventSynchronize (l->ev[c]));
              pthread_mutex_lock (&g_limbo_mutex);
              evpool_put (l->ev[c]);
              pthread_mutex_unlock (&g_limbo_mutex);
            }
            l->nev = 0;
          }
        }
        if (l->nev == 0) { free (l); l = NULL; }
      }
      if (!l) domain->slots[i].flags &= ~SCHRO_MEMORY_DOMAIN_SLOT_IN_USE;
      pthread_mutex_unlock (domain->mutex);
      return;
    }
  }
  pthread_mutex_unlock (domain->mutex);
  sb2h_fatal (__func__, "pointer %p does not belong to this domain", ptr);
}

/* ---- frames ------------------------------------------------------------------ */
static pthread_mutex_t frame_mutex = PTHREAD_MUTEX_INITIALIZER;

SchroFrame *
schro_frame_new (void)
{
  SchroFrame *frame = calloc (1, sizeof (SchroFrame));
  frame->refcount = 1;
  return frame;
}

#define RUP16(x) (((x) + 15) & ~15)

SchroFrame *
schro_frame_new_and_alloc_full (SchroMemoryDomain *domain, SchroFrameFormat format,
    int width, int height, int extension, int upsampled)
{
  SchroFrame *frame = schro_frame_new ();
  const int bpp = sb2h_bpp (format);
  const int hs = SCHRO_FRAME_FORMAT_H_SHIFT (format), vs = SCHRO_FRAME_FORMAT_V_SHIFT (format);
  size_t total = 0, pos = 0;
  int k;

  SB2H_ASSERT (width > 0 && height > 0);
  if (format & 0x100) sb2h_fatal (__func__, "packed formats are outside the picture core");
  frame->format = format;
  frame->width = width;
  frame->height = height;
  frame->domain = domain;
  frame->extension = extension;
  frame->is_upsampled = upsampled;
  for (k = 0; k < 3; k++) {
    SchroFrameData *c = &frame->components[k];
    c->format = format;
    c->width = k ? (width + (1 << hs) - 1) >> hs : width;
    c->height = k ? (height + (1 << vs) - 1) >> vs : height;
    c->stride = RUP16 ((c->width + extension * 2) * bpp);
    if (upsampled) c->stride *= 4;
    c->length = c->stride * (c->height + extension * 2);
    c->h_shift = k ? hs : 0;
    c->v_shift = k ? vs : 0;
    total += (size_t) c->length;
  }
  if (domain)
    frame->regions[0] = schro_memory_domain_alloc (domain, (int) total);
  else
    frame->regions[0] = malloc (total);
  for (k = 0; k < 3; k++) {
    SchroFrameData *c = &frame->components[k];
    c->data = (char *) frame->regions[0] + pos + (size_t) c->stride * extension + (size_t) bpp * extension;
    pos += (size_t) c->length;
  }
  return frame;
}

SchroFrame *
schro_frame_new_and_alloc_extended (SchroMemoryDomain *domain, SchroFrameFormat format,
    int width, int height, int extension)
{
  return schro_frame_new_and_alloc_full (domain, format, width, height, extension, 0);
}

SchroFrame *
schro_frame_new_and_alloc (SchroMemoryDomain *domain, SchroFrameFormat format, int width,
    int height)
{
  return schro_frame_new_and_alloc_full (domain, format, width, height, 0, 0);
}

SchroFrame *
schro_frame_ref (SchroFrame *frame)
{
  pthread_mutex_lock (&frame_mutex);
  frame->refcount++;
  pthread_mutex_unlock (&frame_mutex);
  return frame;
}

void
schro_frame_unref (SchroFrame *frame)
{
  int k, last;
  SB2H_ASSERT (frame && frame->refcount > 0);
  pthread_mutex_lock (&frame_mutex);
  last = (--frame->refcount == 0);
  pthread_mutex_unlock (&frame_mutex);
  if (!last) return;
  if (frame->free) frame->free (frame, frame->priv);
  for (k = 0; k < 3; k++) {
    if (frame->regions[k]) {
      if (frame->domain) schro_memory_domain_memfree (frame->domain, frame->regions[k]);
      else free (frame->regions[k]);
    }
  }
  free (frame);
}

void
schro_upsampled_frame_get_framedata (SchroFrame *upframe, SchroFrameData *fd, int up_index,
    int component)
{
  SB2H_ASSERT (upframe->is_upsampled);
  *fd = upframe->components[component];
  fd->data = (char *) fd->data + (fd->stride >> 2) * up_index;
}

/* copy every component (with its borders and, for upsampled frames, all four phases) */
static void
frame_copy_all (SchroFrame *dest, SchroFrame *src)
{
  Sb2hContext *cx = sb2h_context ();
  int k;
  SB2H_ASSERT (dest->format == src->format && dest->width == src->width &&
      dest->height == src->height);
  if (sb2h_mem_kind (src->regions[0]) == SB2H_MEM_DEVICE) sb2h_frame_use (cx, src->regions[0]);
  if (sb2h_mem_kind (dest->regions[0]) == SB2H_MEM_DEVICE) sb2h_frame_use (cx, dest->regions[0]);
  for (k = 0; k < 3; k++) {
    SchroFrameData *d = &dest->components[k], *s = &src->components[k];
    const int bpp = sb2h_bpp (src->format);
    if (dest->extension == src->extension && dest->is_upsampled == src->is_upsampled) {
      /* same geometry: move the whole plane including borders / phases */
      const int ext = src->extension;
      char *dp = (char *) d->data - (size_t) d->stride * ext - (size_t) bpp * ext;
      char *sp = (char *) s->data - (size_t) s->stride * ext - (size_t) bpp * ext;
      SB2H_ASSERT (d->length == s->length);
      SB2H_CUDA (cudaMemcpyAsync (dp, sp, (size_t) s->length, cudaMemcpyDefault, cx->stream));
    } else {
      sb2h_copy_rect (cx, d->data, d->stride, s->data, s->stride, (size_t) s->width * bpp,
          s->height);
    }
  }
  sb2h_sync (cx);       /* the host side of the copy may be reused / read as soon as we return */
  dest->upsample_done = (dest->is_upsampled == src->is_upsampled &&
      dest->extension == src->extension) ? src->upsample_done : 0;
}

void
schro_frame_to_gpu (SchroFrame *dest, SchroFrame *src)
{
  frame_copy_all (dest, src);
}

void
schro_gpuframe_to_cpu (SchroFrame *dest, SchroFrame *src)
{
  frame_copy_all (dest, src);
}
