/* schro_host.h -- internals shared by the C host layer (schroedinger_b200/host/). */
#ifndef SCHRO_HOST_H
#define SCHRO_HOST_H

#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stddef.h>
#include "schro_b200.h"
#include "schro_b200_compat.h"

/* log + abort, like SCHRO_ASSERT / SCHRO_ERROR (schroedinger/schrodebug.h:55-60) */
void sb2h_fatal (const char *func, const char *fmt, ...);
#define SB2H_CHECK(rc, what) do { if ((rc) != 0) sb2h_fatal (__func__, "%s failed (%d): %s", what, (int) (rc), sb2_last_error ()); } while (0)
#define SB2H_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) sb2h_fatal (__func__, "%s: %s", #call, cudaGetErrorString (e_)); } while (0)
#define SB2H_ASSERT(cond) do { if (!(cond)) sb2h_fatal (__func__, "assertion failed: %s", #cond); } while (0)

enum { SB2H_MEM_PAGEABLE = 0, SB2H_MEM_PINNED = 1, SB2H_MEM_DEVICE = 2 };
int sb2h_mem_kind (const void *ptr);

/* per-thread staging context: one stream, grow-only device and pinned buffers */
enum { SB2H_BUF_IN = 0, SB2H_BUF_OUT, SB2H_BUF_WS, SB2H_BUF_AUX0, SB2H_BUF_AUX1, SB2H_BUF_AUX2,
       SB2H_BUF_AUX3, SB2H_NBUF };
typedef struct {
  cudaStream_t stream;
  void *dev[SB2H_NBUF];
  size_t dev_size[SB2H_NBUF];
  cudaStream_t stream_hi;  /* highest-priority side stream for the latency-bound wavefront kernels */
  cudaEvent_t ev_fork, ev_join;
  cudaEvent_t sync_ev;     /* blocking-sync event: waiting threads sleep instead of spinning */
  void *pin;               /* page-locked staging block for small host arrays that arrive in pageable memory */
  size_t pin_size;
  cudaEvent_t pin_ev;      /* the last DMA out of `pin` */
  int pin_busy;
  volatile int dirty;      /* work was enqueued on `stream` that no call has waited for yet */
  int slot;                /* index in the process-wide context table */
} Sb2hContext;

Sb2hContext *sb2h_context (void);

/* Stream-ordered device frames.  Calls whose frames all live in device memory do not wait for
 * the GPU: they enqueue on the calling thread's stream and return.  Ordering between threads
 * is kept per frame region: sb2h_frame_wrote records the stream's position after a write,
 * sb2h_frame_use makes the calling thread's stream wait for the last write made from another
 * stream (read-after-write, write-after-write).  Memory handed back to a CUDA domain while
 * any stream still has un-waited work is parked until that work has finished (core.c, limbo).
 * Everything that makes results visible to the host (D2H copies, metrics, staged host frames)
 * ends with sb2h_sync. */
void sb2h_sync (Sb2hContext *cx);
void sb2h_frame_use (Sb2hContext *cx, const void *region);
void sb2h_frame_wrote (Sb2hContext *cx, const void *region);
/* the same through a pointer INTO a region (plane pointers of a SchroFrameData): the containing
 * CUDA-domain region is looked up by address; unknown device memory falls back to a device-wide wait */
const void *sb2h_region_of (const void *ptr);
void sb2h_ptr_use (Sb2hContext *cx, const void *ptr);
void sb2h_ptr_wrote (Sb2hContext *cx, const void *ptr);
/* the calling thread's spare region for in-place transforms of CUDA-domain frames (core.c): take returns a region of
 * exactly that domain and size or NULL; put keeps `region` (allocated from `domain`) as the spare and returns 1 when no
 * other thread ever touched it and the thread has no spare yet, else returns 0 and the caller frees it */
void *sb2h_spare_take (Sb2hContext *cx, SchroMemoryDomain *domain, int size);
int sb2h_spare_put (Sb2hContext *cx, SchroMemoryDomain *domain, void *region, int size);
/* process-wide pool of page-locked host blocks, reused by exact size */
void *sb2h_pinned_pool_alloc (size_t bytes);
int sb2h_pinned_pool_free (void *ptr);     /* 0 if ptr is not a pool block */
/* process-wide pool of device blocks, reused by exact size (no cudaMalloc / cudaFree -- and
 * therefore no device-wide synchronisation -- in steady state); a block may be freed by another
 * thread than its allocator and while stream-ordered work on it is in flight (core.c) */
void *sb2h_pool_alloc (size_t bytes);
void sb2h_pool_free (void *ptr);
void *sb2h_dev_buffer (Sb2hContext *cx, int which, size_t bytes);
/* Upload of a host array that may live in pageable memory without a pageable cudaMemcpyAsync (which waits for
 * everything queued on the stream before it and holds the driver's lock while it stages): the bytes are copied
 * by the CPU into the thread's page-locked staging block and DMA'd from there; returns at once */
void sb2h_upload_staged (Sb2hContext *cx, void *dev_dst, const void *host_src, size_t bytes);

/* a u8 pyramid level as a one-picture slab: zero-copy for CUDA-domain frames, else uploaded once into
 * *cache (a pool block the caller frees); returns 1 when it enqueued an upload out of page-locked memory */
int sb2h_level_slab (Sb2hContext *cx, SchroFrame *f, void **cache, sb2_slab *slab);

static inline int sb2h_bpp (SchroFrameFormat format)
{
  switch (SCHRO_FRAME_FORMAT_DEPTH (format)) {
    case SCHRO_FRAME_FORMAT_DEPTH_U8: return 1;
    case SCHRO_FRAME_FORMAT_DEPTH_S16: return 2;
    default: return 4;
  }
}

/* rows x row_bytes rectangle, host (any kind) or device on either side, on cx->stream.
 * Pinned host memory is DMA'd asynchronously; pageable memory is staged by the CUDA
 * runtime (slower, the documented slow path). */
void sb2h_copy_rect (Sb2hContext *cx, void *dst, size_t dst_stride, const void *src,
    size_t src_stride, size_t row_bytes, int rows);

#endif
