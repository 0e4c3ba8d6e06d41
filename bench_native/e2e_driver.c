/*
 * e2e_driver.c -- the host side of bench.py's end-to-end leg: a pool of pthreads drives the
 * drop-in schro_* C API (libschro_b200.so) over pinned host frames, one picture per task,
 * the way libschroedinger's own SchroAsync workers would (schroedinger/schroasync-pthread.c).
 * Python only builds the frames and reads the clock; no interpreter lock sits between two
 * API calls.  Benchmark infrastructure, not part of the product library.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "schro_b200_compat.h"

typedef struct {
  int nthreads, npictures, levels, pic_height, full_core;
  SchroParams *params;
  SchroFrame **coef_host;      /* [npictures] coefficient frames, page-locked */
  SchroFrame **src_host;       /* [npictures] source pictures, page-locked */
  SchroFrame **out_host;       /* [npictures] decoded pictures, page-locked */
  SchroFrame **ref_pyr;        /* [levels + 1] reference pyramid, CUDA domain */
  SchroFrame **coef_dev;       /* [nthreads] */
  SchroFrame **acc_dev;        /* [nthreads] */
  SchroFrame **out_dev;        /* [nthreads] upsampled, extension 32 */
  SchroMotion **motion;        /* [nthreads] */
  SchroFrame **src_pyr;        /* [nthreads * (levels + 1)] */
  /* widen != 0: the coefficients arrive QUANTISED as s16 (half the bytes of the s32 frame) and are
   * dequantised into the s32 coefficient frame on the device (schro_b200_frame_dequantise_widen) */
  int widen;
  SchroFrame **coef16_host;    /* [npictures] quantised coefficients, s16, page-locked */
  SchroFrame **coef16_dev;     /* [nthreads] */
  const int32_t *pairs;        /* (quant_factor, quant_offset + 2) per codeblock */
} Sb2E2eJob;

static Sb2E2eJob g_job;
static pthread_t *g_threads;
static pthread_barrier_t g_start, g_end;
static volatile int g_quit, g_passes = 1;
static volatile unsigned g_sink;

/* optional per-call wall-clock accounting (summed over threads), read with sb2_e2e_times */
enum { T_H2D_COEF, T_IWT, T_RENDER, T_EDGE_UPSAMPLE, T_D2H_OUT, T_H2D_SRC, T_PYRAMID, T_HBM_NEW, T_HBM_SCAN,
       T_HBM_HINT, T_HBM_UNREF, T_N };
static double g_times[64][T_N];
static int g_timing;
static inline double
now (void)
{
  struct timespec a;
  clock_gettime (CLOCK_MONOTONIC, &a);
  return (double) a.tv_sec + 1e-9 * (double) a.tv_nsec;
}
#define TIMED(slot, stmt) do { if (g_timing) { const double t0_ = now (); stmt; g_times[t & 63][slot] += now () - t0_; } else { stmt; } } while (0)

/* development aid (SB2_E2E_STAGES=mask): run only some stages of a picture, to see which of them limit the
 * threaded throughput.  1 = H2D coefficients, 2 = inverse wavelet + render + upsample, 4 = D2H picture,
 * 8 = H2D source, 16 = pyramid, 32 = block matching */
static int g_stage_mask = 0xff;

/* coefficients of picture i into thread t's device coefficient frame */
static void
upload_coefficients (const Sb2E2eJob *j, int t, int i)
{
  if (j->widen) {
    schro_frame_to_gpu (j->coef16_dev[t], j->coef16_host[i]);
    schro_b200_frame_dequantise_widen (j->coef_dev[t], j->coef16_dev[t], j->params, j->pairs);
  } else {
    schro_frame_to_gpu (j->coef_dev[t], j->coef_host[i]);
  }
}

static void
partial_picture (int t, int i)
{
  const Sb2E2eJob *j = &g_job;
  SchroFrame **pyr = j->src_pyr + (size_t) t * (j->levels + 1);
  int l;
  if (g_stage_mask & 1) upload_coefficients (j, t, i);
  if (g_stage_mask & 2) {
    schro_frame_inverse_iwt_transform (j->coef_dev[t], j->params);
    schro_motion_render (j->motion[t], j->acc_dev[t], j->coef_dev[t], 1, j->out_dev[t]);
    schro_frame_mc_edgeextend (j->out_dev[t]);
    j->out_dev[t]->upsample_done = 0;
    schro_upsampled_frame_upsample (j->out_dev[t]);
  }
  if (g_stage_mask & 4) schro_gpuframe_to_cpu (j->out_host[i], j->out_dev[t]);
  if (g_stage_mask & 8) schro_frame_to_gpu (pyr[0], j->src_host[i]);
  if (g_stage_mask & 16) {
    schro_frame_mc_edgeextend (pyr[0]);
    for (l = 0; l < j->levels; l++) {
      schro_frame_downsample (pyr[l + 1], pyr[l]);
      schro_frame_mc_edgeextend (pyr[l + 1]);
    }
  }
  if (g_stage_mask & 32) {
    SchroHierBm *hbm = schro_hbm_new_from_frames (j->params, 0, j->levels, 0, pyr, j->ref_pyr);
    schro_hbm_scan (hbm);
    schro_hierarchical_bm_scan_hint (hbm, 0, 3);
    g_sink += (unsigned) schro_hbm_motion_field (hbm, 0)->motion_vectors[0].metric;
    schro_hbm_unref (hbm);
  }
  if (!(g_stage_mask & (4 | 32))) schro_b200_thread_sync ();
}

static void
one_picture (int t, int i)
{
  const Sb2E2eJob *j = &g_job;
  SchroFrame **pyr = j->src_pyr + (size_t) t * (j->levels + 1);
  SchroHierBm *hbm;
  SchroMotionField *mf;
  int l;

  if (!j->full_core) {
    schro_frame_inverse_iwt_transform (j->coef_host[i], j->params);
    return;
  }
  if (g_stage_mask != 0xff) { partial_picture (t, i); return; }
  /* decode side: coefficients in, decoded picture out */
  TIMED (T_H2D_COEF, upload_coefficients (j, t, i));
  TIMED (T_IWT, schro_frame_inverse_iwt_transform (j->coef_dev[t], j->params));
  /* dest (picture size) gives the rendered area; the iwt-padded coefficient frame is the addframe,
   * as in schrodecoder.c:1784 */
  TIMED (T_RENDER, schro_motion_render (j->motion[t], j->acc_dev[t], j->coef_dev[t], 1, j->out_dev[t]));
  TIMED (T_EDGE_UPSAMPLE, {
    schro_frame_mc_edgeextend (j->out_dev[t]);
    j->out_dev[t]->upsample_done = 0;
    schro_upsampled_frame_upsample (j->out_dev[t]);
  });
  TIMED (T_D2H_OUT, schro_gpuframe_to_cpu (j->out_host[i], j->out_dev[t]));
  /* encode side: source picture in, motion fields out */
  TIMED (T_H2D_SRC, schro_frame_to_gpu (pyr[0], j->src_host[i]));
  TIMED (T_PYRAMID, {
    schro_frame_mc_edgeextend (pyr[0]);
    for (l = 0; l < j->levels; l++) {
      schro_frame_downsample (pyr[l + 1], pyr[l]);
      schro_frame_mc_edgeextend (pyr[l + 1]);
    }
  });
  TIMED (T_HBM_NEW, hbm = schro_hbm_new_from_frames (j->params, 0, j->levels, 0, pyr, j->ref_pyr));
  TIMED (T_HBM_SCAN, schro_hbm_scan (hbm));
  TIMED (T_HBM_HINT, schro_hierarchical_bm_scan_hint (hbm, 0, 3));
  mf = schro_hbm_motion_field (hbm, 0);
  g_sink += (unsigned) mf->motion_vectors[0].metric;      /* the host reads the result */
  TIMED (T_HBM_UNREF, schro_hbm_unref (hbm));
}

static void *
worker (void *arg)
{
  const int t = (int) (size_t) arg;
  for (;;) {
    int i, s;
    pthread_barrier_wait (&g_start);
    if (g_quit) break;
    /* passes follow each other without a barrier: a thread that is done with its pictures of one
     * step starts on the next step's, as the workers of a running codec would */
    for (s = 0; s < g_passes; s++)
      for (i = t; i < g_job.npictures; i += g_job.nthreads) one_picture (t, i);
    pthread_barrier_wait (&g_end);
  }
  schro_b200_thread_release ();
  return NULL;
}

int
sb2_e2e_start (const Sb2E2eJob *job)
{
  int t;
  g_job = *job;
  g_quit = 0;
  if (getenv ("SB2_E2E_STAGES")) g_stage_mask = atoi (getenv ("SB2_E2E_STAGES"));
  pthread_barrier_init (&g_start, NULL, (unsigned) job->nthreads + 1);
  pthread_barrier_init (&g_end, NULL, (unsigned) job->nthreads + 1);
  g_threads = calloc ((size_t) job->nthreads, sizeof (pthread_t));
  for (t = 0; t < job->nthreads; t++)
    if (pthread_create (&g_threads[t], NULL, worker, (void *) (size_t) t)) return -1;
  return 0;
}

/* `steps` steps, a step = every picture of the job once; returns wall seconds */
double
sb2_e2e_run (int steps)
{
  struct timespec a, b;
  g_passes = steps;
  clock_gettime (CLOCK_MONOTONIC, &a);
  pthread_barrier_wait (&g_start);
  pthread_barrier_wait (&g_end);
  clock_gettime (CLOCK_MONOTONIC, &b);
  return (double) (b.tv_sec - a.tv_sec) + 1e-9 * (double) (b.tv_nsec - a.tv_nsec);
}

double
sb2_e2e_step (void)
{
  return sb2_e2e_run (1);
}

/* enable / read the per-call accounting: out[T_N] seconds summed over threads since enabling */
int
sb2_e2e_times (int enable, double *out)
{
  int t, k;
  if (out)
    for (k = 0; k < T_N; k++) {
      out[k] = 0;
      for (t = 0; t < 64; t++) out[k] += g_times[t][k];
    }
  if (enable >= 0) {
    memset (g_times, 0, sizeof (g_times));
    g_timing = enable;
  }
  return T_N;
}

void
sb2_e2e_stop (void)
{
  int t;
  if (!g_threads) return;
  g_quit = 1;
  pthread_barrier_wait (&g_start);
  for (t = 0; t < g_job.nthreads; t++) pthread_join (g_threads[t], NULL);
  free (g_threads);
  g_threads = NULL;
  pthread_barrier_destroy (&g_start);
  pthread_barrier_destroy (&g_end);
}

/* ---- low-delay intra decode (BASELINE configs[1]): compressed slices in, 8-bit pictures out ----
 * A self-contained run: `nthreads` pthreads share the pictures round-robin, each picture is
 *   schro_b200_decode_lowdelay_transform_data -> schro_b200_frame_inverse_iwt_combine (inverse transform with the
 *   conversion to 8 bits fused into its last level) -> schro_gpuframe_to_cpu
 * on the thread's own CUDA-domain frames.  Returns the wall-clock seconds of `steps` passes over the pictures. */
typedef struct {
  int nthreads, npictures, slice_bytes;
  SchroParams *params;
  const uint8_t **slices;      /* [npictures] page-locked slice buffers */
  SchroFrame **out_host;       /* [npictures] page-locked u8 pictures */
  SchroFrame **coef_dev;       /* [nthreads] CUDA-domain coefficient frames (iwt size) */
  SchroFrame **u8_dev;         /* [nthreads] CUDA-domain u8 pictures */
} Sb2LowdelayJob;

typedef struct { const Sb2LowdelayJob *job; int t, steps; pthread_barrier_t *start, *end; } LdWorker;
/* wall-clock seconds per call, summed per thread (sb2_e2e_lowdelay_times reads and clears them) */
static double g_ld_times[64][4];
void
sb2_e2e_lowdelay_times (double *out)
{
  int t, k;
  for (k = 0; k < 4; k++) { out[k] = 0; for (t = 0; t < 64; t++) { out[k] += g_ld_times[t][k]; g_ld_times[t][k] = 0; } }
}

static void *
ld_worker (void *arg)
{
  LdWorker *w = arg;
  const Sb2LowdelayJob *j = w->job;
  const long total = (long) w->steps * j->npictures;
  long k;
  /* one untimed picture first: the thread's streams and staging buffers come into being here (cudaMalloc is a
   * device-wide stall), and go away after the end barrier (cudaFree is another) -- outside the timed region */
  schro_b200_decode_lowdelay_transform_data (j->params, j->slices[w->t % j->npictures], j->slice_bytes, j->coef_dev[w->t]);
  schro_b200_frame_inverse_iwt_combine (j->u8_dev[w->t], j->coef_dev[w->t], j->params, 0);
  schro_gpuframe_to_cpu (j->out_host[w->t % j->npictures], j->u8_dev[w->t]);
  pthread_barrier_wait (w->start);
  for (k = w->t; k < total; k += j->nthreads) {
    const int i = (int) (k % j->npictures);
    const double t0 = now ();
    schro_b200_decode_lowdelay_transform_data (j->params, j->slices[i], j->slice_bytes, j->coef_dev[w->t]);
    const double t1 = now ();
    const double t2 = now ();
    schro_b200_frame_inverse_iwt_combine (j->u8_dev[w->t], j->coef_dev[w->t], j->params, 0);
    const double t3 = now ();
    schro_gpuframe_to_cpu (j->out_host[i], j->u8_dev[w->t]);
    const double t4 = now ();
    g_ld_times[w->t & 63][0] += t1 - t0; g_ld_times[w->t & 63][1] += t2 - t1;
    g_ld_times[w->t & 63][2] += t3 - t2; g_ld_times[w->t & 63][3] += t4 - t3;
  }
  pthread_barrier_wait (w->end);
  schro_b200_thread_release ();
  return NULL;
}

double
sb2_e2e_lowdelay_run (const Sb2LowdelayJob *job, int steps)
{
  pthread_t *th = calloc ((size_t) job->nthreads, sizeof (pthread_t));
  LdWorker *w = calloc ((size_t) job->nthreads, sizeof (LdWorker));
  pthread_barrier_t start, end;
  double t0;
  int t;
  pthread_barrier_init (&start, NULL, (unsigned) job->nthreads + 1);
  pthread_barrier_init (&end, NULL, (unsigned) job->nthreads + 1);
  for (t = 0; t < job->nthreads; t++) {
    w[t].job = job; w[t].t = t; w[t].steps = steps; w[t].start = &start; w[t].end = &end;
    pthread_create (&th[t], NULL, ld_worker, &w[t]);
  }
  pthread_barrier_wait (&start);
  t0 = now ();
  pthread_barrier_wait (&end);
  t0 = now () - t0;
  for (t = 0; t < job->nthreads; t++) pthread_join (th[t], NULL);
  pthread_barrier_destroy (&start);
  pthread_barrier_destroy (&end);
  free (th);
  free (w);
  return t0;
}

/* the same pictures through the batched drop-in: every worker hands `batch` pictures at a time to
 * schro_b200_decode_lowdelay_pictures (one launch per stage and one wait per batch) */
typedef struct { const Sb2LowdelayJob *job; int t, steps, batch; pthread_barrier_t *start, *end; } LdBatchWorker;

static void *
ld_batch_worker (void *arg)
{
  LdBatchWorker *w = arg;
  const Sb2LowdelayJob *j = w->job;
  const long nb = ((long) w->steps * j->npictures) / w->batch;
  const uint8_t **data = calloc ((size_t) w->batch, sizeof (*data));
  SchroFrame **outs = calloc ((size_t) w->batch, sizeof (*outs));
  long b;
  int k;
  for (k = 0; k < w->batch; k++) { data[k] = j->slices[k % j->npictures]; outs[k] = j->out_host[(w->t * w->batch + k) % j->npictures]; }
  schro_b200_decode_lowdelay_pictures (j->params, w->batch, data, j->slice_bytes, outs, 0, 0);     /* untimed: buffers */
  pthread_barrier_wait (w->start);
  for (b = w->t; b < nb; b += j->nthreads) {
    for (k = 0; k < w->batch; k++) {
      const int i = (int) ((b * w->batch + k) % j->npictures);
      data[k] = j->slices[i];
      outs[k] = j->out_host[i];
    }
    schro_b200_decode_lowdelay_pictures (j->params, w->batch, data, j->slice_bytes, outs, 0, 0);
  }
  pthread_barrier_wait (w->end);
  free (data);
  free (outs);
  schro_b200_thread_release ();
  return NULL;
}

double
sb2_e2e_lowdelay_run_batched (const Sb2LowdelayJob *job, int steps, int batch)
{
  pthread_t *th = calloc ((size_t) job->nthreads, sizeof (pthread_t));
  LdBatchWorker *w = calloc ((size_t) job->nthreads, sizeof (LdBatchWorker));
  pthread_barrier_t start, end;
  double t0;
  int t;
  pthread_barrier_init (&start, NULL, (unsigned) job->nthreads + 1);
  pthread_barrier_init (&end, NULL, (unsigned) job->nthreads + 1);
  for (t = 0; t < job->nthreads; t++) {
    w[t].job = job; w[t].t = t; w[t].steps = steps; w[t].batch = batch; w[t].start = &start; w[t].end = &end;
    pthread_create (&th[t], NULL, ld_batch_worker, &w[t]);
  }
  pthread_barrier_wait (&start);
  t0 = now ();
  pthread_barrier_wait (&end);
  t0 = now () - t0;
  for (t = 0; t < job->nthreads; t++) pthread_join (th[t], NULL);
  pthread_barrier_destroy (&start);
  pthread_barrier_destroy (&end);
  free (th);
  free (w);
  return t0;
}
