// launch_rate_probe.cu -- how many kernel launches per second can one process issue from T host threads,
// each on its own stream?  (Is the per-picture e2e leg bound by CUDA calls?  DESIGN.md 6.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/launch_rate_probe tools/launch_rate_probe.cu -lpthread
#include <cstdio>
#include <cstdlib>
#include <pthread.h>
#include <time.h>
#include <cuda_runtime.h>

__global__ void tiny (int *p) { if (p && threadIdx.x == 9999) *p = 1; }

static int g_n = 20000, g_sync_every = 0;
static pthread_barrier_t g_bar;
static double now () { timespec a; clock_gettime (CLOCK_MONOTONIC, &a); return a.tv_sec + 1e-9 * a.tv_nsec; }

static void *worker (void *)
{
  cudaStream_t st;
  cudaEvent_t ev;
  cudaStreamCreateWithFlags (&st, cudaStreamNonBlocking);
  cudaEventCreateWithFlags (&ev, cudaEventDisableTiming | cudaEventBlockingSync);
  tiny<<<1, 32, 0, st>>> (nullptr);
  cudaStreamSynchronize (st);
  pthread_barrier_wait (&g_bar);
  for (int i = 0; i < g_n; i++) {
    tiny<<<1, 32, 0, st>>> (nullptr);
    if (g_sync_every && (i + 1) % g_sync_every == 0) { cudaEventRecord (ev, st); cudaEventSynchronize (ev); }
  }
  cudaStreamSynchronize (st);
  pthread_barrier_wait (&g_bar);
  return nullptr;
}

int main ()
{
  cudaFree (0);
  for (int sync_every : { 0, 40 }) {
    g_sync_every = sync_every;
    for (int T : { 1, 4, 16, 32, 64 }) {
      pthread_t th[64];
      pthread_barrier_init (&g_bar, nullptr, T + 1);
      for (int t = 0; t < T; t++) pthread_create (&th[t], nullptr, worker, nullptr);
      pthread_barrier_wait (&g_bar);
      const double t0 = now ();
      pthread_barrier_wait (&g_bar);
      const double dt = now () - t0;
      for (int t = 0; t < T; t++) pthread_join (th[t], nullptr);
      pthread_barrier_destroy (&g_bar);
      printf ("%2d threads, %s: %8.0f launches/s in total (%.2f us per launch per thread)\n", T,
          sync_every ? "a blocking event wait every 40 launches" : "no waits", T * (double) g_n / dt, dt / g_n * 1e6);
    }
  }
  return 0;
}
