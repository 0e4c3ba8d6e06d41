// vabsdiff_probe.cu -- measured peak of VABSDIFF4.U8.ACC on this GPU (the roofline of the SAD kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/vabsdiff_probe tools/vabsdiff_probe.cu && /tmp/vabsdiff_probe
// Every thread runs ILP independent accumulate chains; the result is lane-operations per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__ (256) probe (unsigned *out, unsigned a0, unsigned b0, int iters)
{
  unsigned acc[ILP], a = a0 + threadIdx.x, b = b0 ^ blockIdx.x;
#pragma unroll
  for (int k = 0; k < ILP; k++) acc[k] = k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < ILP; k++)
        asm volatile ("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a + k), "r"(b + u));
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static void run (unsigned *out, int sms, int clock_khz)
{
  const int iters = 4096, ctas = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate (&e0);
  cudaEventCreate (&e1);
  probe<ILP><<<ctas, 256>>> (out, 1, 2, 16);
  cudaEventRecord (e0);
  probe<ILP><<<ctas, 256>>> (out, 1, 2, iters);
  cudaEventRecord (e1);
  cudaEventSynchronize (e1);
  float ms;
  cudaEventElapsedTime (&ms, e0, e1);
  const double ops = (double) ctas * 256 * iters * 8 * ILP;
  printf ("ILP %d: %.3f ms, %.2f T lane-ops/s, %.1f lanes/clk/SM at %d MHz (max clock)\n", ILP, ms, ops / ms * 1e-9,
      ops / (ms * 1e-3) / sms / (clock_khz * 1e3), clock_khz / 1000);
}

int main ()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties (&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute (&khz, cudaDevAttrClockRate, 0);
  unsigned *out;
  cudaMalloc (&out, (size_t) p.multiProcessorCount * 8 * 256 * 4);
  printf ("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<1> (out, p.multiProcessorCount, khz);
  run<4> (out, p.multiProcessorCount, khz);
  run<8> (out, p.multiProcessorCount, khz);
  return 0;
}
