#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
__global__ void probe (const __grid_constant__ CUtensorMap tm, float *out, int nbytes)
{
  __shared__ alignas (128) float buf[64 * 64];
  __shared__ alignas (8) unsigned long long bar;
  const unsigned b = (unsigned) __cvta_generic_to_shared (&bar);
  const unsigned dst = (unsigned) __cvta_generic_to_shared (buf);
  if (threadIdx.x == 0) {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads ();
  if (threadIdx.x == 0) {
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(nbytes) : "memory");
    asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(dst), "l"(&tm), "r"(0), "r"(0), "r"(b) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x) out[i] = buf[i];
}
int main ()
{
  float *dev, *out;
  cudaMalloc (&dev, 1024 * 1024 * 4); cudaMalloc (&out, 64 * 64 * 4);
  cudaMemset (dev, 0, 1024 * 1024 * 4);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t ge = cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  printf ("entry point: %d %d %p\n", (int) ge, (int) q, fp);
  auto enc = reinterpret_cast<CUresult (*) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)> (fp);
  CUtensorMap tm;
  const cuuint64_t dims[2] = { 1024, 1024 }, strides[1] = { 4096 };
  const cuuint32_t box[2] = { 64, 64 }, es[2] = { 1, 1 };
  printf ("enc %d\n", (int) enc (&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  for (int i = 0; i < 16; i++) printf ("%016llx ", (unsigned long long) tm.opaque[i]);
  printf ("\n");
  probe<<<1, 128>>> (tm, out, 64 * 64 * 4);
  printf ("kernel: %s\n", cudaGetErrorString (cudaDeviceSynchronize ()));
  int drv = 0, rt = 0; cudaDriverGetVersion (&drv); cudaRuntimeGetVersion (&rt);
  printf ("driver %d runtime %d\n", drv, rt);
  return 0;
}
