// probe: 4-D tensor map over a (x, phase, y, picture) byte tensor, one bulk tensor copy into shared memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
constexpr int RW = 112, RH = 104, RB = 4 * RW * RH;
struct Maps { CUtensorMap m[2][3]; };
__global__ void probe (const __grid_constant__ Maps maps, int comp, int x, int y, int pic, unsigned char *out)
{
  extern __shared__ unsigned char raw[];
  unsigned char *buf = reinterpret_cast<unsigned char *> ((reinterpret_cast<size_t> (raw) + 127) & ~(size_t) 127);
  __shared__ alignas (8) unsigned long long bar;
  if (threadIdx.x == 0) {
    const unsigned b = (unsigned) __cvta_generic_to_shared (&bar);
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(RB) : "memory");
    const unsigned dst = (unsigned) __cvta_generic_to_shared (buf);
    const CUtensorMap *tm = comp == 0 ? &maps.m[0][0] : comp == 1 ? &maps.m[0][1] : &maps.m[0][2];
    asm volatile ("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        :: "r"(dst), "l"(tm), "r"(x), "r"(0), "r"(y), "r"(pic), "r"(b) : "memory");
  }
  __syncthreads ();
  {
    const unsigned b = (unsigned) __cvta_generic_to_shared (&bar);
    unsigned done = 0;
    while (!done)
      asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
  }
  for (int i = threadIdx.x; i < RB; i += blockDim.x) out[i] = buf[i];
}
int main (int argc, char **argv)
{
  const int w = argc > 1 ? atoi (argv[1]) : 64, h = argc > 2 ? atoi (argv[2]) : 48, count = 2;
  const int pitch = ((w + 64) + 15) & ~15, stride = 4 * pitch, rows = h + 64;
  const size_t pic_pitch = ((size_t) stride * rows + 255) & ~(size_t) 255;
  std::vector<unsigned char> host (pic_pitch * count);
  for (size_t i = 0; i < host.size (); i++) host[i] = (unsigned char) (i * 2654435761u >> 13);
  unsigned char *dev, *out;
  cudaMalloc (&dev, host.size ()); cudaMalloc (&out, RB);
  cudaMemcpy (dev, host.data (), host.size (), cudaMemcpyHostToDevice);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)> (fp);
  Maps maps; memset (&maps, 0, sizeof (maps));
  const cuuint64_t dims[4] = { (cuuint64_t) (w + 64), 4, (cuuint64_t) rows, (cuuint64_t) count };
  const cuuint64_t strides[3] = { (cuuint64_t) pitch, (cuuint64_t) stride, (cuuint64_t) pic_pitch };
  const cuuint32_t box[4] = { RW, 4, RH, 1 }, es[4] = { 1, 1, 1, 1 };
  for (int c = 0; c < 3; c++) {
    CUresult r = enc (&maps.m[0][c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf ("encode comp %d -> %d (dims %d x 4 x %d x %d, pitch %d)\n", c, (int) r, w + 64, rows, count, pitch);
  }
  cudaFuncSetAttribute (probe, cudaFuncAttributeMaxDynamicSharedMemorySize, RB + 256);
  const int x = 8, y = 12, pic = 1;
  probe<<<1, 256, RB + 256>>> (maps, 1, x, y, pic, out);
  cudaError_t e = cudaDeviceSynchronize ();
  printf ("kernel: %s\n", cudaGetErrorString (e));
  if (e != cudaSuccess) return 1;
  std::vector<unsigned char> got (RB);
  cudaMemcpy (got.data (), out, RB, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int ph = 0; ph < 4; ph++) for (int r = 0; r < RH; r++) for (int c = 0; c < RW; c++) {
    const int gx = x + c, gy = y + r;
    unsigned char want = 0;
    if (gx < w + 64 && gy < rows) want = host[(size_t) pic * pic_pitch + (size_t) gy * stride + (size_t) ph * pitch + gx];
    if (got[(ph * RH + r) * RW + c] != want) bad++;
  }
  printf ("mismatches: %ld of %d\n", bad, RB);
  return bad != 0;
}
