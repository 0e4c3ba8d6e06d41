// hbm_common.cuh -- types and byte-SIMD helpers shared by the block-matching kernels
// (hbm.cu: generic one-row-per-CTA kernel and the C entry points; hbm_wave.cu: the skewed
// multi-row wavefront kernel).
#pragma once
#include "common.cuh"
#include <climits>
#include <cstdio>

namespace sb2 {

struct MotionVector {               // == SchroMotionVector (schroedinger/schromotion.h:20-37)
  uint32_t flags;
  uint32_t metric;
  uint32_t chroma_metric;
  int16_t v[4];
};

struct HbmArgs {
  PlaneSet src, ref;                // 3 u8 components each, edge-extended by `ext`
  const MotionVector *parent;       // field of level shift+1 or nullptr
  MotionVector *field;              // output field
  size_t field_pitch;               // vectors between pictures
  unsigned long long *words;        // [count][rows][cols] published results: bit 63 valid, dx<<16 | dy
  int width, height;                // luma size of this pyramid level
  int cw, ch;                       // chroma size
  int hs, vs;
  int ext;
  int bw, bh;                       // xbsep_luma, ybsep_luma
  int nbx, nby;
  int ref_index;
  int shift, h_range, use_chroma;
  int rows, cols;                   // blocks at this level: ceil(nby/skip), ceil(nbx/skip)
  int count;
  uint32_t flags0;
  unsigned *ticket;                 // CTA ticket counter (zeroed by the init launch)
};

__device__ __forceinline__ int clampi (int x, int lo, int hi) { return min (max (x, lo), hi); }

__device__ __forceinline__ unsigned warp_sum (unsigned v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned long long warp_min64 (unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync (0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// SAD of a w x h block read straight from global memory (L1/texture path), one lane
__device__ __forceinline__ unsigned block_sad (const uint8_t *a, int as, const uint8_t *b, int bs, int w, int h)
{
  unsigned s = 0;
  if (w == 8 && (((size_t) a | (size_t) as) & 7) == 0 && (bs & 3) == 0) {
    // byte-SIMD path: 8-wide source rows are 8-byte aligned (x0 is a multiple of xbsep);
    // the reference row starts anywhere, so it is assembled from aligned words
    for (int y = 0; y < h; y++) {
      const uint2 av = __ldg (reinterpret_cast<const uint2 *> (a + (ptrdiff_t) y * as));   // needs 8-byte alignment
      const uint8_t *br = b + (ptrdiff_t) y * bs;
      const size_t mis = (size_t) br & 3;
      const unsigned *bw_ = reinterpret_cast<const unsigned *> (br - mis);
      const unsigned w0 = __ldg (bw_), w1 = __ldg (bw_ + 1), w2 = mis ? __ldg (bw_ + 2) : 0u;
      const unsigned sh = (unsigned) mis * 8;
      const unsigned b0 = __funnelshift_r (w0, w1, sh), b1 = __funnelshift_r (w1, w2, sh);
      s += __vsadu4 (av.x, b0) + __vsadu4 (av.y, b1);
    }
    return s;
  }
  for (int y = 0; y < h; y++) {
    const uint8_t *ar = a + (ptrdiff_t) y * as, *br = b + (ptrdiff_t) y * bs;
    for (int x = 0; x < w; x++) s += (unsigned) abs ((int) __ldg (ar + x) - (int) __ldg (br + x));
  }
  return s;
}

// Published block results.  A row's neighbours in the row below need only the vector of a
// finished block, so the vector itself is the flag: one relaxed 64-bit word per block
// (bit 63 = valid, dx in bits 16..31, dy in bits 0..15).  No fence on either side -- the
// word is the data, single-copy atomic -- which takes two L2 round trips and two fences off
// the per-block critical path compared with a progress counter + separate vector load.
__device__ __forceinline__ unsigned long long ld_word (const unsigned long long *p)
{
  unsigned long long v;
  asm volatile ("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_word (unsigned long long *p, unsigned long long v)
{
  asm volatile ("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pack_word (int dx, int dy)
{
  return (1ull << 63) | ((unsigned long long) (dx & 0xffff) << 16) | (unsigned long long) (dy & 0xffff);
}

// 8 (or 4) bytes starting at any address, assembled from aligned 32-bit words
__device__ __forceinline__ uint2 load8_unaligned (const uint8_t *p)
{
  const size_t mis = (size_t) p & 3;
  const unsigned *w = reinterpret_cast<const unsigned *> (p - mis);
  const unsigned w0 = __ldg (w), w1 = __ldg (w + 1), w2 = mis ? __ldg (w + 2) : 0u;
  const unsigned sh = (unsigned) mis * 8;
  return make_uint2 (__funnelshift_r (w0, w1, sh), __funnelshift_r (w1, w2, sh));
}
__device__ __forceinline__ unsigned load4_unaligned (const uint8_t *p)
{
  const size_t mis = (size_t) p & 3;
  const unsigned *w = reinterpret_cast<const unsigned *> (p - mis);
  const unsigned w0 = __ldg (w), w1 = mis ? __ldg (w + 1) : 0u;
  return __funnelshift_r (w0, w1, (unsigned) mis * 8);
}

// The same with the alignment work hoisted out of the row loop: every row of a block starts at
// the same byte offset inside its word (strides are multiples of 4), so the aligned base, the
// shift and "needs a third word" are computed once per block position.
struct RowRef { const unsigned *w; unsigned sh; bool three; };
__device__ __forceinline__ RowRef row_ref (const uint8_t *p)
{
  RowRef r;
  const unsigned mis = (unsigned) ((size_t) p & 3);
  r.w = reinterpret_cast<const unsigned *> (p - mis);
  r.sh = mis * 8;
  r.three = mis != 0;
  return r;
}
__device__ __forceinline__ uint2 row_load8 (const RowRef &r, int word_off)
{
  const unsigned w0 = __ldg (r.w + word_off), w1 = __ldg (r.w + word_off + 1), w2 = r.three ? __ldg (r.w + word_off + 2) : 0u;
  return make_uint2 (__funnelshift_r (w0, w1, r.sh), __funnelshift_r (w1, w2, r.sh));
}
__device__ __forceinline__ unsigned row_load4 (const RowRef &r, int word_off)
{
  const unsigned w0 = __ldg (r.w + word_off), w1 = r.three ? __ldg (r.w + word_off + 1) : 0u;
  return __funnelshift_r (w0, w1, r.sh);
}


// hbm_wave.cu: static-candidate pre-pass + wavefront kernel for the codec's usual geometry
// (8x8 blocks, 4:2:0, luma-only scan).  Returns false when the geometry is not covered.
bool hbm_wave_supported (const HbmArgs &A, int h_range);
size_t hbm_wave_workspace_bytes (int rows, int cols, int count);
int hbm_wave_launch (const HbmArgs &A, int h_range, void *workspace, size_t workspace_bytes, cudaStream_t st,
    double bytes);

}  // namespace sb2
