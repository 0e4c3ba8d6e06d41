// obmc_common.cuh -- types and device helpers shared by the OBMC kernels (obmc.cu: the generic
// per-pixel kernel, the scatter kernel and the C entry point; obmc_blocks.cu: the block-per-warp
// kernel on TMA-staged reference regions).
#pragma once
#include "common.cuh"
#include <cstdio>

namespace sb2 {

struct MotionVector {               // == SchroMotionVector, schroedinger/schromotion.h:20-37
  uint32_t flags;                   // pred_mode:2 using_global:1 split:2 unused:3 scan:8
  uint32_t metric;
  uint32_t chroma_metric;
  int16_t v[4];                     // vec: dx0 dx1 dy0 dy1 / dc: dc0 dc1 dc2
};
static_assert (sizeof (MotionVector) == 20, "SchroMotionVector is 20 bytes");

struct ObmcArgs {
  PlaneSet ref0, ref1, acc, res, out;
  const MotionVector *mvs;
  size_t mv_pitch;                  // vectors between consecutive pictures
  int w[SB2_MAX_COMPONENTS], h[SB2_MAX_COMPONENTS];
  int xbsep[SB2_MAX_COMPONENTS], ybsep[SB2_MAX_COMPONENTS];
  int xblen[SB2_MAX_COMPONENTS], yblen[SB2_MAX_COMPONENTS];
  int hs[SB2_MAX_COMPONENTS], vs[SB2_MAX_COMPONENTS];
  unsigned char wx[SB2_MAX_COMPONENTS][64], wy[SB2_MAX_COMPONENTS][64];
  int nbx, nby, prec, w1, w2, bits;
  int ncomp, add, res_is_s32, has_ref1, has_acc;
};

__device__ __forceinline__ int w16 (int x) { return (int) (short) x; }
__device__ __forceinline__ int clampi (int x, int lo, int hi) { return min (max (x, lo), hi); }

// half-pel sample (u,v) + block pixel (a,b): phase ((v&1)<<1)|(u&1) at (u>>1, v>>1)
// (schroedinger/schroframe.c:2186-2200)
__device__ __forceinline__ int halfpel (const uint8_t *ref, int rstride, int u, int v, int a, int b)
{
  const int ph = ((v & 1) << 1) | (u & 1);
  return __ldg (ref + (ptrdiff_t) ph * (rstride >> 2) + (ptrdiff_t) ((v >> 1) + b) * rstride + (u >> 1) + a);
}

// schromotion8.c:303-335 + schroframe.c:2288-2482
__device__ __forceinline__ int fetch (const uint8_t *ref, int rstride, int prec, int bx, int by,
    int dx, int dy, int max_fast_x, int max_fast_y, int a, int b)
{
  int px = (bx << prec) + dx, py = (by << prec) + dy;
  const int e = 32 << prec;
  px = clampi (px, -e, max_fast_x + e - 1);
  py = clampi (py, -e, max_fast_y + e - 1);
  if (prec == 0) return __ldg (ref + (ptrdiff_t) (py + b) * rstride + px + a);
  if (prec == 1) return halfpel (ref, rstride, px, py, a, b);
  if (prec == 2) { px <<= 1; py <<= 1; }
  const int hx = px >> 2, hy = py >> 2, rx = px & 3, ry = py & 3;
  const int s00 = halfpel (ref, rstride, hx, hy, a, b);
  if ((rx | ry) == 0) return s00;
  if (ry == 0 && rx == 2) return (s00 + halfpel (ref, rstride, hx + 1, hy, a, b) + 1) >> 1;
  if (ry == 2 && rx == 0) return (s00 + halfpel (ref, rstride, hx, hy + 1, a, b) + 1) >> 1;
  // orc_combine4_nxm_u8 (schroorc.orc:1635-1662): weights sum to 16, fits 16 bits
  const int s01 = halfpel (ref, rstride, hx + 1, hy, a, b);
  const int s10 = halfpel (ref, rstride, hx, hy + 1, a, b);
  const int s11 = halfpel (ref, rstride, hx + 1, hy + 1, a, b);
  return ((4 - ry) * (4 - rx) * s00 + (4 - ry) * rx * s01 + ry * (4 - rx) * s10 + ry * rx * s11 + 8) >> 4;
}

// ---- block table + unified 4-tap fetch ------------------------------------------------
// Every sub-pel case of schroframe.c:2288-2413 is the same 4-tap sum
//   (w00*s00 + w01*s01 + w10*s10 + w11*s11 + 8) >> 4,  weights summing to 16:
// the copy case is w00 = 16, the two avgub cases are 8/8 ((8a+8b+8)>>4 == (a+b+1)>>1),
// prec 0/1 are single taps.  So the per-block work (vector decode, clamp, phase
// selection, weights) is done once per CTA into a table and every pixel issues up to
// 2x2 blocks x 2 refs x 4 taps of independent loads before any arithmetic.
struct BlkRef { int o[4]; unsigned w; };
struct BlkEnt { BlkRef r[2]; short mode, fast, dc, pad; };
constexpr int MAX_ENT = 256;

__device__ __forceinline__ void make_blkref (BlkRef &br, int rstride, int prec, int bx, int by, int dx, int dy,
    int max_fast_x, int max_fast_y)
{
  int px = (bx << prec) + dx, py = (by << prec) + dy;
  const int e = 32 << prec;
  px = clampi (px, -e, max_fast_x + e - 1);
  py = clampi (py, -e, max_fast_y + e - 1);
  const int q = rstride >> 2;
  if (prec == 0) {
    br.o[0] = br.o[1] = br.o[2] = br.o[3] = py * rstride + px;
    br.w = 16u;
    return;
  }
  int rx = 0, ry = 0, hx = px, hy = py;
  if (prec >= 2) {
    if (prec == 2) { px <<= 1; py <<= 1; }
    hx = px >> 2; hy = py >> 2; rx = px & 3; ry = py & 3;
  }
  // half-pel sample (u,v): phase ((v&1)<<1)|(u&1) at (u>>1, v>>1)
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int u = hx + (t & 1), v = hy + (t >> 1);
    br.o[t] = (((v & 1) << 1) | (u & 1)) * q + (v >> 1) * rstride + (u >> 1);
  }
  const unsigned w00 = (4 - ry) * (4 - rx), w01 = (4 - ry) * rx, w10 = ry * (4 - rx), w11 = ry * rx;
  br.w = w00 | (w01 << 8) | (w10 << 16) | (w11 << 24);
}

// one pixel through the 4-tap sum (taps with zero weight are not loaded)
__device__ __forceinline__ int fetch1 (const uint8_t *ref, const BlkRef &br, int pix)
{
  const unsigned w = br.w;
  int acc = 8;
  acc += (int) (w & 0xff) * (int) __ldg (ref + br.o[0] + pix);
  if (w & 0x0000ff00u) acc += (int) ((w >> 8) & 0xff) * (int) __ldg (ref + br.o[1] + pix);
  if (w & 0x00ff0000u) acc += (int) ((w >> 16) & 0xff) * (int) __ldg (ref + br.o[2] + pix);
  if (w & 0xff000000u) acc += (int) (w >> 24) * (int) __ldg (ref + br.o[3] + pix);
  return acc >> 4;
}

template <bool SIMPLE>
__device__ __forceinline__ int obmc_combine (const ObmcArgs &A, int mode, bool fast, bool noscale, int dc, int s0, int s1)
{
  if (SIMPLE) {
    const int avg = (s0 + s1 + 1) >> 1;
    const int one = (mode == 1) ? s0 : s1;
    const int dcv = fast ? w16 (dc + 128) : ((dc + 128) & 0xff);
    return mode == 0 ? dcv : (mode == 3 ? avg : one);
  }
  if (mode == 0) return fast ? w16 (dc + 128) : ((dc + 128) & 0xff);
  if (mode == 3) {
    if (fast) {
      int t = w16 (s0 * w16 (A.w1 << (6 - A.bits)));
      const int u = w16 (s1 * w16 (A.w2 << (6 - A.bits)));
      t = w16 (t + u);
      t = w16 (t + 32);
      return t >> 6;
    }
    int t = w16 (s0 * w16 (A.w1));
    const int u = w16 (s1 * w16 (A.w2));
    t = w16 (t + u);
    t = w16 (t + ((1 << A.bits) >> 1));
    return clampi (t >> A.bits, 0, 255);
  }
  const int s = (mode == 1) ? s0 : s1;
  if (fast) {
    int t = w16 (s * w16 ((A.w1 + A.w2) << (6 - A.bits)));
    t = w16 (t + 32);
    return t >> 6;
  }
  if (noscale) return s;
  return ((s * (A.w1 + A.w2) + (1 << (A.bits - 1))) >> A.bits) & 0xff;
}

// 4 bytes starting at any address, assembled from aligned 32-bit words
__device__ __forceinline__ unsigned ldg_u32_unaligned (const uint8_t *p)
{
  const size_t mis = (size_t) p & 3;
  const unsigned *w = reinterpret_cast<const unsigned *> (p - mis);
  const unsigned w0 = __ldg (w), w1 = mis ? __ldg (w + 1) : 0u;
  return __funnelshift_r (w0, w1, (unsigned) mis * 8);
}

// 4-tap sum of four adjacent pixels, two per packed 16-bit pair: returns (p0 | p1<<16, p2 | p3<<16)
__device__ __forceinline__ uint2 fetch4x4 (const uint8_t *ref, const BlkRef &br, int pix)
{
  const unsigned w = br.w;
  unsigned lo = 0x00080008u, hi = 0x00080008u;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const unsigned wt = (w >> (8 * t)) & 0xff;
    if (t == 0 || wt) {
      const unsigned b = ldg_u32_unaligned (ref + br.o[t] + pix);
      lo += wt * __byte_perm (b, 0, 0x4140);      // byte0 | byte1 << 16
      hi += wt * __byte_perm (b, 0, 0x4342);      // byte2 | byte3 << 16
    }
  }
  return make_uint2 ((lo >> 4) & 0x0fff0fffu, (hi >> 4) & 0x0fff0fffu);
}


// obmc_blocks.cu: reference regions staged by TMA, one block per warp pass, for blocks of at most 32
// (row, 8-pixel item) lanes on frames with a 32-pixel border.  Returns SB2_OK when it launched,
// SB2_ERR_UNSUPPORTED when the caller should take another kernel.
int obmc_blocks_launch (const ObmcArgs &A, const sb2_slab *ref0, const sb2_slab *ref1, int count, bool staged, cudaStream_t st);

}  // namespace sb2
