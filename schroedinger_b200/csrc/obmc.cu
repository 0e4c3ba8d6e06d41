// obmc.cu -- overlapped-block motion-compensation renderer for sm_100a.
//
// Bit-exact replacement for schro_motion_render / schro_motion_render_u8
// (schroedinger/schromotion.c:95-155, schroedinger/schromotion8.c:700-929).
//
// The reference scatters block by block into an s16 accumulator strip.  Here every
// output pixel GATHERS the (at most four, for xblen <= 2*xbsep) blocks that cover it:
// fetches the sub-pel reference sample from the four half-pel phase planes, applies the
// prediction mode, multiplies by the OBMC window and finishes (residual add + clamp, or
// residual subtract) in registers -- the accumulator never touches memory unless the
// caller asks for it.  Equality with the reference holds modulo 2^16, which is exactly
// what its Orc addw/mullw accumulate (schromotion8.c:15-167).

#include "obmc_common.cuh"
#include <cstdlib>

namespace sb2 {

__global__ void __launch_bounds__ (256)
obmc_kernel (const ObmcArgs A)
{
  const int comp = blockIdx.z % A.ncomp, pic = blockIdx.z / A.ncomp;
  const int width = A.w[comp], height = A.h[comp];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, (width - xoff) / xbsep);
  const int max_y_blocks = min (A.nby - 1, (height - yoff) / ybsep);
  const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));
  const unsigned char *wx = A.wx[comp], *wy = A.wy[comp];

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  const int j0 = (y + yoff - yblen + 1 > 0) ? (y + yoff - yblen + ybsep) / ybsep : 0;
  const int j1 = min (A.nby - 1, (y + yoff) / ybsep);
  const int i0 = (x + xoff - xblen + 1 > 0) ? (x + xoff - xblen + xbsep) / xbsep : 0;
  const int i1 = min (A.nbx - 1, (x + xoff) / xbsep);

  int sum = 0;
  for (int j = j0; j <= j1; j++) {
    const int by = ybsep * j - yoff, b = y - by;
    for (int i = i0; i <= i1; i++) {
      const int bx = xbsep * i - xoff, a = x - bx;
      const MotionVector *mv = mvs + (size_t) j * A.nbx + i;
      // 20-byte struct: flags at +0, vectors at +12 (4-byte aligned)
      const unsigned flags = __ldg (&mv->flags);
      const int2 vv = make_int2 (__ldg (reinterpret_cast<const int *> (mv->v)),
          __ldg (reinterpret_cast<const int *> (mv->v) + 1));
      const int v0 = (short) (vv.x & 0xffff), v1 = vv.x >> 16, v2 = (short) (vv.y & 0xffff), v3 = vv.y >> 16;
      const int mode = flags & 3;
      const bool fast = (i >= 1 && i < max_x_blocks && j >= 1 && j < max_y_blocks);
      int v;
      if (mode == 0) {
        const int dc = (comp == 0 ? v0 : comp == 1 ? v1 : v2) + 128;
        v = fast ? w16 (dc) : (dc & 0xff);
      } else if (mode == 3) {
        const int s0 = fetch (ref0, rs0, prec, bx, by, v0 >> A.hs[comp], v2 >> A.vs[comp], max_fast_x, max_fast_y, a, b);
        const int s1 = fetch (ref1, rs1, prec, bx, by, v1 >> A.hs[comp], v3 >> A.vs[comp], max_fast_x, max_fast_y, a, b);
        if (simple) {
          v = (s0 + s1 + 1) >> 1;
        } else if (fast) {          // block_acc_biref, schromotion8.c:127-167
          int t = w16 (s0 * w16 (A.w1 << (6 - A.bits)));
          const int u = w16 (s1 * w16 (A.w2 << (6 - A.bits)));
          t = w16 (t + u);
          t = w16 (t + 32);
          v = t >> 6;
        } else {                    // orc_combine2_nxm_u8, schroorc.orc:1737-1756
          int t = w16 (s0 * w16 (A.w1));
          const int u = w16 (s1 * w16 (A.w2));
          t = w16 (t + u);
          t = w16 (t + ((1 << A.bits) >> 1));
          v = clampi (t >> A.bits, 0, 255);
        }
      } else {
        const int s = (mode == 1)
            ? fetch (ref0, rs0, prec, bx, by, v0 >> A.hs[comp], v2 >> A.vs[comp], max_fast_x, max_fast_y, a, b)
            : fetch (ref1, rs1, prec, bx, by, v1 >> A.hs[comp], v3 >> A.vs[comp], max_fast_x, max_fast_y, a, b);
        if (fast) {
          if (simple) v = s;
          else {                    // block_acc_scaled, schromotion8.c:41-71
            int t = w16 (s * w16 ((A.w1 + A.w2) << (6 - A.bits)));
            t = w16 (t + 32);
            v = t >> 6;
          }
        } else {
          if (noscale) v = s;       // schromotion8.c:384-398
          else v = ((s * (A.w1 + A.w2) + (1 << (A.bits - 1))) >> A.bits) & 0xff;
        }
      }
      int w_x = wx[a], w_y = wy[b];
      if (!fast) {
        // border blocks absorb the weight of the missing neighbour (schromotion8.c:673-693)
        if (x < xoff) w_x += wx[2 * xoff - a - 1];
        if (x >= A.nbx * xbsep - xoff) w_x += wx[2 * (xblen - xoff) - a - 1];
        if (y < yoff) w_y += wy[2 * yoff - b - 1];
        if (y >= A.nby * ybsep - yoff) w_y += wy[2 * (yblen - yoff) - b - 1];
      }
      sum += v * w_x * w_y;
    }
  }

  const int a16 = w16 (sum);
  if (A.add) {
    // orc_rrshift6_add_s16_2d / _s32_2d (schroorc.orc:636-660)
    const char *rrow = plane_ptr (A.res, pic, comp) + (size_t) y * A.res.stride[comp];
    const int r = A.res_is_s32 ? w16 (reinterpret_cast<const int *> (rrow)[x])
                               : (int) reinterpret_cast<const short *> (rrow)[x];
    int t = w16 (a16 + 32) >> 6;
    t = w16 (r + t);
    reinterpret_cast<uint8_t *> (plane_ptr (A.out, pic, comp))[(size_t) y * A.out.stride[comp] + x] =
        (uint8_t) clampi (t, 0, 255);
    if (A.has_acc)
      reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp])[x] = (short) a16;
  } else {
    // orc_rrshift6_sub_s16_2d (schroorc.orc:663-673)
    short *r = reinterpret_cast<short *> (plane_ptr (A.res, pic, comp) + (size_t) y * A.res.stride[comp]) + x;
    const int t = w16 (a16 - 8160) >> 6;
    *r = (short) w16 (*r - t);
    if (A.has_acc)
      reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp])[x] = (short) t;
  }
}

// ---- v4: block-major scatter into a shared-memory accumulator ---------------------------
// The gather kernels above are bound by L1 wavefronts: neighbouring pixels belong to
// different blocks with different vectors, so every byte load of a warp touches ~16 cache
// lines.  Here the work item is (block, block row, group of 4 pixels): the three items of
// a block row read 12 contiguous reference bytes, each as one unaligned 32-bit word built
// from two aligned loads; taps are applied to two pixels at a time in packed 16-bit lanes.
// Contributions are added into a tile accumulator in shared memory with atomics (integer
// adds commute, so the sum is bit-exact whatever the order); a thread keeps its (row, group)
// for every block it visits, so the item loop has no divisions and no barriers.
#ifndef OBMC_MINB
#define OBMC_MINB 6     // 40 registers: latency of the scattered reference loads is hidden by resident warps
#endif
constexpr int O4_W = 64, O4_H = 32;
constexpr int O4_P = O4_W + 12;    // accumulator pitch: 16-byte aligned rows, consecutive block rows on different banks

template <bool SIMPLE>
__global__ void __launch_bounds__ (256, OBMC_MINB)
obmc_kernel_v4 (const ObmcArgs A)
{
  __shared__ __align__ (16) BlkEnt tab[MAX_ENT];
  __shared__ __align__ (16) int acc[O4_H][O4_P];
  __shared__ unsigned char s_wx[64], s_wy[64];
  const int comp = blockIdx.z % A.ncomp, pic = blockIdx.z / A.ncomp;
  const int width = A.w[comp], height = A.h[comp];
  const int tx0 = blockIdx.x * O4_W, ty0 = blockIdx.y * O4_H;
  if (tx0 >= width || ty0 >= height) return;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  // block separations are powers of two in every Dirac preset: shifts instead of the ~25-instruction
  // software divide that every thread of the CTA would otherwise run six times (operands are >= 0)
  const int xsh = (xbsep & (xbsep - 1)) == 0 ? __ffs (xbsep) - 1 : -1;
  const int ysh = (ybsep & (ybsep - 1)) == 0 ? __ffs (ybsep) - 1 : -1;
  auto divx = [&] (int v) { return xsh >= 0 ? v >> xsh : v / xbsep; };
  auto divy = [&] (int v) { return ysh >= 0 ? v >> ysh : v / ybsep; };
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, divx (width - xoff));
  const int max_y_blocks = min (A.nby - 1, divy (height - yoff));
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));

  if (threadIdx.x < 64) {
    s_wx[threadIdx.x] = A.wx[comp][threadIdx.x];
    s_wy[threadIdx.x] = A.wy[comp][threadIdx.x];
  }
  for (int t = threadIdx.x; t < O4_P * O4_H / 4; t += blockDim.x) reinterpret_cast<int4 *> (&acc[0][0])[t] = make_int4 (0, 0, 0, 0);

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  const int tw = min (O4_W, width - tx0), th = min (O4_H, height - ty0);
  const int x1 = tx0 + tw - 1, y1 = ty0 + th - 1;
  const int ti0 = (tx0 + xoff - xblen + 1 > 0) ? divx (tx0 + xoff - xblen + xbsep) : 0;
  const int ti1 = min (A.nbx - 1, divx (x1 + xoff));
  const int tj0 = (ty0 + yoff - yblen + 1 > 0) ? divy (ty0 + yoff - yblen + ybsep) : 0;
  const int tj1 = min (A.nby - 1, divy (y1 + yoff));
  const int tni = ti1 - ti0 + 1, tnj = tj1 - tj0 + 1;

  for (int t = threadIdx.x; t < tni * tnj; t += blockDim.x) {
    const int jj = t / tni, ii = t - jj * tni;
    const int i = ti0 + ii, j = tj0 + jj;
    const MotionVector *mv = mvs + (size_t) j * A.nbx + i;
    const unsigned flags = __ldg (&mv->flags);
    const int v01 = __ldg (reinterpret_cast<const int *> (mv->v)), v23 = __ldg (reinterpret_cast<const int *> (mv->v) + 1);
    const int v0 = (short) (v01 & 0xffff), v1 = v01 >> 16, v2 = (short) (v23 & 0xffff), v3 = v23 >> 16;
    BlkEnt e;
    e.mode = (short) (flags & 3);
    e.fast = (short) (i >= 1 && i < max_x_blocks && j >= 1 && j < max_y_blocks);
    e.dc = (short) (comp == 0 ? v0 : comp == 1 ? v1 : v2);
    e.pad = 0;
    const int bx = xbsep * i - xoff, by = ybsep * j - yoff;
    make_blkref (e.r[0], rs0, prec, bx, by, v0 >> A.hs[comp], v2 >> A.vs[comp], max_fast_x, max_fast_y);
    make_blkref (e.r[1], rs1, prec, bx, by, v1 >> A.hs[comp], v3 >> A.vs[comp], max_fast_x, max_fast_y);
    tab[t] = e;
  }
  __syncthreads ();

  // ---- scatter: every (block, block row, 4-pixel group) item adds into the tile accumulator.
  // A thread keeps the same (row, group) for every block it visits, so no per-item divisions;
  // overlapping blocks meet in shared-memory atomics (integer adds commute: bit-exact in any order).
  const int gx = (xblen + 3) >> 2;                 // 4-pixel groups per block row
  const int ipb = gx * yblen;                      // items per block
  const int nslot = blockDim.x / ipb;              // blocks in flight per pass
  const int slot = threadIdx.x / ipb, rem = threadIdx.x - slot * ipb;
  const int r = rem / gx, g = rem - r * gx;
  if (slot < nslot) {
    int bi_ = slot % tni, bj = slot / tni;
    const int step_i = nslot % tni, step_j = nslot / tni;
    for (; bj < tnj; ) {
      const int i = ti0 + bi_, j = tj0 + bj;
      const int bx = xbsep * i - xoff, by = ybsep * j - yoff;
      const int y = by + r, xg = bx + 4 * g;
      if (!(y < ty0 || y > y1 || xg > x1 || xg + 3 < tx0)) {
        const BlkEnt &e = tab[bj * tni + bi_];
        const int mode = e.mode;
        const bool fast = e.fast != 0;
        int v[4];
        if (mode == 0) {
          const int dcv = fast ? w16 ((int) e.dc + 128) : (((int) e.dc + 128) & 0xff);
          v[0] = v[1] = v[2] = v[3] = dcv;
        } else {
          uint2 s0 = make_uint2 (0, 0), s1 = make_uint2 (0, 0);
          if (mode & 1) s0 = fetch4x4 (ref0, e.r[0], r * rs0 + 4 * g);
          if (mode & 2) s1 = fetch4x4 (ref1, e.r[1], r * rs1 + 4 * g);
          if (SIMPLE) {
            uint2 p;
            if (mode == 3) {
              p.x = ((s0.x + s1.x + 0x00010001u) >> 1) & 0x00ff00ffu;     // avgub, two lanes
              p.y = ((s0.y + s1.y + 0x00010001u) >> 1) & 0x00ff00ffu;
            } else {
              p = (mode == 1) ? s0 : s1;
            }
            v[0] = p.x & 0xffff; v[1] = p.x >> 16; v[2] = p.y & 0xffff; v[3] = p.y >> 16;
          } else {
            const int a0[4] = { (int) (s0.x & 0xffff), (int) (s0.x >> 16), (int) (s0.y & 0xffff), (int) (s0.y >> 16) };
            const int a1[4] = { (int) (s1.x & 0xffff), (int) (s1.x >> 16), (int) (s1.y & 0xffff), (int) (s1.y >> 16) };
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = obmc_combine<false> (A, mode, fast, noscale, e.dc, a0[k], a1[k]);
          }
        }
        int w_y = s_wy[r];
        if (!fast) {
          if (y < yoff) w_y += s_wy[2 * yoff - r - 1];
          if (y >= A.nby * ybsep - yoff) w_y += s_wy[2 * (yblen - yoff) - r - 1];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int a = 4 * g + k, x = xg + k;
          if (a < xblen && x >= tx0 && x <= x1) {
            int w_x = s_wx[a];
            if (!fast) {
              if (x < xoff) w_x += s_wx[2 * xoff - a - 1];
              if (x >= A.nbx * xbsep - xoff) w_x += s_wx[2 * (xblen - xoff) - a - 1];
            }
            atomicAdd (&acc[y - ty0][x - tx0], v[k] * w_x * w_y);
          }
        }
      }
      bi_ += step_i; bj += step_j;
      if (bi_ >= tni) { bi_ -= tni; bj++; }
    }
  }
  __syncthreads ();

  // ---- finish: four pixels per thread ---------------------------------------------------
  for (int t = threadIdx.x; t < (O4_W / 4) * th; t += blockDim.x) {
    const int ry = t / (O4_W / 4), lx = (t - ry * (O4_W / 4)) * 4;
    const int x = tx0 + lx, y = ty0 + ry;
    if (x >= width) continue;
    const int npx = min (4, width - x);
    const int4 av = *reinterpret_cast<const int4 *> (&acc[ry][lx]);
    const int sum[4] = { av.x, av.y, av.z, av.w };
    const size_t ro = (size_t) y * A.res.stride[comp];
    if (A.add) {
      int r[4];
      const char *rrow = plane_ptr (A.res, pic, comp) + ro;
      if (npx == 4) {
        if (A.res_is_s32) {
          const int4 q = *reinterpret_cast<const int4 *> (rrow + (size_t) x * 4);
          r[0] = w16 (q.x); r[1] = w16 (q.y); r[2] = w16 (q.z); r[3] = w16 (q.w);
        } else {
          const int2 q = *reinterpret_cast<const int2 *> (rrow + (size_t) x * 2);
          r[0] = (q.x << 16) >> 16; r[1] = q.x >> 16; r[2] = (q.y << 16) >> 16; r[3] = q.y >> 16;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          r[k] = k < npx ? (A.res_is_s32 ? w16 (reinterpret_cast<const int *> (rrow)[x + k])
                                         : (int) reinterpret_cast<const short *> (rrow)[x + k]) : 0;
      }
      unsigned packed = 0;
      int a16[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a16[k] = w16 (sum[k]);
        int tt = w16 (a16[k] + 32) >> 6;
        tt = w16 (r[k] + tt);
        packed |= (unsigned) clampi (tt, 0, 255) << (8 * k);
      }
      uint8_t *orow = reinterpret_cast<uint8_t *> (plane_ptr (A.out, pic, comp)) + (size_t) y * A.out.stride[comp] + x;
      if (npx == 4) *reinterpret_cast<unsigned *> (orow) = packed;
      else for (int k = 0; k < npx; k++) orow[k] = (uint8_t) (packed >> (8 * k));
      if (A.has_acc) {
        short *arow = reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp]) + x;
        for (int k = 0; k < npx; k++) arow[k] = (short) a16[k];
      }
    } else {
      short *rrow = reinterpret_cast<short *> (plane_ptr (A.res, pic, comp) + ro) + x;
      short *arow = A.has_acc ? reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp]) + x : nullptr;
      for (int k = 0; k < npx; k++) {
        const int tt = w16 (w16 (sum[k]) - 8160) >> 6;
        rrow[k] = (short) w16 (rrow[k] - tt);
        if (arow) arow[k] = (short) tt;
      }
    }
  }
}

// ---- the reference's per-pixel renderer (global motion) ----------------------------------------
// schro_motion_render_ref (schroedinger/schromotionref.c:245-330), which the reference switches to whenever
// params->have_global_motion is set (schroedinger/schromotion.c:113-121): a pixel is the rounded sum of the
// (at most four) blocks covering it, each fetched pixel by pixel with clamped coordinates
// (schro_upsampled_frame_get_pixel_precN, schroedinger/schroframe.c:2033-2046, 2124-2143, 2209-2265); blocks
// flagged using_global take their vector from the picture's global-motion model at that pixel (:22-41).  The
// prediction is clamped to 0..255 before the residual is added.  One thread per pixel.
struct RefRenderArgs {
  ObmcArgs o;
  int gm[2][10];                      // b0 b1 a_exp a00 a01 a10 a11 c_exp c0 c1 per reference
};

__device__ __forceinline__ int ref_ramp (int x, int offset)
{
  if (offset == 1) return x == 0 ? 3 : 5;
  return 1 + (6 * x + offset - 1) / (2 * offset - 1);
}

__device__ __forceinline__ int ref_pixel_prec1 (const uint8_t *ref, int rstride, int w, int h, int x, int y)
{
  x = clampi (x, 0, w * 2 - 2);
  y = clampi (y, 0, h * 2 - 2);
  return halfpel (ref, rstride, x, y, 0, 0);
}

__device__ __forceinline__ int ref_pixel (const uint8_t *ref, int rstride, int w, int h, int x, int y, int prec)
{
  if (prec == 0) return __ldg (ref + (ptrdiff_t) clampi (y, 0, h - 1) * rstride + clampi (x, 0, w - 1));
  if (prec == 1) return ref_pixel_prec1 (ref, rstride, w, h, x, y);
  if (prec == 2) { x <<= 1; y <<= 1; }
  const int hx = x >> 2, hy = y >> 2, rx = x & 3, ry = y & 3;
  const int v = (4 - ry) * (4 - rx) * ref_pixel_prec1 (ref, rstride, w, h, hx, hy)
      + (4 - ry) * rx * ref_pixel_prec1 (ref, rstride, w, h, hx + 1, hy)
      + ry * (4 - rx) * ref_pixel_prec1 (ref, rstride, w, h, hx, hy + 1)
      + ry * rx * ref_pixel_prec1 (ref, rstride, w, h, hx + 1, hy + 1);
  return (v + 8) >> 4;
}

__global__ void __launch_bounds__ (256)
obmc_ref_kernel (const RefRenderArgs R)
{
  const ObmcArgs &A = R.o;
  const int comp = blockIdx.z % A.ncomp, pic = blockIdx.z / A.ncomp;
  const int width = A.w[comp], height = A.h[comp];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) / 2, yoff = (yblen - ybsep) / 2;
  const int W = xbsep * A.nbx, H = ybsep * A.nby;
  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;
  const int i0 = (x + xoff) / xbsep - 1, j0 = (y + yoff) / ybsep - 1;
  int value = 0;
#pragma unroll
  for (int dj = 0; dj < 2; dj++)
#pragma unroll
    for (int di = 0; di < 2; di++) {
      const int i = i0 + di, j = j0 + dj;
      if (i < 0 || j < 0 || i >= A.nbx || j >= A.nby) continue;
      const int xmin = i * xbsep - xoff, ymin = j * ybsep - yoff;
      const int xmax = (i + 1) * xbsep + xoff, ymax = (j + 1) * ybsep + yoff;
      if (x < xmin || y < ymin || x >= xmax || y >= ymax) continue;
      int wx = 8, wy = 8;
      if (!(xoff == 0 || x < xoff || x >= W - xoff)) {
        if (x - xmin < 2 * xoff) wx = ref_ramp (x - xmin, xoff);
        else if (xmax - 1 - x < 2 * xoff) wx = ref_ramp (xmax - 1 - x, xoff);
      }
      if (!(yoff == 0 || y < yoff || y >= H - yoff)) {
        if (y - ymin < 2 * yoff) wy = ref_ramp (y - ymin, yoff);
        else if (ymax - 1 - y < 2 * yoff) wy = ref_ramp (ymax - 1 - y, yoff);
      }
      const MotionVector *mv = mvs + (size_t) j * A.nbx + i;
      const unsigned flags = mv->flags;
      const int mode = flags & 3;
      int v;
      if (mode == 0) {
        v = (int) mv->v[comp] + 128;
      } else {
        int s[2] = { 0, 0 };
#pragma unroll
        for (int r = 0; r < 2; r++) {
          if (!((mode >> r) & 1)) continue;
          int dx, dy;
          if ((flags >> 2) & 1) {
            const int *g = R.gm[r];
            const int alpha = g[2], beta = g[7];
            const int scale = (1 << beta) - (g[8] * x + g[9] * y);
            dx = (scale * (g[3] * x + g[4] * y + (1 << alpha) * g[0])) >> (alpha + beta);
            dy = (scale * (g[5] * x + g[6] * y + (1 << alpha) * g[1])) >> (alpha + beta);
          } else {
            dx = mv->v[r]; dy = mv->v[2 + r];
          }
          dx >>= A.hs[comp]; dy >>= A.vs[comp];
          s[r] = ref_pixel (r ? ref1 : ref0, r ? rs1 : rs0, width, height, (x << A.prec) + dx, (y << A.prec) + dy, A.prec);
        }
        v = mode == 3 ? A.w1 * s[0] + A.w2 * s[1] : (A.w1 + A.w2) * s[mode == 1 ? 0 : 1];
        v = (v + (1 << (A.bits - 1))) >> A.bits;
      }
      value += v * wx * wy;
    }
  const int line = clampi ((value + 32) >> 6, 0, 255) - 128;
  if (A.has_acc)
    reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp])[x] = (short) line;
  short *rrow = reinterpret_cast<short *> (plane_ptr (A.res, pic, comp) + (size_t) y * A.res.stride[comp]);
  if (A.add)
    reinterpret_cast<uint8_t *> (plane_ptr (A.out, pic, comp))[(size_t) y * A.out.stride[comp] + x] =
        (uint8_t) clampi ((int) rrow[x] + line + 128, 0, 255);
  else
    rrow[x] = (short) ((int) rrow[x] - line);
}

// schroedinger/schromotion.c:40-79
static int get_ramp (int x, int offset)
{
  if (offset == 1) return x == 0 ? 3 : 5;
  return 1 + (6 * x + offset - 1) / (2 * offset - 1);
}

static void obmc_weights (unsigned char *w, int len, int off)
{
  for (int i = 0; i < len; i++) {
    int v;
    if (off == 0) v = 8;
    else if (i < 2 * off) v = get_ramp (i, off);
    else if (len - 1 - i < 2 * off) v = get_ramp (len - 1 - i, off);
    else v = 8;
    w[i] = (unsigned char) v;
  }
}

}  // namespace sb2

using namespace sb2;

// 0: pick by geometry; 1 / 2 / 3: force the TMA block / scatter / per-pixel kernel (tests run every
// case through all three); the environment variable SB2_OBMC_KERNEL sets the initial value
static int g_obmc_variant = -1;
static thread_local int g_obmc_last = 0;
extern "C" void sb2_obmc_force_kernel (int which) { g_obmc_variant = which < 0 || which > 4 ? 0 : which; }
extern "C" int sb2_obmc_last_kernel (void) { return g_obmc_last; }
static int forced_variant ()
{
  if (g_obmc_variant < 0) {
    const char *v = getenv ("SB2_OBMC_KERNEL");
    g_obmc_variant = v ? atoi (v) : 0;
    if (g_obmc_variant < 0 || g_obmc_variant > 4) g_obmc_variant = 0;
  }
  return g_obmc_variant;
}

extern "C" int
sb2_obmc_render (const sb2_obmc_params *p, const void *motion_vectors, size_t mv_picture_pitch,
    const sb2_slab *ref0, const sb2_slab *ref1, const sb2_slab *acc, const sb2_slab *residual,
    int residual_is_s32, int add, const sb2_slab *out, void *stream)
{
  if (!p || !motion_vectors || !ref0 || !residual)
    return set_error (SB2_ERR_ARG, "sb2_obmc_render: null argument");
  if (add && !out) return set_error (SB2_ERR_ARG, "sb2_obmc_render: add needs an output slab");
  if (!add && residual_is_s32)
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_obmc_render: the subtract direction is s16 only (as the reference)");
  if (p->mv_precision < 0 || p->mv_precision > 3)
    return set_error (SB2_ERR_ARG, "sb2_obmc_render: mv_precision %d", p->mv_precision);
  if (p->xblen < p->xbsep || p->yblen < p->ybsep || p->xblen > 64 || p->yblen > 64 || p->xbsep < 1 || p->ybsep < 1)
    return set_error (SB2_ERR_ARG, "sb2_obmc_render: bad block geometry %dx%d sep %dx%d", p->xblen, p->yblen, p->xbsep, p->ybsep);
  const int ncomp = residual->ncomp, count = residual->count;
  if (ncomp < 1 || ncomp > SB2_MAX_COMPONENTS || ref0->ncomp != ncomp || ref0->count != count ||
      (ref1 && (ref1->ncomp != ncomp || ref1->count != count)) || (out && (out->ncomp != ncomp || out->count != count)) ||
      (acc && (acc->ncomp != ncomp || acc->count != count)))
    return set_error (SB2_ERR_ARG, "sb2_obmc_render: slab shapes differ");

  ObmcArgs A;
  A.ref0 = planeset_from_slab (ref0);
  A.ref1 = planeset_from_slab (ref1 ? ref1 : ref0);
  A.acc = planeset_from_slab (acc ? acc : residual);
  A.res = planeset_from_slab (residual);
  A.out = planeset_from_slab (out ? out : residual);
  A.mvs = static_cast<const MotionVector *> (motion_vectors);
  A.mv_pitch = mv_picture_pitch;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.prec = p->mv_precision;
  A.w1 = p->picture_weight_1;
  A.w2 = p->picture_weight_2;
  A.bits = p->picture_weight_bits;
  A.ncomp = ncomp;
  A.add = add;
  A.res_is_s32 = residual_is_s32;
  A.has_ref1 = ref1 != nullptr;
  A.has_acc = acc != nullptr;
  int maxw = 0, maxh = 0;
  double bytes = 0;
  // The rendered area is that of `dest` (schromotion8.c:722-751: motion->width = comp->width of
  // dest), i.e. of `acc`; callers pass a picture-size dest with an iwt-padded (taller) addframe
  // (schrodecoder.c:1784, schroencoder.c:2447).  Without an acc slab the output (or, in the
  // subtract direction, the residual) gives the size.  No plane may be smaller than that area.
  const sb2_slab *area = acc ? acc : (add ? out : residual);
  for (int c = 0; c < ncomp; c++) {
    const sb2_slab *all[3] = { residual, out, acc };
    for (const sb2_slab *s : all)
      if (s && (s->width[c] < area->width[c] || s->height[c] < area->height[c]))
        return set_error (SB2_ERR_ARG, "sb2_obmc_render: component %d: a %dx%d plane is smaller than the %dx%d render area",
            c, s->width[c], s->height[c], area->width[c], area->height[c]);
  }
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    const int hs = c ? p->chroma_h_shift : 0, vs = c ? p->chroma_v_shift : 0;
    A.w[c] = c < ncomp ? area->width[c] : 0;
    A.h[c] = c < ncomp ? area->height[c] : 0;
    A.hs[c] = hs;
    A.vs[c] = vs;
    A.xbsep[c] = p->xbsep >> hs;
    A.ybsep[c] = p->ybsep >> vs;
    A.xblen[c] = p->xblen >> hs;
    A.yblen[c] = p->yblen >> vs;
    if (c < ncomp) {
      if (A.xbsep[c] < 1 || A.ybsep[c] < 1)
        return set_error (SB2_ERR_ARG, "sb2_obmc_render: block separation vanishes in component %d", c);
      if (A.nbx * A.xbsep[c] < A.w[c] || A.nby * A.ybsep[c] < A.h[c])
        return set_error (SB2_ERR_ARG, "sb2_obmc_render: %dx%d blocks do not cover component %d", A.nbx, A.nby, c);
      obmc_weights (A.wx[c], A.xblen[c], (A.xblen[c] - A.xbsep[c]) / 2);
      obmc_weights (A.wy[c], A.yblen[c], (A.yblen[c] - A.ybsep[c]) / 2);
      maxw = max (maxw, A.w[c]);
      maxh = max (maxh, A.h[c]);
      // algorithmic bytes: 4 phases of each reference + residual + output
      const double px = (double) A.w[c] * A.h[c] * count;
      bytes += px * (4.0 * (ref1 ? 2 : 1) + (residual_is_s32 ? 4 : 2) + (add ? 1 : 4));
    }
  }
  bytes += (double) A.nbx * A.nby * 20 * count;
  dim3 grid (ceil_div (maxw, 32), ceil_div (maxh, 8), ncomp * count);
  {
    LaunchScope scope (add ? "obmc_render_add" : "obmc_render_sub", bytes, as_stream (stream));
    const int force = forced_variant ();
    if (force == 4) {
      const int rc = obmc_blocks_launch (A, ref0, ref1, count, false, as_stream (stream));
      if (rc == SB2_OK) {
        g_obmc_last = 4;
        return check_cuda (cudaGetLastError (), "obmc_kernel_blocks launch");
      }
      return set_error (SB2_ERR_UNSUPPORTED, "sb2_obmc_render: the block kernel does not cover this geometry");
    }
    // 2: the scatter kernel needs 4-byte aligned output rows, 16-byte aligned residual rows, reference
    // planes whose rows are 4-byte aligned (word loads) and a block table that fits
    bool v4_ok = true;
    for (int c = 0; c < ncomp && v4_ok; c++) {
      if (out && ((out->stride[c] | out->offset[c]) & 3)) v4_ok = false;
      if ((residual->stride[c] | residual->offset[c]) & 15) v4_ok = false;
      if (((A.xblen[c] + 3) >> 2) * A.yblen[c] > 256) v4_ok = false;     // one block's items must fit a CTA pass
      const int ni = (O4_W + A.xblen[c]) / A.xbsep[c] + 2, nj = (O4_H + A.yblen[c]) / A.ybsep[c] + 2;
      if (ni * nj > MAX_ENT) v4_ok = false;
      if ((ref0->stride[c] & 3) || (ref1 && (ref1->stride[c] & 3))) v4_ok = false;
    }
    if (out && (((size_t) out->base | out->picture_pitch) & 3)) v4_ok = false;
    if (((size_t) residual->base | residual->picture_pitch) & 15) v4_ok = false;
    // 1: the block-per-warp kernel on TMA-staged reference regions (obmc_blocks.cu): blocks of at most
    // 32 lanes (8-pixel items x rows), 32-pixel borders.  Measured at 2160p (32 pictures, +-16 pixel
    // vectors): 6.2 ms against 5.6 ms for the scatter kernel, so it is the choice only
    // where the scatter kernel does not apply -- or when forced
    if (force == 1 || (force == 0 && !v4_ok)) {
      const int rc = obmc_blocks_launch (A, ref0, ref1, count, true, as_stream (stream));
      if (rc == SB2_OK) {
        g_obmc_last = 1;
        return check_cuda (cudaGetLastError (), "obmc_kernel_blocks launch");
      }
      if (force == 1) return set_error (SB2_ERR_UNSUPPORTED, "sb2_obmc_render: the TMA kernel does not cover this geometry");
    }
    const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
    if (force == 2 && !v4_ok) return set_error (SB2_ERR_UNSUPPORTED, "sb2_obmc_render: the scatter kernel does not cover this geometry");
    if (v4_ok && force != 3) {
      g_obmc_last = 2;
      dim3 g4 (ceil_div (maxw, O4_W), ceil_div (maxh, O4_H), ncomp * count);
      if (simple) obmc_kernel_v4<true><<<g4, 256, 0, as_stream (stream)>>> (A);
      else obmc_kernel_v4<false><<<g4, 256, 0, as_stream (stream)>>> (A);
    } else {
      // 3: one thread per pixel, any geometry
      g_obmc_last = 3;
      obmc_kernel<<<grid, 256, 0, as_stream (stream)>>> (A);
    }
  }
  return check_cuda (cudaGetLastError (), "obmc_kernel launch");
}

extern "C" int
sb2_obmc_render_ref (const sb2_obmc_params *p, const int *global_motion, const void *motion_vectors,
    size_t mv_picture_pitch, const sb2_slab *ref0, const sb2_slab *ref1, const sb2_slab *acc,
    const sb2_slab *residual, int add, const sb2_slab *out, void *stream)
{
  if (!p || !motion_vectors || !ref0 || !residual) return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: null argument");
  if (add && !out) return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: add needs an output slab");
  if (p->mv_precision < 0 || p->mv_precision > 3 || p->picture_weight_bits < 1)
    return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: mv_precision %d / weight bits %d", p->mv_precision, p->picture_weight_bits);
  if (p->xblen < p->xbsep || p->yblen < p->ybsep || p->xbsep < 1 || p->ybsep < 1 || p->xblen > 2 * p->xbsep || p->yblen > 2 * p->ybsep)
    return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: bad block geometry %dx%d sep %dx%d", p->xblen, p->yblen, p->xbsep, p->ybsep);
  const int ncomp = residual->ncomp, count = residual->count;
  if (ncomp < 1 || ncomp > SB2_MAX_COMPONENTS || ref0->ncomp != ncomp || ref0->count != count ||
      (ref1 && (ref1->ncomp != ncomp || ref1->count != count)) || (out && (out->ncomp != ncomp || out->count != count)) ||
      (acc && (acc->ncomp != ncomp || acc->count != count)))
    return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: slab shapes differ");
  RefRenderArgs R;
  ObmcArgs &A = R.o;
  A.ref0 = planeset_from_slab (ref0);
  A.ref1 = planeset_from_slab (ref1 ? ref1 : ref0);
  A.acc = planeset_from_slab (acc ? acc : residual);
  A.res = planeset_from_slab (residual);
  A.out = planeset_from_slab (out ? out : residual);
  A.mvs = static_cast<const MotionVector *> (motion_vectors);
  A.mv_pitch = mv_picture_pitch;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.prec = p->mv_precision;
  A.w1 = p->picture_weight_1;
  A.w2 = p->picture_weight_2;
  A.bits = p->picture_weight_bits;
  A.ncomp = ncomp;
  A.add = add;
  A.res_is_s32 = 0;
  A.has_ref1 = ref1 != nullptr;
  A.has_acc = acc != nullptr;
  for (int r = 0; r < 2; r++)
    for (int k = 0; k < 10; k++) R.gm[r][k] = global_motion ? global_motion[10 * r + k] : 0;
  // the rendered area is that of dest (schromotionref.c:267: comp = dest->components + k)
  const sb2_slab *area = acc ? acc : (add ? out : residual);
  int maxw = 0, maxh = 0;
  double bytes = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    A.w[c] = A.h[c] = A.xbsep[c] = A.ybsep[c] = A.xblen[c] = A.yblen[c] = A.hs[c] = A.vs[c] = 0;
    if (c >= ncomp) continue;
    const int hs = c ? p->chroma_h_shift : 0, vs = c ? p->chroma_v_shift : 0;
    A.w[c] = area->width[c];
    A.h[c] = area->height[c];
    const sb2_slab *all[4] = { residual, out, acc, ref0 };
    for (const sb2_slab *s : all)
      if (s && (s->width[c] < A.w[c] || s->height[c] < A.h[c]))
        return set_error (SB2_ERR_ARG, "sb2_obmc_render_ref: component %d: a plane is smaller than the rendered area", c);
    A.xbsep[c] = p->xbsep >> hs; A.ybsep[c] = p->ybsep >> vs;
    A.xblen[c] = p->xblen >> hs; A.yblen[c] = p->yblen >> vs;
    A.hs[c] = hs; A.vs[c] = vs;
    maxw = max (maxw, A.w[c]);
    maxh = max (maxh, A.h[c]);
    bytes += (double) A.w[c] * A.h[c] * count * (4.0 * (ref1 ? 2 : 1) + 2 + (add ? 1 : 2));
  }
  dim3 grid (ceil_div (maxw, 32), ceil_div (maxh, 8), ncomp * count);
  {
    LaunchScope scope (add ? "obmc_render_ref_add" : "obmc_render_ref_sub", bytes, as_stream (stream));
    obmc_ref_kernel<<<grid, 256, 0, as_stream (stream)>>> (R);
  }
  return check_cuda (cudaGetLastError (), "obmc_ref_kernel launch");
}
