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

// ---- v2: block table in shared memory + unified 4-tap fetch -------------------------
// Every sub-pel case of schroframe.c:2288-2413 is the same 4-tap sum
//   (w00*s00 + w01*s01 + w10*s10 + w11*s11 + 8) >> 4,  weights summing to 16:
// the copy case is w00 = 16, the two avgub cases are 8/8 ((8a+8b+8)>>4 == (a+b+1)>>1),
// prec 0/1 are single taps.  So the per-block work (vector decode, clamp, phase
// selection, weights) is done once per CTA into a table and every pixel issues up to
// 2x2 blocks x 2 refs x 4 taps of independent loads before any arithmetic.
struct BlkRef { int o[4]; unsigned w; };
struct BlkEnt { BlkRef r[2]; short mode, fast, dc, pad; };
constexpr int OT_W = 32, OT_H = 8;            // pixel tile
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

__device__ __forceinline__ int fetch4 (const uint8_t *ref, const BlkRef &br, int pix)
{
  // taps with zero weight are not loaded
  const unsigned w = br.w;
  int acc = 8;
  acc += (int) (w & 0xff) * (int) __ldg (ref + br.o[0] + pix);
  if (w & 0x0000ff00u) acc += (int) ((w >> 8) & 0xff) * (int) __ldg (ref + br.o[1] + pix);
  if (w & 0x00ff0000u) acc += (int) ((w >> 16) & 0xff) * (int) __ldg (ref + br.o[2] + pix);
  if (w & 0xff000000u) acc += (int) (w >> 24) * (int) __ldg (ref + br.o[3] + pix);
  return acc >> 4;
}

__global__ void __launch_bounds__ (256)
obmc_kernel_v2 (const ObmcArgs A)
{
  __shared__ BlkEnt tab[MAX_ENT];
  // OBMC window and per-column / per-row covering-block ranges: per-lane indexed, so they
  // must not live in the (uniform-access) constant bank of the kernel arguments
  __shared__ unsigned char s_wx[64], s_wy[64];
  __shared__ short s_i0[OT_W], s_i1[OT_W], s_j0[OT_H], s_j1[OT_H];
  const int comp = blockIdx.z % A.ncomp, pic = blockIdx.z / A.ncomp;
  const int width = A.w[comp], height = A.h[comp];
  const int tx0 = blockIdx.x * OT_W, ty0 = blockIdx.y * OT_H;
  if (tx0 >= width || ty0 >= height) return;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, (width - xoff) / xbsep);
  const int max_y_blocks = min (A.nby - 1, (height - yoff) / ybsep);
  const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));
  const unsigned char *wx = s_wx, *wy = s_wy;
  if (threadIdx.x < 64) {
    s_wx[threadIdx.x] = A.wx[comp][threadIdx.x];
    s_wy[threadIdx.x] = A.wy[comp][threadIdx.x];
  } else if (threadIdx.x < 64 + OT_W) {
    const int x = tx0 + (int) threadIdx.x - 64;
    s_i0[threadIdx.x - 64] = (short) ((x + xoff - xblen + 1 > 0) ? (x + xoff - xblen + xbsep) / xbsep : 0);
    s_i1[threadIdx.x - 64] = (short) min (A.nbx - 1, (x + xoff) / xbsep);
  } else if (threadIdx.x < 64 + OT_W + OT_H) {
    const int y = ty0 + (int) threadIdx.x - 64 - OT_W;
    s_j0[threadIdx.x - 64 - OT_W] = (short) ((y + yoff - yblen + 1 > 0) ? (y + yoff - yblen + ybsep) / ybsep : 0);
    s_j1[threadIdx.x - 64 - OT_W] = (short) min (A.nby - 1, (y + yoff) / ybsep);
  }

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  // blocks overlapping this tile
  const int x1 = min (tx0 + OT_W, width) - 1, y1 = min (ty0 + OT_H, height) - 1;
  const int ti0 = (tx0 + xoff - xblen + 1 > 0) ? (tx0 + xoff - xblen + xbsep) / xbsep : 0;
  const int ti1 = min (A.nbx - 1, (x1 + xoff) / xbsep);
  const int tj0 = (ty0 + yoff - yblen + 1 > 0) ? (ty0 + yoff - yblen + ybsep) / ybsep : 0;
  const int tj1 = min (A.nby - 1, (y1 + yoff) / ybsep);
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

  const int x = tx0 + (threadIdx.x & 31), y = ty0 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;

  const int j0 = s_j0[threadIdx.x >> 5], j1 = s_j1[threadIdx.x >> 5];
  const int i0 = s_i0[threadIdx.x & 31], i1 = s_i1[threadIdx.x & 31];

  int sum = 0;
  for (int j = j0; j <= j1; j++) {
    const int b = y - (ybsep * j - yoff);
    for (int i = i0; i <= i1; i++) {
      const int a = x - (xbsep * i - xoff);
      const BlkEnt &e = tab[(j - tj0) * tni + (i - ti0)];
      const int mode = e.mode;
      const bool fast = e.fast != 0;
      int s0 = 0, s1 = 0;
      if (mode & 1) s0 = fetch4 (ref0, e.r[0], b * rs0 + a);
      if (mode & 2) s1 = fetch4 (ref1, e.r[1], b * rs1 + a);
      int v;
      if (mode == 0) {
        const int dc = (int) e.dc + 128;
        v = fast ? w16 (dc) : (dc & 0xff);
      } else if (mode == 3) {
        if (simple) {
          v = (s0 + s1 + 1) >> 1;
        } else if (fast) {
          int t = w16 (s0 * w16 (A.w1 << (6 - A.bits)));
          const int u = w16 (s1 * w16 (A.w2 << (6 - A.bits)));
          t = w16 (t + u);
          t = w16 (t + 32);
          v = t >> 6;
        } else {
          int t = w16 (s0 * w16 (A.w1));
          const int u = w16 (s1 * w16 (A.w2));
          t = w16 (t + u);
          t = w16 (t + ((1 << A.bits) >> 1));
          v = clampi (t >> A.bits, 0, 255);
        }
      } else {
        const int s = (mode == 1) ? s0 : s1;
        if (fast) {
          if (simple) v = s;
          else {
            int t = w16 (s * w16 ((A.w1 + A.w2) << (6 - A.bits)));
            t = w16 (t + 32);
            v = t >> 6;
          }
        } else {
          if (noscale) v = s;
          else v = ((s * (A.w1 + A.w2) + (1 << (A.bits - 1))) >> A.bits) & 0xff;
        }
      }
      int w_x = wx[a], w_y = wy[b];
      if (!fast) {
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
    short *r = reinterpret_cast<short *> (plane_ptr (A.res, pic, comp) + (size_t) y * A.res.stride[comp]) + x;
    const int t = w16 (a16 - 8160) >> 6;
    *r = (short) w16 (*r - t);
    if (A.has_acc)
      reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp])[x] = (short) t;
  }
}

// ---- v3: four horizontally adjacent pixels per thread ------------------------------
// Same block table as v2; a thread owns pixels x..x+3 of one row and loops over the UNION
// of the blocks covering them (2x2 for xblen <= 2*xbsep), with a zero window weight where a
// pixel lies outside a block.  Table reads, mode logic and loop control are paid once per
// four pixels, residual loads and output stores are 128 / 32 bits wide.
constexpr int O3_W = 128, O3_H = 8;

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

template <bool SIMPLE>
__global__ void __launch_bounds__ (256)
obmc_kernel_v3 (const ObmcArgs A)
{
  __shared__ __align__ (16) BlkEnt tab[MAX_ENT];
  __shared__ unsigned char s_wx[64], s_wy[64];
  __shared__ short s_i0[O3_W], s_i1[O3_W], s_j0[O3_H], s_j1[O3_H];
  const int comp = blockIdx.z % A.ncomp, pic = blockIdx.z / A.ncomp;
  const int width = A.w[comp], height = A.h[comp];
  const int tx0 = blockIdx.x * O3_W, ty0 = blockIdx.y * O3_H;
  if (tx0 >= width || ty0 >= height) return;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, (width - xoff) / xbsep);
  const int max_y_blocks = min (A.nby - 1, (height - yoff) / ybsep);
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));

  if (threadIdx.x < 64) {
    s_wx[threadIdx.x] = A.wx[comp][threadIdx.x];
    s_wy[threadIdx.x] = A.wy[comp][threadIdx.x];
  }
  if (threadIdx.x < O3_W) {
    const int x = min (tx0 + (int) threadIdx.x, width - 1);
    s_i0[threadIdx.x] = (short) ((x + xoff - xblen + 1 > 0) ? (x + xoff - xblen + xbsep) / xbsep : 0);
    s_i1[threadIdx.x] = (short) min (A.nbx - 1, (x + xoff) / xbsep);
  } else if (threadIdx.x < O3_W + O3_H) {
    const int y = min (ty0 + (int) threadIdx.x - O3_W, height - 1);
    s_j0[threadIdx.x - O3_W] = (short) ((y + yoff - yblen + 1 > 0) ? (y + yoff - yblen + ybsep) / ybsep : 0);
    s_j1[threadIdx.x - O3_W] = (short) min (A.nby - 1, (y + yoff) / ybsep);
  }

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  const int x1 = min (tx0 + O3_W, width) - 1, y1 = min (ty0 + O3_H, height) - 1;
  const int ti0 = (tx0 + xoff - xblen + 1 > 0) ? (tx0 + xoff - xblen + xbsep) / xbsep : 0;
  const int ti1 = min (A.nbx - 1, (x1 + xoff) / xbsep);
  const int tj0 = (ty0 + yoff - yblen + 1 > 0) ? (ty0 + yoff - yblen + ybsep) / ybsep : 0;
  const int tj1 = min (A.nby - 1, (y1 + yoff) / ybsep);
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

  const int lx = (threadIdx.x & 31) * 4;
  const int x = tx0 + lx, y = ty0 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  const int npx = min (4, width - x);

  const int j0 = s_j0[threadIdx.x >> 5], j1 = s_j1[threadIdx.x >> 5];
  const int i0 = s_i0[lx], i1 = s_i1[lx + 3];

  int sum[4] = { 0, 0, 0, 0 };
  for (int j = j0; j <= j1; j++) {
    const int b = y - (ybsep * j - yoff);
    int wy_plain = s_wy[b], wy_fold = wy_plain;
    if (y < yoff) wy_fold += s_wy[2 * yoff - b - 1];
    if (y >= A.nby * ybsep - yoff) wy_fold += s_wy[2 * (yblen - yoff) - b - 1];
    for (int i = i0; i <= i1; i++) {
      const BlkEnt &e = tab[(j - tj0) * tni + (i - ti0)];
      const int mode = e.mode;
      const bool fast = e.fast != 0;
      const int dc = e.dc;
      const int bx = xbsep * i - xoff;
      const int w_y = fast ? wy_plain : wy_fold;
      const BlkRef r0 = e.r[0], r1 = e.r[1];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int a = x + k - bx;
        const bool in = a >= 0 && a < xblen && k < npx;
        const int ac = min (max (a, 0), xblen - 1);
        int w_x = s_wx[ac];
        if (!fast) {
          if (x + k < xoff) w_x += s_wx[max (2 * xoff - ac - 1, 0)];
          if (x + k >= A.nbx * xbsep - xoff) w_x += s_wx[min (max (2 * (xblen - xoff) - ac - 1, 0), 63)];
        }
        int s0 = 0, s1 = 0;
        if (in && (mode & 1)) s0 = fetch4 (ref0, r0, b * rs0 + ac);
        if (in && (mode & 2)) s1 = fetch4 (ref1, r1, b * rs1 + ac);
        const int v = obmc_combine<SIMPLE> (A, mode, fast, noscale, dc, s0, s1);
        sum[k] += in ? v * w_x * w_y : 0;
      }
    }
  }

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
      int t = w16 (a16[k] + 32) >> 6;
      t = w16 (r[k] + t);
      packed |= (unsigned) clampi (t, 0, 255) << (8 * k);
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
      const int t = w16 (w16 (sum[k]) - 8160) >> 6;
      rrow[k] = (short) w16 (rrow[k] - t);
      if (arow) arow[k] = (short) t;
    }
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
    // the table kernel needs every tile's block list to fit its shared-memory table
    bool table_ok = true;
    for (int c = 0; c < ncomp; c++) {
      const int ni = (OT_W + A.xblen[c]) / A.xbsep[c] + 2, nj = (OT_H + A.yblen[c]) / A.ybsep[c] + 2;
      if (ni * nj > MAX_ENT) table_ok = false;
    }
    // the 4-pixel kernel needs 4-byte aligned output rows and 16-byte aligned residual rows
    bool v3_ok = table_ok;
    for (int c = 0; c < ncomp && v3_ok; c++) {
      const int ni = (O3_W + A.xblen[c]) / A.xbsep[c] + 2, nj = (O3_H + A.yblen[c]) / A.ybsep[c] + 2;
      if (ni * nj > MAX_ENT) v3_ok = false;
      if (out && ((out->stride[c] | out->offset[c]) & 3)) v3_ok = false;
      if ((residual->stride[c] | residual->offset[c]) & 15) v3_ok = false;
    }
    if (out && (((size_t) out->base | out->picture_pitch) & 3)) v3_ok = false;
    if (((size_t) residual->base | residual->picture_pitch) & 15) v3_ok = false;
    const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
    // the scatter kernel additionally needs non-overlapping same-colour blocks, reference planes
    // whose rows are 4-byte aligned (word loads) and a block table that fits
    bool v4_ok = v3_ok;
    for (int c = 0; c < ncomp && v4_ok; c++) {
      if (((A.xblen[c] + 3) >> 2) * A.yblen[c] > 256) v4_ok = false;     // one block's items must fit a CTA pass
      const int ni = (O4_W + A.xblen[c]) / A.xbsep[c] + 2, nj = (O4_H + A.yblen[c]) / A.ybsep[c] + 2;
      if (ni * nj > MAX_ENT) v4_ok = false;
      if ((ref0->stride[c] & 3) || (ref1 && (ref1->stride[c] & 3))) v4_ok = false;
    }
    if (v4_ok) {
      // (a compact grid without the out-of-plane chroma CTAs was measured 3 % slower here)
      dim3 g4 (ceil_div (maxw, O4_W), ceil_div (maxh, O4_H), ncomp * count);
      if (simple) obmc_kernel_v4<true><<<g4, 256, 0, as_stream (stream)>>> (A);
      else obmc_kernel_v4<false><<<g4, 256, 0, as_stream (stream)>>> (A);
    } else if (v3_ok) {
      dim3 g3 (ceil_div (maxw, O3_W), ceil_div (maxh, O3_H), ncomp * count);
      if (simple) obmc_kernel_v3<true><<<g3, 256, 0, as_stream (stream)>>> (A);
      else obmc_kernel_v3<false><<<g3, 256, 0, as_stream (stream)>>> (A);
    } else if (table_ok) obmc_kernel_v2<<<grid, 256, 0, as_stream (stream)>>> (A);
    else obmc_kernel<<<grid, 256, 0, as_stream (stream)>>> (A);
  }
  return check_cuda (cudaGetLastError (), "obmc_kernel launch");
}
