// obmc_blocks.cu -- OBMC renderer: TMA-staged reference regions, one block per warp pass.
//
// Same result, bit for bit, as schro_motion_render_u8 (schroedinger/schromotion8.c:700-929) with the
// sub-pel block fetch of schroedinger/schroframe.c:2288-2482.
//
// A CTA owns a 64x32 tile of one component of one picture.  One bulk tensor copy per reference
// (cp.async.bulk.tensor.4d, UTMALDG) stages the region every block of the tile can read -- 112 bytes x
// 4 half-pel phases x 72 rows -- in shared memory behind an mbarrier while the CTA decodes its blocks'
// vectors into a table.  Then the unit of work is a BLOCK: a warp takes one block (12x12: 24 lanes =
// 12 rows x two 8-pixel items; 6x6 chroma blocks: five blocks per pass), so mode, sub-pel case, tap
// offsets and weights are the same for every lane of a block and nothing is decided per pixel:
//   * the sub-pel cases of schroframe.c:2288-2413 at quarter-pel precision are a copy, the byte-wise
//     rounded average of two half-pel samples (avgub == VAVG over four pixels at once) or the exact
//     mean of four; only eighth-pel vectors need the general 4-tap sum;
//   * the two-reference average of the common weights (1,1,1) is one more byte-wise average;
//   * the OBMC window weights of a lane's pixels are constants of the lane (interior blocks), the
//     accumulator address is block base + lane constant; the accumulator tile carries a margin as
//     wide as a block, so no pixel is ever tested against the tile;
//   * contributions meet in a shared-memory accumulator with reductions (no return value), the
//     finish pass adds the residual and stores.
// A block whose window is not inside the staged region (the +-4000 outliers of a stream) reads its
// taps from global memory through the same code (generic addressing is not used: two call sites).

#include "obmc_common.cuh"
#include <cuda.h>
#include <cstddef>
#include <cstring>
#include <mutex>

namespace sb2 {

constexpr int T6_W = 64, T6_H = 32;            // output tile
constexpr int T6_THREADS = 256;
constexpr int R6_MX = 16, R6_MY = 20;          // staged margin left / top (x must be a multiple of 16: see tools/tma_probe.cu)
constexpr int R6_W = 112, R6_H = 72;           // 16 + 64 + 32 columns, 20 + 32 + 20 rows
constexpr int R6_PLANE = R6_W * R6_H;          // one copy per phase: [phase][row][x], 28 words a row (rows 8 apart share banks, no others)
constexpr int R6_BYTES = 4 * R6_PLANE;         // 32256 per reference
constexpr int R6_BIAS = 8192;                  // added to stored tap offsets (a block may start above / left of the region)
constexpr int A6_MX = 16, A6_MY = 12;          // accumulator margin: a block never leaves it
constexpr int A6_H = T6_H + 2 * A6_MY;         // 56 rows of 96 used columns
constexpr int A6_P = 99;                       // row pitch in ints, 3 mod 32: the 24 lanes of a 12x12 block hit 24 banks (see the lane map)
constexpr int T6_MAXB = 200;                   // blocks overlapping a tile (chroma 6/4: 18 x 10)
constexpr int B6 = 32;                         // frame extension the renderer requires

struct Blk6 {
  unsigned short flags;                        // mode | fast << 2 | staged << 3 (bit r: reference r's window lies inside the staged region)
  short dc;
  unsigned short accoff, pad;                  // int index of block pixel (0, 0) in the accumulator tile
  unsigned w[2];                               // tap weights of the 4-tap form, one byte each
  unsigned short o[2][4];                      // R6_BIAS + byte offset of tap t of block pixel (0, 0) in the staged region
};
static_assert (sizeof (Blk6) == 32, "two 16-byte reads per table entry");

struct Maps6 { CUtensorMap m[2][3]; };

struct Smem6 {
  alignas (128) int acc[A6_H * A6_P + 4];
  alignas (16) Blk6 tab[T6_MAXB];
  alignas (16) int go[T6_MAXB][2][4];                                   // tap offsets into the reference planes (for blocks not staged)
  unsigned char wx[64], wy[64];
  alignas (16) unsigned lane[32][12];                      // per lane of a pass: wb[8], accumulator word, region byte, need | late << 8 | on << 16, -
  alignas (8) unsigned long long bar;
  alignas (128) unsigned char ref[2][R6_BYTES + 128];      // [128 spare bytes][region]; LAST: absent when nothing is staged
};
constexpr size_t SMEM6_UNSTAGED = offsetof (Smem6, ref);

__device__ __forceinline__ unsigned s6_u32 (const void *p) { return (unsigned) __cvta_generic_to_shared (p); }

// eight bytes starting at any shared / global address, from aligned 32-bit words
__device__ __forceinline__ uint2 lds8 (unsigned addr)
{
  const unsigned al = addr & ~3u, sh = (addr & 3u) * 8u;
  unsigned w0, w1, w2;
  asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(al));
  asm volatile ("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(al));
  asm volatile ("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(al));
  return make_uint2 (__funnelshift_r (w0, w1, sh), __funnelshift_r (w1, w2, sh));
}
// (global: only the words holding the `need` bytes asked for are touched -- the last row of a slab has nothing after it)
__device__ __forceinline__ uint2 ldg8 (const uint8_t *p, int need)
{
  // two 8-byte loads: a warp's request touches a line per block row, so fewer, wider requests
  const unsigned mis = (unsigned) ((size_t) p & 7);
  const uint2 *w = reinterpret_cast<const uint2 *> (p - mis);
  const uint2 a = __ldg (w), b = (int) mis + need > 8 ? __ldg (w + 1) : make_uint2 (0u, 0u);
  const bool hi = (mis & 4) != 0;
  const unsigned w0 = hi ? a.y : a.x, w1 = hi ? b.x : a.y, w2 = hi ? b.y : b.x;
  const unsigned sh = (mis & 3) * 8;
  return make_uint2 (__funnelshift_r (w0, w1, sh), __funnelshift_r (w1, w2, sh));
}

// shared-memory reduction, predicated on `on` without a branch
__device__ __forceinline__ void red_if (unsigned on, unsigned addr, unsigned v)
{
  asm volatile ("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.shared.add.u32 [%1], %2; }" :: "r"(on), "r"(addr), "r"(v) : "memory");
}

// exact (a + b + c + d + 2) >> 2 of four packed bytes
__device__ __forceinline__ unsigned mean4 (unsigned a, unsigned b, unsigned c, unsigned d)
{
  const unsigned M = 0x00ff00ffu;
  const unsigned lo = (a & M) + (b & M) + (c & M) + (d & M) + 0x00020002u;
  const unsigned hi = ((a >> 8) & M) + ((b >> 8) & M) + ((c >> 8) & M) + ((d >> 8) & M) + 0x00020002u;
  return ((lo >> 2) & M) | (((hi >> 2) & M) << 8);
}

// the general 4-tap sum (w0 s0 + w1 s1 + w2 s2 + w3 s3 + 8) >> 4 of four packed bytes, weights summing to 16
__device__ __forceinline__ unsigned taps4 (const unsigned (&s)[4], unsigned w)
{
  unsigned lo = 0x00080008u, hi = 0x00080008u;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const unsigned wt = (w >> (8 * t)) & 0xff;
    lo += wt * __byte_perm (s[t], 0, 0x4140);
    hi += wt * __byte_perm (s[t], 0, 0x4342);
  }
  lo = (lo >> 4) & 0x00ff00ffu;
  hi = (hi >> 4) & 0x00ff00ffu;
  return __byte_perm (lo, hi, 0x6420);           // p0 p1 p2 p3
}

// eight predicted pixels of one reference: LD (offset) -> uint2 reads eight bytes of tap window data
template <typename LD>
__device__ __forceinline__ uint2 predict8 (unsigned w, const int (&o)[4], LD ld)
{
  if (w == 16u) return ld (o[0]);                                         // integer / half-pel position: a copy
  if (w == 0x00000808u || w == 0x00080008u) {                             // between two half-pel samples: avgub
    const uint2 a = ld (o[0]), b = ld (w == 0x00000808u ? o[1] : o[2]);
    return make_uint2 (__vavgu4 (a.x, b.x), __vavgu4 (a.y, b.y));
  }
  const uint2 a = ld (o[0]), b = ld (o[1]), c = ld (o[2]), d = ld (o[3]);
  if (w == 0x04040404u) return make_uint2 (mean4 (a.x, b.x, c.x, d.x), mean4 (a.y, b.y, c.y, d.y));
  const unsigned sx[4] = { a.x, b.x, c.x, d.x }, sy[4] = { a.y, b.y, c.y, d.y };
  return make_uint2 (taps4 (sx, w), taps4 (sy, w));
}

// STAGED = false: the same kernel without the staged regions -- every block reads its taps from global
// memory (the path the staged kernel keeps for outlier vectors).  A quarter of the shared memory and
// half the registers: four CTAs an SM instead of two.
template <bool SIMPLE, bool STAGED>
__global__ void __launch_bounds__ (T6_THREADS, STAGED ? 2 : 4)
obmc_kernel_blocks (const ObmcArgs A, const TileGrid tiles, const __grid_constant__ Maps6 maps)
{
  extern __shared__ unsigned char smem_raw[];
  Smem6 &S = *reinterpret_cast<Smem6 *> ((reinterpret_cast<size_t> (smem_raw) + 127) & ~(size_t) 127);
  const TilePos tp = tile_pos (tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int width = A.w[comp], height = A.h[comp];
  const int tx0 = tp.bx * T6_W, ty0 = tp.by * T6_H;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  // block separations are powers of two in every Dirac preset: shifts instead of software divides (operands >= 0)
  const int xsh = (xbsep & (xbsep - 1)) == 0 ? __ffs (xbsep) - 1 : -1;
  const int ysh = (ybsep & (ybsep - 1)) == 0 ? __ffs (ybsep) - 1 : -1;
  auto divx = [&] (int v) { return xsh >= 0 ? v >> xsh : v / xbsep; };
  auto divy = [&] (int v) { return ysh >= 0 ? v >> ysh : v / ybsep; };
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, divx (width - xoff));
  const int max_y_blocks = min (A.nby - 1, divy (height - yoff));
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));

  // ---- one bulk tensor copy per reference: the region around the tile, all four phases
  const int rx0 = tx0 - R6_MX, ry0 = ty0 - R6_MY;
  const int nref = A.has_ref1 ? 2 : 1;
  if (STAGED && threadIdx.x == 0) {
    const unsigned bar = s6_u32 (&S.bar);
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nref * R6_BYTES) : "memory");
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r >= nref) break;
      // (selected, not indexed: the descriptor is read from the parameter space itself)
      const CUtensorMap *tm = comp == 0 ? &maps.m[r][0] : comp == 1 ? &maps.m[r][1] : &maps.m[r][2];
#pragma unroll
      for (int ph = 0; ph < 4; ph++) {
        const unsigned dst = s6_u32 (&S.ref[r][128]) + ph * R6_PLANE;
        asm volatile (
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            :: "r"(dst), "l"(tm), "r"(rx0 + B6), "r"(ph), "r"(ry0 + B6), "r"(pic), "r"(bar) : "memory");
      }
    }
  }
  if (threadIdx.x < 64) {
    S.wx[threadIdx.x] = A.wx[comp][threadIdx.x];
    S.wy[threadIdx.x] = A.wy[comp][threadIdx.x];
  }
  for (int t = threadIdx.x; t < (A6_H * A6_P + 3) / 4; t += blockDim.x) reinterpret_cast<int4 *> (S.acc)[t] = make_int4 (0, 0, 0, 0);

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  // ---- blocks overlapping the tile -> table
  const int tw = min (T6_W, width - tx0), th = min (T6_H, height - ty0);
  const int x1 = tx0 + tw - 1, y1 = ty0 + th - 1;
  const int ti0 = (tx0 + xoff - xblen + 1 > 0) ? divx (tx0 + xoff - xblen + xbsep) : 0;
  const int ti1 = min (A.nbx - 1, divx (x1 + xoff));
  const int tj0 = (ty0 + yoff - yblen + 1 > 0) ? divy (ty0 + yoff - yblen + ybsep) : 0;
  const int tj1 = min (A.nby - 1, divy (y1 + yoff));
  const int tni = ti1 - ti0 + 1, tnj = tj1 - tj0 + 1, nblk = tni * tnj;
  const int ipr = (xblen + 7) >> 3;                            // 8-pixel items per block row
  const int lpb = yblen * ipr;                                 // lanes per block

  // A lane adds eight products per pass, step k at accumulator word base + k.  An item holding at
  // most four pixels (the second item of a 12-wide block) does them at steps 4..7, so that at every
  // step the active lanes of a 12x12 block fall on different banks (row pitch 99 == 3 mod 32:
  // 3r and 3r + 4 for r < 12 are 24 different residues).
  if (threadIdx.x >= T6_THREADS - 32) {                        // the last warp: the table rarely needs it
    const int ln = threadIdx.x & 31;
    const int sb = ln / lpb, rem = ln - sb * lpb;
    const int lr = rem / ipr, lh = rem - lr * ipr;
    const bool on = sb < 32 / lpb;
    const int nd = max (1, min (8, xblen - 8 * lh));           // pixels of the item
    const int lt = (lh > 0 && nd <= 4) ? 4 : 0;                // first step of its pixels
    const int w_y = on ? A.wy[comp][lr] : 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      // step k handles item pixel k - lt: its weight sits in the byte lane of the data word it multiplies (dp4a)
      const int q = k - lt, a = 8 * lh + q;
      const int wv = (on && q >= 0 && a < xblen) ? (int) A.wx[comp][a] * w_y : 0;
      S.lane[ln][k] = (unsigned) wv << (8 * (q & 3));
    }
    S.lane[ln][8] = (unsigned) (lr * A6_P + 8 * lh - lt);
    S.lane[ln][9] = (unsigned) (lr * R6_W + 8 * lh);
    S.lane[ln][10] = (unsigned) nd | ((unsigned) lt << 8) | ((on ? 1u : 0u) << 16);
    S.lane[ln][11] = (unsigned) sb | ((unsigned) lr << 8) | ((unsigned) lh << 16);
  }

  for (int t = threadIdx.x; t < nblk; t += blockDim.x) {
    const int jj = t / tni, ii = t - jj * tni;
    const int i = ti0 + ii, j = tj0 + jj;
    const MotionVector *mv = mvs + (size_t) j * A.nbx + i;
    const unsigned flags = __ldg (&mv->flags);
    const int v01 = __ldg (reinterpret_cast<const int *> (mv->v)), v23 = __ldg (reinterpret_cast<const int *> (mv->v) + 1);
    const int v0 = (short) (v01 & 0xffff), v1 = v01 >> 16, v2 = (short) (v23 & 0xffff), v3 = v23 >> 16;
    Blk6 e;
    const int bx = xbsep * i - xoff, by = ybsep * j - yoff;
    unsigned fl = (flags & 3) | ((i >= 1 && i < max_x_blocks && j >= 1 && j < max_y_blocks) ? 4u : 0u);
    e.dc = (short) (comp == 0 ? v0 : comp == 1 ? v1 : v2);
    e.accoff = (unsigned short) ((by - ty0 + A6_MY) * A6_P + (bx - tx0 + A6_MX));
    e.pad = 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      // clamped position, half-pel decomposition and weights exactly as make_blkref, in picture coordinates
      int px = (bx << prec) + ((r ? v1 : v0) >> A.hs[comp]), py = (by << prec) + ((r ? v3 : v2) >> A.vs[comp]);
      const int ee = 32 << prec;
      px = clampi (px, -ee, max_fast_x + ee - 1);
      py = clampi (py, -ee, max_fast_y + ee - 1);
      int rx = 0, ry = 0, hx = px << 1, hy = py << 1;           // prec 0: integer position = even half-pel position
      if (prec == 1) { hx = px; hy = py; }
      else if (prec >= 2) {
        if (prec == 2) { px <<= 1; py <<= 1; }
        hx = px >> 2; hy = py >> 2; rx = px & 3; ry = py & 3;
      }
      const unsigned w00 = (4 - ry) * (4 - rx), w01 = (4 - ry) * rx, w10 = ry * (4 - rx), w11 = ry * rx;
      e.w[r] = w00 | (w01 << 8) | (w10 << 16) | (w11 << 24);
      // every lane of a block reads eight bytes per row, so the whole block (rounded up to 8-pixel items)
      // must lie inside the region, plus the three bytes an unaligned read may touch past its last
      bool inside = true;
      int go[4];
      const int rs = r ? rs1 : rs0;
#pragma unroll
      for (int t2 = 0; t2 < 4; t2++) {
        const int u = hx + (t2 & 1), v = hy + (t2 >> 1);
        const int ph = ((v & 1) << 1) | (u & 1);
        const int ox = (u >> 1) - rx0, oy = (v >> 1) - ry0;
        inside = STAGED && inside && ox >= 0 && oy >= 0 && ox + 8 * ipr + 4 <= R6_W && oy + yblen <= R6_H;
        e.o[r][t2] = (unsigned short) (ph * R6_PLANE + oy * R6_W + ox + R6_BIAS);
        go[t2] = ph * (rs >> 2) + (v >> 1) * rs + (u >> 1);
      }
      if (inside) fl |= 8u << r;
      else {
#pragma unroll
        for (int t2 = 0; t2 < 4; t2++) S.go[t][r][t2] = go[t2];
      }
    }
    e.flags = (unsigned short) fl;
    S.tab[t] = e;
  }
  __syncthreads ();
  if (STAGED) {
    const unsigned bar = s6_u32 (&S.bar);
    unsigned done = 0, tries = 0;
    while (!done) {
      asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(bar) : "memory");
      if (!done && ++tries > (1u << 24)) __trap ();           // a copy that never lands is an error, not a hang
    }
  }
  const unsigned reg0 = STAGED ? s6_u32 (&S.ref[0][128]) - R6_BIAS : 0u, reg1 = STAGED ? s6_u32 (&S.ref[1][128]) - R6_BIAS : 0u;

  // ---- blocks: lane = (block of the pass, block row r, 8-pixel item h); constants from S.lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bpp = 32 / lpb;                                    // blocks per warp pass (lpb <= 32 is a launch condition)
  unsigned wb[8];
  {
    const uint4 l0 = *reinterpret_cast<const uint4 *> (&S.lane[lane][0]), l1 = *reinterpret_cast<const uint4 *> (&S.lane[lane][4]);
    wb[0] = l0.x; wb[1] = l0.y; wb[2] = l0.z; wb[3] = l0.w; wb[4] = l1.x; wb[5] = l1.y; wb[6] = l1.z; wb[7] = l1.w;
  }
  const uint4 l2 = *reinterpret_cast<const uint4 *> (&S.lane[lane][8]);
  const unsigned acc_lane = s6_u32 (S.acc) + 4u * l2.x;
  const unsigned reg_lane = l2.y;
  const int need = l2.z & 0xff, late = (l2.z >> 8) & 0xff;
  const bool lane_on = (l2.z >> 16) != 0;
  const int sblk = (int) l2.w & 0xff, r = ((int) l2.w >> 8) & 0xff, h = (int) l2.w >> 16;
  const unsigned tab0 = s6_u32 (S.tab);

#pragma unroll 1
  for (int b0 = warp * bpp; b0 < nblk; b0 += (T6_THREADS / 32) * bpp) {
    const int t = b0 + sblk;
    if (!lane_on || t >= nblk) continue;
    uint4 e0, e1;
    asm volatile ("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(tab0 + 32u * t));
    asm volatile ("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(tab0 + 32u * t + 16u));
    const int mode = e0.x & 3;
    const bool fast = (e0.x & 4) != 0;
    const int dc = (int) e0.x >> 16;
    const unsigned dst = acc_lane + 4u * (e0.y & 0xffffu);
    uint2 p = make_uint2 (0, 0);
    if (mode != 0) {
      uint2 p0 = make_uint2 (0, 0), p1 = make_uint2 (0, 0);
      if (mode & 1) {
        if (STAGED && (e0.x & 8)) {
          const int o[4] = { (int) (e1.x & 0xffff), (int) (e1.x >> 16), (int) (e1.y & 0xffff), (int) (e1.y >> 16) };
          const unsigned base = reg0 + reg_lane;
          p0 = predict8 (e0.z, o, [&] (int off) { return lds8 (base + off); });
        } else {
          const int4 gv = *reinterpret_cast<const int4 *> (&S.go[t][0][0]);
          const int o[4] = { gv.x, gv.y, gv.z, gv.w };
          const uint8_t *base = ref0 + (ptrdiff_t) r * rs0 + 8 * h;
          p0 = predict8 (e0.z, o, [&] (int off) { return ldg8 (base + off, need); });
        }
      }
      if (mode & 2) {
        if (STAGED && (e0.x & 16)) {
          const int o[4] = { (int) (e1.z & 0xffff), (int) (e1.z >> 16), (int) (e1.w & 0xffff), (int) (e1.w >> 16) };
          const unsigned base = reg1 + reg_lane;
          p1 = predict8 (e0.w, o, [&] (int off) { return lds8 (base + off); });
        } else {
          const int4 gv = *reinterpret_cast<const int4 *> (&S.go[t][1][0]);
          const int o[4] = { gv.x, gv.y, gv.z, gv.w };
          const uint8_t *base = ref1 + (ptrdiff_t) r * rs1 + 8 * h;
          p1 = predict8 (e0.w, o, [&] (int off) { return ldg8 (base + off, need); });
        }
      }
      if (SIMPLE) p = mode == 3 ? make_uint2 (__vavgu4 (p0.x, p1.x), __vavgu4 (p0.y, p1.y)) : (mode == 1 ? p0 : p1);
      else {
        {
          // general weights: per pixel, 16-bit wrapping arithmetic of the reference
          int v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int s0 = (int) (((k < 4 ? p0.x : p0.y) >> (8 * (k & 3))) & 0xff);
            const int s1 = (int) (((k < 4 ? p1.x : p1.y) >> (8 * (k & 3))) & 0xff);
            v[k] = obmc_combine<false> (A, mode, fast, noscale, dc, s0, s1);
          }
          if (fast) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int w = (int) (wb[k] >> (8 * (k & 3)));
              red_if (wb[k], dst + 4 * k, (unsigned) ((late ? v[k & 3] : v[k]) * w));
            }
            continue;
          }
          // picture-edge blocks absorb the weight of the missing neighbour (schromotion8.c:673-693)
          const int arow = (int) (e0.y & 0xffffu) / A6_P, acol = (int) (e0.y & 0xffffu) - arow * A6_P;
          const int bx = acol - A6_MX + tx0, by = arow - A6_MY + ty0, y = by + r;
          int wy2 = S.wy[r];
          if (y < yoff) wy2 += S.wy[2 * yoff - r - 1];
          if (y >= A.nby * ybsep - yoff) wy2 += S.wy[2 * (yblen - yoff) - r - 1];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int a = 8 * h + k, x = bx + a;
            if (a >= xblen) break;
            int wx2 = S.wx[a];
            if (x < xoff) wx2 += S.wx[2 * xoff - a - 1];
            if (x >= A.nbx * xbsep - xoff) wx2 += S.wx[2 * (xblen - xoff) - a - 1];
            asm volatile ("red.shared.add.u32 [%0], %1;" :: "r"(dst + 4 * (k + late)), "r"(v[k] * wx2 * wy2) : "memory");
          }
          continue;
        }
      }
    }
    if (fast && mode != 0) {
      // ---- the common case: product of byte k of the item and its weight in one dp4a, one reduction a step
      const unsigned plo = late ? 0u : p.x, phi = late ? p.x : p.y;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const unsigned c = __dp4a (k < 4 ? plo : phi, wb[k], 0u);
        red_if (wb[k], dst + 4 * k, c);                        // (a lane constant: steps without a pixel are skipped, not branched over)
      }
    } else if (fast) {
      const int dcv = w16 (dc + 128);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int w = (int) (wb[k] >> (8 * (k & 3)));
        red_if (wb[k], dst + 4 * k, (unsigned) (dcv * w));
      }
    } else {
      // picture-edge blocks absorb the weight of the missing neighbour (schromotion8.c:673-693)
      const int arow = (int) (e0.y & 0xffffu) / A6_P, acol = (int) (e0.y & 0xffffu) - arow * A6_P;
      const int bx = acol - A6_MX + tx0, by = arow - A6_MY + ty0, y = by + r;
      int wy2 = S.wy[r];
      if (y < yoff) wy2 += S.wy[2 * yoff - r - 1];
      if (y >= A.nby * ybsep - yoff) wy2 += S.wy[2 * (yblen - yoff) - r - 1];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int a = 8 * h + k, x = bx + a;
        if (a >= xblen) break;
        int wx2 = S.wx[a];
        if (x < xoff) wx2 += S.wx[2 * xoff - a - 1];
        if (x >= A.nbx * xbsep - xoff) wx2 += S.wx[2 * (xblen - xoff) - a - 1];
        const int v = mode == 0 ? ((dc + 128) & 0xff) : (int) (((k < 4 ? p.x : p.y) >> (8 * (k & 3))) & 0xff);
        asm volatile ("red.shared.add.u32 [%0], %1;" :: "r"(dst + 4 * (k + late)), "r"(v * wx2 * wy2) : "memory");
      }
    }
  }
  __syncthreads ();

  // ---- finish: four pixels per thread (schromotion8.c:809-921; schroorc.orc:636-673)
  for (int t = threadIdx.x; t < (T6_W / 4) * th; t += blockDim.x) {
    const int ry = t / (T6_W / 4), lx = (t - ry * (T6_W / 4)) * 4;
    const int x = tx0 + lx, y = ty0 + ry;
    if (x >= width) continue;
    const int npx = min (4, width - x);
    const int *arow_s = &S.acc[(ry + A6_MY) * A6_P + lx + A6_MX];
    const int sum[4] = { arow_s[0], arow_s[1], arow_s[2], arow_s[3] };
    const size_t ro = (size_t) y * A.res.stride[comp];
    if (A.add) {
      int rr[4];
      const char *rrow = plane_ptr (A.res, pic, comp) + ro;
      const bool vec = npx == 4 && (((size_t) (rrow + (size_t) x * (A.res_is_s32 ? 4 : 2))) & (A.res_is_s32 ? 15 : 7)) == 0;
      if (vec) {
        if (A.res_is_s32) {
          const int4 q = *reinterpret_cast<const int4 *> (rrow + (size_t) x * 4);
          rr[0] = w16 (q.x); rr[1] = w16 (q.y); rr[2] = w16 (q.z); rr[3] = w16 (q.w);
        } else {
          const int2 q = *reinterpret_cast<const int2 *> (rrow + (size_t) x * 2);
          rr[0] = (q.x << 16) >> 16; rr[1] = q.x >> 16; rr[2] = (q.y << 16) >> 16; rr[3] = q.y >> 16;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          rr[k] = k < npx ? (A.res_is_s32 ? w16 (reinterpret_cast<const int *> (rrow)[x + k])
                                          : (int) reinterpret_cast<const short *> (rrow)[x + k]) : 0;
      }
      unsigned packed = 0;
      int a16[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a16[k] = w16 (sum[k]);
        int tt = w16 (a16[k] + 32) >> 6;
        tt = w16 (rr[k] + tt);
        packed |= (unsigned) clampi (tt, 0, 255) << (8 * k);
      }
      uint8_t *orow = reinterpret_cast<uint8_t *> (plane_ptr (A.out, pic, comp)) + (size_t) y * A.out.stride[comp] + x;
      if (npx == 4 && (((size_t) orow) & 3) == 0) *reinterpret_cast<unsigned *> (orow) = packed;
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

// ---- host side -----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiled6) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled6 encode_tiled6 ()
{
  static EncodeTiled6 fn = nullptr;
  static std::once_flag once;
  std::call_once (once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled6> (p);
  });
  return fn;
}

// the four phase planes of component c of every picture of a reference slab as a 4-D tensor of
// bytes: (x, phase, y, picture), origin at the corner of the 32-pixel border
static bool make_ref_map6 (CUtensorMap *tm, const sb2_slab *s, int c)
{
  EncodeTiled6 enc = encode_tiled6 ();
  if (!enc) return false;
  const int stride = s->stride[c];
  if ((stride & 63) || (s->picture_pitch & 15)) return false;                 // phase pitch must be a multiple of 16
  if (s->offset[c] < (size_t) B6 * stride + B6) return false;
  char *base = static_cast<char *> (s->base) + s->offset[c] - (size_t) B6 * stride - B6;
  if ((size_t) base & 15) return false;
  const cuuint64_t dims[4] = { (cuuint64_t) (s->width[c] + 2 * B6), 4, (cuuint64_t) (s->height[c] + 2 * B6), (cuuint64_t) s->count };
  if (dims[0] > (cuuint64_t) (stride >> 2)) return false;
  const cuuint64_t strides[3] = { (cuuint64_t) (stride >> 2), (cuuint64_t) stride, (cuuint64_t) s->picture_pitch };
  const cuuint32_t box[4] = { R6_W, 1, R6_H, 1 };            // one phase a copy
  const cuuint32_t estr[4] = { 1, 1, 1, 1 };
  return enc (tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int obmc_blocks_launch (const ObmcArgs &A, const sb2_slab *ref0, const sb2_slab *ref1, int count, bool staged, cudaStream_t st)
{
  for (int c = 0; c < A.ncomp; c++) {
    const int ipr = (A.xblen[c] + 7) >> 3;
    // a block's lanes fit a warp, the block fits the accumulator margin, every block of a tile fits the table
    if (A.yblen[c] * ipr > 32 || 8 * ipr > A6_MX || A.yblen[c] > A6_MY || A.xblen[c] < 1) return SB2_ERR_UNSUPPORTED;
    const int ni = (T6_W + A.xblen[c] - 2) / A.xbsep[c] + 1, nj = (T6_H + A.yblen[c] - 2) / A.ybsep[c] + 1;
    if (ni * nj > T6_MAXB) return SB2_ERR_UNSUPPORTED;
  }
  if (A.ncomp > 3 || count > 65535) return SB2_ERR_UNSUPPORTED;
  Maps6 maps;
  memset (&maps, 0, sizeof (maps));
  for (int c = 0; c < A.ncomp; c++) {
    if (staged) {
      if (!make_ref_map6 (&maps.m[0][c], ref0, c)) return SB2_ERR_UNSUPPORTED;
      if (ref1 && !make_ref_map6 (&maps.m[1][c], ref1, c)) return SB2_ERR_UNSUPPORTED;
    }
    // the global path (all of the unstaged kernel, the outliers of the staged one) loads 8-byte words
    if ((ref0->stride[c] & 31) || (ref1 && (ref1->stride[c] & 31)) || (ref0->offset[c] & 7) || (ref1 && (ref1->offset[c] & 7)))
      return SB2_ERR_UNSUPPORTED;
  }
  if ((((size_t) ref0->base | ref0->picture_pitch) & 7) || (ref1 && (((size_t) ref1->base | ref1->picture_pitch) & 7)))
    return SB2_ERR_UNSUPPORTED;
  TileGrid tiles;
  const dim3 grid = make_tile_grid (tiles, A.ncomp, A.w, A.h, T6_W, T6_H, count);
  const size_t smem = (staged ? sizeof (Smem6) : SMEM6_UNSTAGED) + 128;
  const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once (once, [&] {
    const int big = (int) sizeof (Smem6) + 128, small = (int) SMEM6_UNSTAGED + 128;
    attr_err = cudaFuncSetAttribute (obmc_kernel_blocks<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute (obmc_kernel_blocks<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (attr_err == cudaSuccess && small > 48 * 1024) {
      attr_err = cudaFuncSetAttribute (obmc_kernel_blocks<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute (obmc_kernel_blocks<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    }
  });
  if (attr_err != cudaSuccess) return SB2_ERR_UNSUPPORTED;
  if (staged) {
    if (simple) obmc_kernel_blocks<true, true><<<grid, T6_THREADS, smem, st>>> (A, tiles, maps);
    else obmc_kernel_blocks<false, true><<<grid, T6_THREADS, smem, st>>> (A, tiles, maps);
  } else {
    if (simple) obmc_kernel_blocks<true, false><<<grid, T6_THREADS, smem, st>>> (A, tiles, maps);
    else obmc_kernel_blocks<false, false><<<grid, T6_THREADS, smem, st>>> (A, tiles, maps);
  }
  return SB2_OK;
}

}  // namespace sb2
