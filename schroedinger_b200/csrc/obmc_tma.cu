// obmc_tma.cu -- OBMC renderer: TMA-staged reference windows, gather per output pixel.
//
// Same result, bit for bit, as schro_motion_render_u8 (schroedinger/schromotion8.c:700-929) with the
// sub-pel block fetch of schroedinger/schroframe.c:2288-2482; obmc.cu holds the kernels for every
// geometry this one does not cover and the C entry point.
//
// A CTA owns a 64x64 tile of one component of one picture.  The reference samples every block of the
// tile can need lie within +-(motion range) of the tile: ONE bulk tensor copy per reference
// (cp.async.bulk.tensor.4d: 112 bytes x 4 half-pel phases x 104 rows, the phases being a dimension of
// the tensor because the frame layout keeps them side by side in every row) brings that region into
// shared memory behind an mbarrier while the CTA decodes the motion vectors of its blocks into a
// table (mode, clamped position, phase offsets and bilinear weights per reference -- every sub-pel
// case of schroframe.c:2288-2413 is one 4-tap sum because (8a+8b+8)>>4 == avgub(a,b)).  A thread then
// owns four adjacent pixels of four rows and GATHERS the blocks that cover them: taps are read from
// the staged region (no global loads, no per-pixel address clamping), applied to two pixels at a
// time in packed 16-bit lanes, multiplied by the OBMC window -- whose picture-edge folding
// (schromotion8.c:673-693) is a function of pixel and block index only, so it is part of the
// per-thread weights -- and summed in registers; the finish (rrshift6 add + clamp, or subtract) and
// the stores follow in the same thread.  No accumulator in memory, no atomics.  A block whose
// vector points outside the staged region (outliers) reads its taps from global memory instead.

#include "obmc_common.cuh"
#include <cuda.h>
#include <cstring>
#include <mutex>

namespace sb2 {

constexpr int T5_W = 64, T5_H = 64;            // output tile
constexpr int T5_THREADS = 256;                // 16 four-pixel groups x 16 rows, four rows per thread (two CTAs per SM:
                                               // 128 registers per thread -- 512 threads at 64 registers spilled)
constexpr int T5_ROWSTEP = T5_THREADS / 16;
constexpr int R5_MX = 16, R5_MY = 20;          // staged margin left / top of the tile.  The x coordinate of a tensor copy must be
                                               // a multiple of 16 bytes (measured on B200: anything else raises "illegal
                                               // instruction"), so the margin is 16 and not the 20 the vertical one has
constexpr int R5_W = 112, R5_H = 104;          // staged region: 16 + 64 + 32 columns, 20 + 64 + 20 rows
constexpr int R5_ROW = 4 * R5_W;               // the copy lands as [row][phase][x]: bytes per row
constexpr int R5_BYTES = R5_ROW * R5_H;        // 46592 per reference
constexpr int R5_BIAS = 8192;                  // added to the stored tap offsets: a block that starts above / left of the region
                                               // (only its lower / right part lies in the tile) has a negative corner offset
constexpr int T5_MAXOUT = 32;                  // (block, reference) pairs of a tile whose window is not inside the staged region
constexpr int T5_MAXB = 400;                   // blocks overlapping a tile (4:2:0 chroma with 6/4 blocks: 18 x 18)
constexpr int BORDER = 32;                     // frame extension the renderer requires (schromotion8.c:303-335)

struct Blk5 {
  short mode, fast, dc, staged;                // staged: bit r set = reference r's taps lie inside the staged region
  unsigned w[2];                               // tap weights, four bytes
  unsigned short o[2][4];                      // R5_BIAS + byte offset of tap t of pixel (a = 0, b = 0) inside the staged region;
                                               // not staged: o[r][0] = index into the outlier table
};
static_assert (sizeof (Blk5) == 32, "two 16-byte reads per table entry");

struct ObmcMaps { CUtensorMap m[2][3]; };


struct Smem5 {
  alignas (128) unsigned char ref[2][R5_BYTES + 128];      // [128 spare bytes][region] per reference: a group that starts
                                                           // left of its block reads a few bytes in front of the region
  alignas (16) Blk5 tab[T5_MAXB];
  int out_o[T5_MAXOUT][4];                                 // outliers: tap offsets into the reference plane (as BlkRef::o)
  int nout;
  unsigned char wx[64], wy[64];
  alignas (8) unsigned long long bar;
};

__device__ __forceinline__ unsigned smem_u32 (const void *p) { return (unsigned) __cvta_generic_to_shared (p); }

__device__ __forceinline__ unsigned lds_u32_unaligned (unsigned addr)
{
  const unsigned al = addr & ~3u;
  unsigned w0, w1;
  asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(al));
  asm volatile ("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(al));
  return __funnelshift_r (w0, w1, (addr & 3u) * 8u);
}

// 4-tap sum of four adjacent pixels out of the staged region: returns (p0 | p1<<16, p2 | p3<<16)
__device__ __forceinline__ uint2 fetch4x4_staged (unsigned region, const Blk5 &e, int r, int pix)
{
  const unsigned w = e.w[r];
  unsigned lo = 0x00080008u, hi = 0x00080008u;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const unsigned wt = (w >> (8 * t)) & 0xff;
    const unsigned b = lds_u32_unaligned (region + e.o[r][t] + pix);
    lo += wt * __byte_perm (b, 0, 0x4140);
    hi += wt * __byte_perm (b, 0, 0x4342);
  }
  return make_uint2 ((lo >> 4) & 0x0fff0fffu, (hi >> 4) & 0x0fff0fffu);
}

template <bool SIMPLE>
__global__ void __launch_bounds__ (T5_THREADS, 2)
obmc_kernel_tma (const ObmcArgs A, const TileGrid tiles, const __grid_constant__ ObmcMaps maps)
{
  extern __shared__ unsigned char smem_raw[];
  Smem5 &S = *reinterpret_cast<Smem5 *> ((reinterpret_cast<size_t> (smem_raw) + 127) & ~(size_t) 127);
  const TilePos tp = tile_pos (tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int width = A.w[comp], height = A.h[comp];
  const int tx0 = tp.bx * T5_W, ty0 = tp.by * T5_H;

  const int xbsep = A.xbsep[comp], ybsep = A.ybsep[comp], xblen = A.xblen[comp], yblen = A.yblen[comp];
  const int xoff = (xblen - xbsep) >> 1, yoff = (yblen - ybsep) >> 1;
  const int prec = A.prec;
  const int max_fast_x = (width - xblen) << prec, max_fast_y = (height - yblen) << prec;
  const int max_x_blocks = min (A.nbx - 1, (width - xoff) / xbsep);
  const int max_y_blocks = min (A.nby - 1, (height - yoff) / ybsep);
  const bool noscale = (A.w1 + A.w2 == (1 << A.bits));

  // ---- one bulk tensor copy per reference: the region around the tile, all four phases
  const int rx0 = tx0 - R5_MX, ry0 = ty0 - R5_MY;           // picture coordinates of the region's corner
  const int nref = A.has_ref1 ? 2 : 1;
  if (threadIdx.x == 0) {
    const unsigned bar = smem_u32 (&S.bar);
    S.nout = 0;
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nref * R5_BYTES) : "memory");
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r >= nref) break;
      const unsigned dst = smem_u32 (&S.ref[r][128]);
      // (selected, not indexed: the descriptor must be read from the parameter space itself)
      const CUtensorMap *tm = comp == 0 ? &maps.m[r][0] : comp == 1 ? &maps.m[r][1] : &maps.m[r][2];
      asm volatile (
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
          :: "r"(dst), "l"(tm), "r"(rx0 + BORDER), "r"(0), "r"(ry0 + BORDER), "r"(pic), "r"(bar) : "memory");
    }
  }

  if (threadIdx.x < 64) {
    S.wx[threadIdx.x] = A.wx[comp][threadIdx.x];
    S.wy[threadIdx.x] = A.wy[comp][threadIdx.x];
  }
  __syncthreads ();

  const uint8_t *ref0 = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref0, pic, comp));
  const uint8_t *ref1 = A.has_ref1 ? reinterpret_cast<const uint8_t *> (plane_ptr (A.ref1, pic, comp)) : ref0;
  const int rs0 = A.ref0.stride[comp], rs1 = A.has_ref1 ? A.ref1.stride[comp] : rs0;
  const MotionVector *mvs = A.mvs + (size_t) pic * A.mv_pitch;

  // ---- blocks overlapping the tile -> table
  const int tw = min (T5_W, width - tx0), th = min (T5_H, height - ty0);
  const int x1 = tx0 + tw - 1, y1 = ty0 + th - 1;
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
    Blk5 e;
    e.mode = (short) (flags & 3);
    e.fast = (short) (i >= 1 && i < max_x_blocks && j >= 1 && j < max_y_blocks);
    e.dc = (short) (comp == 0 ? v0 : comp == 1 ? v1 : v2);
    e.staged = 0;
    const int bx = xbsep * i - xoff, by = ybsep * j - yoff;
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
      // only the part of the block that lies inside the tile is read (plus up to three columns on
      // either side: threads work on aligned groups of four pixels)
      const int a_lo = max (0, tx0 - bx) - 3, a_hi = min (xblen, tx0 + tw - bx) + 3;
      const int b_lo = max (0, ty0 - by), b_hi = min (yblen, ty0 + th - by);
      bool inside = true;
      int go[4];
      const int rs = r ? rs1 : rs0;
#pragma unroll
      for (int t2 = 0; t2 < 4; t2++) {
        const int u = hx + (t2 & 1), v = hy + (t2 >> 1);
        const int ph = ((v & 1) << 1) | (u & 1);
        const int ox = (u >> 1) - rx0, oy = (v >> 1) - ry0;      // window corner inside the region
        inside = inside && ox + a_lo >= 0 && oy + b_lo >= 0 && ox + a_hi + 4 <= R5_W && oy + b_hi <= R5_H;
        e.o[r][t2] = (unsigned short) (oy * R5_ROW + ph * R5_W + ox + R5_BIAS);
        go[t2] = ph * (rs >> 2) + (v >> 1) * rs + (u >> 1);
      }
      const bool used = (e.mode >> r) & 1;
      if (inside) e.staged |= (short) (1 << r);
      else if (used) {
        const int k = atomicAdd (&S.nout, 1);
        if (k < T5_MAXOUT) {
          e.o[r][0] = (unsigned short) k;
#pragma unroll
          for (int t2 = 0; t2 < 4; t2++) S.out_o[k][t2] = go[t2];
        } else {
          e.o[r][0] = 0xffff;                                     // table full: the vector is decoded again at every use
        }
      }
    }
    S.tab[t] = e;
  }
  __syncthreads ();

  // ---- per-thread geometry: four pixels x0..x0+3 of rows y (two of them), the block columns /
  // rows that cover them and their window weights (picture-edge folding included)
  const int g = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int x0 = tx0 + 4 * g;
  const bool col_in = x0 < width;
  const int npx = min (4, width - x0);
  const int xl = min (x0 + 3, width - 1);
  const int i_lo = (x0 + xoff - xblen + 1 > 0) ? (x0 + xoff - xblen + xbsep) / xbsep : 0;
  const int i_hi = min (A.nbx - 1, (xl + xoff) / xbsep);
  unsigned wxp[3] = { 0, 0, 0 };                                // packed weights of my four pixels for block column i_lo + c
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int i = i_lo + c;
    if (i > i_hi) continue;
    const int bx = xbsep * i - xoff;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int x = x0 + k, a = x - bx;
      unsigned w = 0;
      if (a >= 0 && a < xblen && k < npx) {
        w = S.wx[a];
        if (x < xoff) w += S.wx[2 * xoff - a - 1];
        if (x >= A.nbx * xbsep - xoff) w += S.wx[2 * (xblen - xoff) - a - 1];
      }
      wxp[c] |= w << (8 * k);
    }
  }

  // the staged regions must have landed before the first tap is read
  {
    const unsigned bar = smem_u32 (&S.bar);
    unsigned done = 0;
    while (!done)
      asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(bar) : "memory");
  }
  const unsigned reg0 = smem_u32 (&S.ref[0][128]) - R5_BIAS, reg1 = smem_u32 (&S.ref[1][128]) - R5_BIAS;

#pragma unroll 1
  for (int q = 0; q < T5_H / T5_ROWSTEP; q++) {
    const int y = ty0 + ty + T5_ROWSTEP * q;
    if (y >= height || !col_in) continue;
    const int j_lo = (y + yoff - yblen + 1 > 0) ? (y + yoff - yblen + ybsep) / ybsep : 0;
    const int j_hi = min (A.nby - 1, (y + yoff) / ybsep);
    int sum[4] = { 0, 0, 0, 0 };
    for (int j = j_lo; j <= j_hi; j++) {
      const int by = ybsep * j - yoff, b = y - by;
      int w_y = S.wy[b];
      if (y < yoff) w_y += S.wy[2 * yoff - b - 1];
      if (y >= A.nby * ybsep - yoff) w_y += S.wy[2 * (yblen - yoff) - b - 1];
      int rowsum[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int i = i_lo + c;
        if (i > i_hi) break;
        const Blk5 e = S.tab[(j - tj0) * tni + (i - ti0)];
        const int mode = e.mode;
        const bool fast = e.fast != 0;
        const int a0 = x0 - (xbsep * i - xoff);                 // block column of my first pixel (may be negative)
        // a block whose vector leaves the staged region (the +-4000 outliers of a stream): its taps come
        // from global memory (tap offsets decoded once per tile into a small side table)
        auto fetch_outlier = [&] (const uint8_t *ref, int rs, const MotionVector *mv, int r) -> uint2 {
          BlkRef br;
          br.w = e.w[r];
          if (e.o[r][0] != 0xffff) {
#pragma unroll
            for (int t2 = 0; t2 < 4; t2++) br.o[t2] = S.out_o[e.o[r][0]][t2];
          } else {
            make_blkref (br, rs, prec, xbsep * i - xoff, by, (int) mv->v[r] >> A.hs[comp], (int) mv->v[2 + r] >> A.vs[comp],
                max_fast_x, max_fast_y);
          }
          if (a0 >= 0) return fetch4x4 (ref, br, b * rs + a0);     // (reads up to three bytes past the block: inside the plane)
          unsigned p[4];
#pragma unroll
          for (int k = 0; k < 4; k++) p[k] = (unsigned) fetch1 (ref, br, b * rs + max (a0 + k, 0));
          return make_uint2 (p[0] | (p[1] << 16), p[2] | (p[3] << 16));
        };
        int v[4];
        if (mode == 0) {
          const int dcv = fast ? w16 ((int) e.dc + 128) : (((int) e.dc + 128) & 0xff);
          v[0] = v[1] = v[2] = v[3] = dcv;
        } else {
          uint2 s0 = make_uint2 (0, 0), s1 = make_uint2 (0, 0);
          if (mode & 1) {
            if (e.staged & 1) s0 = fetch4x4_staged (reg0, e, 0, b * R5_ROW + a0);
            else s0 = fetch_outlier (ref0, rs0, mvs + (size_t) j * A.nbx + i, 0);
          }
          if (mode & 2) {
            if (e.staged & 2) s1 = fetch4x4_staged (reg1, e, 1, b * R5_ROW + a0);
            else s1 = fetch_outlier (ref1, rs1, mvs + (size_t) j * A.nbx + i, 1);
          }
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
            const int p0[4] = { (int) (s0.x & 0xffff), (int) (s0.x >> 16), (int) (s0.y & 0xffff), (int) (s0.y >> 16) };
            const int p1[4] = { (int) (s1.x & 0xffff), (int) (s1.x >> 16), (int) (s1.y & 0xffff), (int) (s1.y >> 16) };
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = obmc_combine<false> (A, mode, fast, noscale, e.dc, p0[k], p1[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) rowsum[k] += v[k] * (int) ((wxp[c] >> (8 * k)) & 0xff);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) sum[k] += rowsum[k] * w_y;
    }

    // ---- finish (schromotion8.c:809-921; schroorc.orc:636-673)
    const size_t ro = (size_t) y * A.res.stride[comp];
    if (A.add) {
      int r[4];
      const char *rrow = plane_ptr (A.res, pic, comp) + ro;
      if (npx == 4) {
        if (A.res_is_s32) {
          const int4 q = *reinterpret_cast<const int4 *> (rrow + (size_t) x0 * 4);
          r[0] = w16 (q.x); r[1] = w16 (q.y); r[2] = w16 (q.z); r[3] = w16 (q.w);
        } else {
          const int2 q = *reinterpret_cast<const int2 *> (rrow + (size_t) x0 * 2);
          r[0] = (q.x << 16) >> 16; r[1] = q.x >> 16; r[2] = (q.y << 16) >> 16; r[3] = q.y >> 16;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          r[k] = k < npx ? (A.res_is_s32 ? w16 (reinterpret_cast<const int *> (rrow)[x0 + k])
                                         : (int) reinterpret_cast<const short *> (rrow)[x0 + k]) : 0;
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
      uint8_t *orow = reinterpret_cast<uint8_t *> (plane_ptr (A.out, pic, comp)) + (size_t) y * A.out.stride[comp] + x0;
      if (npx == 4) *reinterpret_cast<unsigned *> (orow) = packed;
      else for (int k = 0; k < npx; k++) orow[k] = (uint8_t) (packed >> (8 * k));
      if (A.has_acc) {
        short *arow = reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp]) + x0;
        for (int k = 0; k < npx; k++) arow[k] = (short) a16[k];
      }
    } else {
      short *rrow = reinterpret_cast<short *> (plane_ptr (A.res, pic, comp) + ro) + x0;
      short *arow = A.has_acc ? reinterpret_cast<short *> (plane_ptr (A.acc, pic, comp) + (size_t) y * A.acc.stride[comp]) + x0 : nullptr;
      for (int k = 0; k < npx; k++) {
        const int tt = w16 (w16 (sum[k]) - 8160) >> 6;
        rrow[k] = (short) w16 (rrow[k] - tt);
        if (arow) arow[k] = (short) tt;
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled ()
{
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once (once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn> (p);
  });
  return fn;
}

// the four phase planes of component c of every picture of a reference slab as a 4-D tensor of
// bytes: (x, phase, y, picture), origin at the corner of the 32-pixel border
static bool make_ref_map (CUtensorMap *tm, const sb2_slab *s, int c)
{
  EncodeTiledFn enc = encode_tiled ();
  if (!enc) return false;
  const int stride = s->stride[c];
  if ((stride & 63) || (s->picture_pitch & 15)) return false;                 // phase pitch must be a multiple of 16
  const size_t corner = s->offset[c] - (size_t) BORDER * stride - BORDER;
  if (s->offset[c] < (size_t) BORDER * stride + BORDER) return false;
  char *base = static_cast<char *> (s->base) + corner;
  if ((size_t) base & 15) return false;
  const cuuint64_t dims[4] = { (cuuint64_t) (s->width[c] + 2 * BORDER), 4, (cuuint64_t) (s->height[c] + 2 * BORDER),
                               (cuuint64_t) s->count };
  if (dims[0] > (cuuint64_t) (stride >> 2)) return false;
  const cuuint64_t strides[3] = { (cuuint64_t) (stride >> 2), (cuuint64_t) stride, (cuuint64_t) s->picture_pitch };
  const cuuint32_t box[4] = { R5_W, 4, R5_H, 1 };
  const cuuint32_t estr[4] = { 1, 1, 1, 1 };
  return enc (tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int obmc_tma_launch (const ObmcArgs &A, const sb2_slab *ref0, const sb2_slab *ref1, int count, cudaStream_t st)
{
  for (int c = 0; c < A.ncomp; c++) {
    // one block of overlap at most, and every block of a tile fits the table
    if (A.xblen[c] > 2 * A.xbsep[c] || A.yblen[c] > 2 * A.ybsep[c] || A.xblen[c] + 3 > R5_W || A.yblen[c] > R5_H) return SB2_ERR_UNSUPPORTED;
    const int ni = (T5_W + A.xblen[c]) / A.xbsep[c] + 2, nj = (T5_H + A.yblen[c]) / A.ybsep[c] + 2;
    if (ni * nj > T5_MAXB) return SB2_ERR_UNSUPPORTED;
    // three covering block columns per four-pixel group at most
    if ((A.xblen[c] + 2) / A.xbsep[c] + 1 > 3) return SB2_ERR_UNSUPPORTED;
  }
  if (A.ncomp > 3 || count > 65535) return SB2_ERR_UNSUPPORTED;
  ObmcMaps maps;
  memset (&maps, 0, sizeof (maps));
  for (int c = 0; c < A.ncomp; c++) {
    if (!make_ref_map (&maps.m[0][c], ref0, c)) return SB2_ERR_UNSUPPORTED;
    if (ref1 && !make_ref_map (&maps.m[1][c], ref1, c)) return SB2_ERR_UNSUPPORTED;
  }
  TileGrid tiles;
  const dim3 grid = make_tile_grid (tiles, A.ncomp, A.w, A.h, T5_W, T5_H, count);
  const size_t smem = sizeof (Smem5) + 128;
  const bool simple = (A.w1 == 1 && A.w2 == 1 && A.bits == 1);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once (once, [&] {
    attr_err = cudaFuncSetAttribute (obmc_kernel_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute (obmc_kernel_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  });
  if (attr_err != cudaSuccess) return SB2_ERR_UNSUPPORTED;
  if (simple) obmc_kernel_tma<true><<<grid, T5_THREADS, smem, st>>> (A, tiles, maps);
  else obmc_kernel_tma<false><<<grid, T5_THREADS, smem, st>>> (A, tiles, maps);
  return SB2_OK;
}


}  // namespace sb2
