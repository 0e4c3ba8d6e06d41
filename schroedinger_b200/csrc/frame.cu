// frame.cu -- reference-frame preparation for sm_100a: motion-compensation edge
// extension, the 8-tap half-pel upsampler (all three phases and every border in one
// launch) and the 4-tap pyramid downsampler.  u8, bit-exact with
//   schro_frame_mc_edgeextend        schroedinger/schroframe.c:1940-1997
//   schro_upsampled_frame_upsample   schroedinger/schroframe.c:2000-2030
//     (schro_frame_upsample_vert / _horiz, :1515-1645)
//   schro_frame_downsample           schroedinger/schroframe.c:1400-1513
//
// The reference fills the phase planes pass by pass (vertical, horizontal, horizontal of
// vertical) and patches the borders after each pass from specific sources.  Here every
// pixel of the extended phase planes is a closed-form function of phase 0 (DESIGN.md
// "upsample"), evaluated from one shared-memory tile: phase 0 is read once, the three
// phases are written once, nothing is re-read.

#include "common.cuh"
#include <cstdio>
#include <cstdlib>

namespace sb2 {

struct FrameArgs {
  TileGrid tiles;
  PlaneSet planes;              // phase-0 pixel (0,0) of every component
  int w[SB2_MAX_COMPONENTS];
  int h[SB2_MAX_COMPONENTS];
  int ncomp;
  int ext;
  int fuse_edge;                // upsample: also write the phase-0 border (edge extension)
};

__device__ __forceinline__ int clampi (int x, int lo, int hi) { return min (max (x, lo), hi); }
__device__ __forceinline__ int clamp255 (int x) { return min (max (x, 0), 255); }

// ---- edge extension ---------------------------------------------------------------
// one thread per border pixel: rows [-ext, h+ext) x columns outside [0,w), plus the
// rows above/below for columns [0,w)
__global__ void __launch_bounds__ (256)
edgeextend_kernel (const FrameArgs a)
{
  const int comp = blockIdx.z % a.ncomp, pic = blockIdx.z / a.ncomp;
  const int w = a.w[comp], h = a.h[comp], ext = a.ext;
  uint8_t *p = reinterpret_cast<uint8_t *> (plane_ptr (a.planes, pic, comp));
  const int stride = a.planes.stride[comp];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (((((size_t) p | (size_t) stride) & 3) == 0) && (w & 3) == 0 && (ext & 3) == 0 && ext > 0) {
    // word path.  Side strips: one item per (row, side, word of the strip), the edge pixel
    // replicated into all four bytes.  Caps: one item per (word column incl. the corners, top /
    // bottom), the source word of row 0 / h-1 (corner columns: the replicated corner pixel) stored
    // to `ext` rows, consecutive threads on consecutive words.
    const int wps = ext >> 2;                                   // words per strip
    const int nside = h * 2 * wps;
    for (int i = tid; i < nside; i += nth) {
      const int y = i / (2 * wps), c = i - y * (2 * wps);
      const bool right = c >= wps;
      uint8_t *row = p + (ptrdiff_t) y * stride;
      const unsigned v = right ? row[w - 1] : row[0];
      unsigned *dst = reinterpret_cast<unsigned *> (right ? row + w : row - ext) + (right ? c - wps : c);
      *dst = v * 0x01010101u;
    }
    const int wpr = (w + 2 * ext) >> 2;                         // words per extended row
    for (int i = tid; i < 2 * wpr; i += nth) {
      const bool bottom = i >= wpr;
      const int c = bottom ? i - wpr : i;                       // word column of the extended row
      const int x = 4 * c - ext;
      const uint8_t *srow = p + (ptrdiff_t) (bottom ? h - 1 : 0) * stride;
      unsigned v;
      if (x < 0) v = srow[0] * 0x01010101u;
      else if (x >= w) v = srow[w - 1] * 0x01010101u;
      else v = *reinterpret_cast<const unsigned *> (srow + x);
      uint8_t *d0 = p + (ptrdiff_t) (bottom ? h : -ext) * stride + x;
      for (int r = 0; r < ext; r++) *reinterpret_cast<unsigned *> (d0 + (ptrdiff_t) r * stride) = v;
    }
    return;
  }
  const int side = 2 * ext * (h + 2 * ext);       // left+right strips, all rows
  const int caps = 2 * ext * w;                   // top+bottom caps, interior columns
  for (int i = tid; i < side + caps; i += nth) {
    int x, y;
    if (i < side) {
      y = i / (2 * ext) - ext;
      const int c = i % (2 * ext);
      x = c < ext ? c - ext : w + (c - ext);
    } else {
      const int k = i - side;
      const int r = k / w;
      x = k % w;
      y = r < ext ? r - ext : h + (r - ext);
    }
    p[(ptrdiff_t) y * stride + x] = p[(ptrdiff_t) clampi (y, 0, h - 1) * stride + clampi (x, 0, w - 1)];
  }
}

// ---- upsample ---------------------------------------------------------------------

__device__ __forceinline__ int taps8 (const uint8_t *s, int step)
{
  // (-1, 3, -7, 21, 21, -7, 3, -1), +16 >> 5   (schroframe.c:1562, 1615)
  int acc = 16;
  acc += 21 * ((int) s[3 * step] + (int) s[4 * step]);
  acc -= 7 * ((int) s[2 * step] + (int) s[5 * step]);
  acc += 3 * ((int) s[1 * step] + (int) s[6 * step]);
  acc -= ((int) s[0] + (int) s[7 * step]);
  return clamp255 (acc >> 5);
}

// ---- upsample, the usual case: whole words, dp4a ------------------------------------------
// Every pixel of the four extended phase planes is a function of E, the edge-replicated phase 0
// (E(x,y) = phase0[clamp y][clamp x]):
//   phase 1 = horizontal filter of E, except columns x < 0 and x >= w-1: E itself   (schroframe.c:2022-2024)
//   phase 2 = vertical filter of E,   except rows y < 0 and y >= h-1:    E itself   (:2018-2020)
//   phase 3 = horizontal filter of phase 2, except rows y < 0 / y >= h-1: phase 1, and there
//             columns x < 0 / x >= w-1: phase 2                                      (:2026-2028)
// so one kernel serves interior and border tiles alike: the tile of E is loaded with clamped
// coordinates, the filters run on whole 4-pixel words everywhere and the exceptions are a word
// select per row (vertical) or a byte mask per word (horizontal).
//   * 8 taps x 4 pixels = 8 dp4a: a pixel's window is two words of four bytes, u8 data times s8
//     taps (-1,3,-7,21 | 21,-7,3,-1), chained through the accumulator that starts at the rounding 16;
//   * horizontally the eight byte windows of a word come from three funnel shifts on each side;
//   * vertically a thread walks down a word column: the window words G_k(s) = byte k of rows
//     s..s+3 slide by one PRMT per pixel per row, and each is used twice (low taps of row s,
//     high taps of row s-4);
//   * >> 5, then cvt.pack.sat (I2IP) clamps and packs two pixels an instruction.
constexpr int U3_W = 128, U3_H = 56;             // output tile; 34 word columns x 7 strips of 8 rows = 238 of 256 threads
constexpr int U3_RS = 8;                         // rows per vertical strip
constexpr int U3_WORDS = U3_W / 4 + 2;           // tile words incl. one halo word each side
constexpr int U3_PITCH = U3_WORDS + 1;
constexpr unsigned TAPS_LO = 0x15f903ffu;        // bytes (-1, 3, -7, 21)
constexpr unsigned TAPS_HI = 0xff03f915u;        // bytes (21, -7, 3, -1)

__device__ __forceinline__ int dp4a_us (unsigned data, unsigned taps, int acc)
{
  int d;
  asm ("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(data), "r"(taps), "r"(acc));
  return d;
}

// clamp four sums >> SHIFT to bytes and pack them, pixel 0 lowest
template <int SHIFT = 5>
__device__ __forceinline__ unsigned pack4_sat (int a0, int a1, int a2, int a3)
{
  unsigned hi, d;
  asm ("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a3 >> SHIFT), "r"(a2 >> SHIFT), "r"(0));
  asm ("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a1 >> SHIFT), "r"(a0 >> SHIFT), "r"(hi));
  return d;
}

// horizontal filter of pixels b[0..3] given the words holding b[-4..-1], b[0..3], b[4..7]
__device__ __forceinline__ unsigned horiz4_dp4a (unsigned wm, unsigned w0, unsigned wp)
{
  unsigned u[8];                                   // u[j] = bytes b[j-3 .. j]
  u[0] = __funnelshift_r (wm, w0, 8);
  u[1] = __funnelshift_r (wm, w0, 16);
  u[2] = __funnelshift_r (wm, w0, 24);
  u[3] = w0;
  u[4] = __funnelshift_r (w0, wp, 8);
  u[5] = __funnelshift_r (w0, wp, 16);
  u[6] = __funnelshift_r (w0, wp, 24);
  u[7] = wp;
  int acc[4];
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = dp4a_us (u[i + 4], TAPS_HI, dp4a_us (u[i], TAPS_LO, 16));
  return pack4_sat (acc[0], acc[1], acc[2], acc[3]);
}

// bytes of a word starting at column x that lie in [lo, hi]
__device__ __forceinline__ unsigned byte_mask (int x, int lo, int hi)
{
  unsigned m = 0xffffffffu;
  const int cut_lo = lo - x, cut_hi = x + 3 - hi;
  if (cut_lo > 0) m = cut_lo >= 4 ? 0u : m << (8 * cut_lo);
  if (cut_hi > 0) m = cut_hi >= 4 ? 0u : m & (0xffffffffu >> (8 * cut_hi));
  return m;
}

// store the bytes of `v` selected by `m` (a whole word when all four are)
__device__ __forceinline__ void store_masked (uint8_t *p, unsigned v, unsigned m)
{
  if (m == 0xffffffffu) *reinterpret_cast<unsigned *> (p) = v;
  else {
#pragma unroll
    for (int k = 0; k < 4; k++) if ((m >> (8 * k)) & 1) p[k] = (uint8_t) (v >> (8 * k));
  }
}

// EDGE = false: a tile whose every tap, every output and every store lies inside the picture proper
// (nine tiles in ten at 2160p) -- no clamps, no masks, no exception rows, loop bounds known at compile time.
template <bool EDGE>
__device__ __forceinline__ void upsample_tile_words (const FrameArgs &a, unsigned (&s0)[U3_H + 7][U3_PITCH],
    unsigned (&sv)[U3_H][U3_PITCH], uint8_t *p0, int stride, int w, int h, int x0, int y0)
{
  const int ext = a.ext, q = stride >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- E: a warp per row (32 words = one 128-byte line), then the two halo words of every row
  if (!EDGE) {
    const uint8_t *src = p0 + (ptrdiff_t) (y0 - 3 + warp) * stride + x0 + 4 * lane;
#pragma unroll
    for (int ty = warp; ty < U3_H + 7; ty += 8, src += (ptrdiff_t) 8 * stride)
      s0[ty][lane + 1] = __ldg (reinterpret_cast<const unsigned *> (src));
    if (threadIdx.x < 2 * (U3_H + 7)) {
      const int ty = threadIdx.x >> 1, c = (threadIdx.x & 1) * (U3_WORDS - 1);
      s0[ty][c] = __ldg (reinterpret_cast<const unsigned *> (p0 + (ptrdiff_t) (y0 - 3 + ty) * stride + x0 - 4 + 4 * c));
    }
  } else {
    // clamped rows; words that straddle a picture edge are gathered byte by byte
    for (int i = threadIdx.x; i < (U3_H + 7) * U3_WORDS; i += blockDim.x) {
      const int ty = i / U3_WORDS, c = i - ty * U3_WORDS;
      const int yy = clampi (y0 + ty - 3, 0, h - 1), x = x0 - 4 + 4 * c;
      const uint8_t *row = p0 + (ptrdiff_t) yy * stride;
      unsigned v;
      if (x >= 0 && x + 3 <= w - 1) v = __ldg (reinterpret_cast<const unsigned *> (row + x));
      else {
        v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) v |= (unsigned) row[clampi (x + k, 0, w - 1)] << (8 * k);
      }
      s0[ty][c] = v;
    }
  }
  __syncthreads ();

  // ---- phase 2: a thread owns word column c over a strip of U3_RS rows
  if (threadIdx.x < U3_WORDS * (U3_H / U3_RS)) {
    const int strip = threadIdx.x / U3_WORDS, c = threadIdx.x - strip * U3_WORDS;
    const int t0 = strip * U3_RS;
    const int x = x0 - 4 + 4 * c;
    const bool mine = c >= 1 && c <= U3_W / 4;                // halo columns are computed for phase 3, not stored
    const unsigned xm = !mine ? 0u : EDGE ? byte_mask (x, -ext, w + ext - 1) : 0xffffffffu;
    uint8_t *dst = p0 + (ptrdiff_t) (y0 + t0) * stride + 2 * q + x;
    unsigned g[5][4];                              // g[s % 5][k] = byte k of rows s .. s+3
#pragma unroll
    for (int k = 0; k < 4; k++) g[1][k] = 0;               // (slot of s = -4, the first window shifted in)
#pragma unroll
    for (int t = 0; t < U3_RS + 7; t++) {
      // row t0 + t enters the window: slot (t - 3) now holds rows t0+t-3 .. t0+t
      const unsigned r = s0[t0 + t][c];
      const int cur = (t + 2) % 5, prev = (t + 1) % 5;        // slot of s = t - 3 (cur), of s - 1 (prev); t < 3 primes
#pragma unroll
      for (int k = 0; k < 4; k++) g[cur][k] = __byte_perm (g[prev][k], r, 0x0321 | ((4 + k) << 12));
      if (t >= 7) {
        const int ty = t0 + t - 7, y = y0 + ty;              // output row: low taps rows ty..ty+3 (s = t-7), high taps s = t-3
        const int lo = (t - 7 + 5) % 5;                      // slot of s = t - 7
        int acc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = dp4a_us (g[cur][k], TAPS_HI, dp4a_us (g[lo][k], TAPS_LO, 16));
        unsigned v = pack4_sat (acc[0], acc[1], acc[2], acc[3]);
        if (EDGE && (y < 0 || y >= h - 1)) v = s0[ty + 3][c];
        sv[ty][c] = v;
        if (!EDGE) { if (mine) *reinterpret_cast<unsigned *> (dst) = v; }
        else if (xm && y < h + ext) store_masked (dst, v, xm);
        dst += stride;
      }
    }
  }
  __syncthreads ();

  // ---- phases 1 and 3 (and the phase-0 border when asked): a warp per row, a word per lane
  const int c = lane + 1, x = x0 + 4 * lane;
  if (EDGE && x >= w + ext) return;
  const unsigned xm = EDGE ? byte_mask (x, -ext, w + ext - 1) : 0xffffffffu;
  const unsigned fm = EDGE ? byte_mask (x, 0, w - 2) : 0xffffffffu;     // columns that are filtered
  uint8_t *o = p0 + (ptrdiff_t) (y0 + warp) * stride + x;
#pragma unroll
  for (int ty = warp; ty < U3_H; ty += 8, o += (ptrdiff_t) 8 * stride) {
    const int y = y0 + ty;
    if (EDGE && y >= h + ext) break;
    const unsigned e = s0[ty + 3][c], v2 = sv[ty][c];
    unsigned v1 = e, v3;
    if (!EDGE) {
      v1 = horiz4_dp4a (s0[ty + 3][c - 1], e, s0[ty + 3][c + 1]);
      v3 = horiz4_dp4a (sv[ty][c - 1], v2, sv[ty][c + 1]);
      *reinterpret_cast<unsigned *> (o + q) = v1;
      *reinterpret_cast<unsigned *> (o + 3 * q) = v3;
    } else {
      if (fm) {
        const unsigned f = horiz4_dp4a (s0[ty + 3][c - 1], e, s0[ty + 3][c + 1]);
        v1 = (f & fm) | (e & ~fm);
      }
      v3 = v1;
      if (y >= 0 && y < h - 1) {
        v3 = v2;
        if (fm) {
          const unsigned f = horiz4_dp4a (sv[ty][c - 1], v2, sv[ty][c + 1]);
          v3 = (f & fm) | (v2 & ~fm);
        }
      }
      store_masked (o + q, v1, xm);
      store_masked (o + 3 * q, v3, xm);
      if (a.fuse_edge) {
        const unsigned om = (y < 0 || y >= h) ? xm : (xm & ~byte_mask (x, 0, w - 1));     // outside the picture
        if (om) store_masked (o, e, om);
      }
    }
  }
}

__global__ void __launch_bounds__ (256)
upsample_kernel_words (const FrameArgs a)
{
  __shared__ unsigned s0[U3_H + 7][U3_PITCH];      // E: word c = pixels x0-4+4c .., row t = y0-3+t
  __shared__ unsigned sv[U3_H][U3_PITCH];          // phase 2 of the same words
  static_assert (U3_W == 128 && U3_H % 8 == 0, "a warp per row of 32 words, eight rows a round");

  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp], ext = a.ext;
  const int x0 = tp.bx * U3_W - ext, y0 = tp.by * U3_H - ext;
  if (x0 >= w + ext || y0 >= h + ext) return;
  uint8_t *p0 = reinterpret_cast<uint8_t *> (plane_ptr (a.planes, pic, comp));
  const int stride = a.planes.stride[comp];
  const bool interior = x0 - 4 >= 0 && x0 + U3_W + 3 <= w - 2 && y0 - 3 >= 0 && y0 + U3_H - 1 + 4 <= h - 2;
  if (interior) upsample_tile_words<false> (a, s0, sv, p0, stride, w, h, x0, y0);
  else upsample_tile_words<true> (a, s0, sv, p0, stride, w, h, x0, y0);
}

// ---- upsample, any alignment: one pixel at a time ---------------------------------------
#ifndef U2_TW
#define U2_TW 128
#define U2_TH 64
#endif
constexpr int U2_W = U2_TW, U2_H = U2_TH;        // output tile
constexpr int U2_WORDS = U2_W / 4 + 2;
constexpr int U2_PITCH = U2_WORDS + 1;

__global__ void __launch_bounds__ (256)
upsample_kernel_pixel (const FrameArgs a)
{
  __shared__ unsigned s0[U2_H + 7][U2_PITCH];
  __shared__ unsigned sv[U2_H][U2_PITCH];

  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp], ext = a.ext;
  const int x0 = tp.bx * U2_W - ext, y0 = tp.by * U2_H - ext;
  if (x0 >= w + ext || y0 >= h + ext) return;
  uint8_t *p0 = reinterpret_cast<uint8_t *> (plane_ptr (a.planes, pic, comp));
  const int stride = a.planes.stride[comp];
  const int q = stride >> 2;

  // per-pixel rules over the tile.
  // phase 1 (schroframe.c:2022-2024): horizontal filter of phase 0 row clamp(y); left border =
  //   phase 0 column 0, columns >= w-1 = phase 0 column w-1
  // phase 2 (:2018-2020): vertical filter; rows above = phase 0 row 0, last row and below =
  //   phase 0 row h-1, side borders replicate phase 2 itself
  // phase 3 (:2026-2028): horizontal filter of phase 2; rows above / last row and below are
  //   copies of phase 1, side borders come from phase 2
  uint8_t *b0 = reinterpret_cast<uint8_t *> (&s0[0][0]);
  uint8_t *bv = reinterpret_cast<uint8_t *> (&sv[0][0]);
  constexpr int BW = U2_W + 8;                     // byte tile width: x0-3 .. x0+U2_W+4
  for (int i = threadIdx.x; i < (U2_H + 7) * BW; i += blockDim.x) {
    const int ty = i / BW, tx = i - ty * BW;
    const int yy = clampi (y0 + ty - 3, 0, h - 1), xx = clampi (x0 + tx - 3, 0, w - 1);
    b0[ty * BW + tx] = p0[(ptrdiff_t) yy * stride + xx];
  }
  __syncthreads ();
  for (int i = threadIdx.x; i < U2_H * BW; i += blockDim.x) {
    const int ty = i / BW, tx = i - ty * BW;
    const int y = y0 + ty;
    int v;
    if (y >= h - 1 || y < 0) v = b0[(ty + 3) * BW + tx];
    else v = taps8 (&b0[ty * BW + tx], BW);
    bv[ty * BW + tx] = (uint8_t) v;
  }
  __syncthreads ();
  for (int i = threadIdx.x; i < U2_H * U2_W; i += blockDim.x) {
    const int ty = i / U2_W, tx = i - ty * U2_W;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w + ext || y >= h + ext) continue;
    int v1;
    if (x < 0 || x >= w - 1) v1 = b0[(ty + 3) * BW + tx + 3];
    else v1 = taps8 (&b0[(ty + 3) * BW + tx], 1);
    const int v2 = bv[ty * BW + tx + 3];
    int v3;
    if (y < 0 || y >= h - 1) v3 = v1;
    else if (x < 0 || x >= w - 1) v3 = v2;
    else v3 = taps8 (&bv[ty * BW + tx], 1);
    uint8_t *o = p0 + (ptrdiff_t) y * stride + x;
    if (a.fuse_edge && (x < 0 || x >= w || y < 0 || y >= h)) o[0] = b0[(ty + 3) * BW + tx + 3];
    o[q] = (uint8_t) v1;
    o[2 * q] = (uint8_t) v2;
    o[3 * q] = (uint8_t) v3;
  }
}

// ---- downsample ---------------------------------------------------------------------
struct DownArgs {
  TileGrid tiles;
  PlaneSet src, dst;
  int sw[SB2_MAX_COMPONENTS], sh[SB2_MAX_COMPONENTS];
  int dw[SB2_MAX_COMPONENTS], dh[SB2_MAX_COMPONENTS];
  int ncomp;
  int dst_ext;                  // > 0: also replicate the result into dst's border
};

// ---- downsample, the usual case: whole words ----------------------------------------------
// Vertical (6,26,26,6)+32 >> 6 on source words (two pixels per 16-bit lane pair), result in shared
// memory; horizontal: an output pixel's four taps are one unaligned word of that intermediate --
// a funnel shift and one dp4a.  Clamped loads make border tiles the same computation as interior
// ones (the intermediate of a clamped column is the clamped column of the intermediate); only their
// stores differ (partial words, the fused border replication).
#ifndef D2_TW
#define D2_TW 64
#define D2_TH 32
#endif
constexpr int D2_W = D2_TW, D2_H = D2_TH;                // output tile
constexpr int D2_WORDS = (2 * D2_W) / 4 + 3;       // intermediate words: source columns 2*x0-4 .. 2*x0+2*D2_W+8
constexpr int D2_BW = 2 * D2_W + 2;                // byte tile of the per-pixel kernel
constexpr unsigned DTAPS = 0x061a1a06u;            // bytes (6, 26, 26, 6)
static_assert (D2_W == 64 && D2_H == 32 && D2_WORDS == 35, "thread maps of downsample_tile_words");

template <bool EDGE>
__device__ __forceinline__ void downsample_tile_words (const DownArgs &a, unsigned (&st)[D2_H][D2_WORDS + 1],
    uint8_t *so, const uint8_t *s, uint8_t *d, int ss, int dstr, int sw, int sh, int dw, int dh, int x0, int y0)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- vertical: word c of intermediate row ty from source rows 2y-1 .. 2y+2
  auto vert_word = [&] (int ty, int c) {
    unsigned r[4];
    const int x = 2 * x0 - 4 + 4 * c, yb = 2 * (y0 + ty) - 1;
    if (!EDGE) {
      const uint8_t *col = s + (ptrdiff_t) yb * ss + x;
#pragma unroll
      for (int j = 0; j < 4; j++) r[j] = __ldg (reinterpret_cast<const unsigned *> (col + (ptrdiff_t) j * ss));
    } else {
      const bool whole = x >= 0 && x + 3 <= sw - 1;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint8_t *row = s + (ptrdiff_t) clampi (yb + j, 0, sh - 1) * ss;
        if (whole) r[j] = __ldg (reinterpret_cast<const unsigned *> (row + x));
        else {
          r[j] = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) r[j] |= (unsigned) row[clampi (x + k, 0, sw - 1)] << (8 * k);
        }
      }
    }
    unsigned lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      lo[j] = __byte_perm (r[j], 0, 0x4140);
      hi[j] = __byte_perm (r[j], 0, 0x4342);
    }
    const unsigned vlo = ((6u * (lo[0] + lo[3]) + 26u * (lo[1] + lo[2]) + 0x00200020u) >> 6) & 0x00ff00ffu;
    const unsigned vhi = ((6u * (hi[0] + hi[3]) + 26u * (hi[1] + hi[2]) + 0x00200020u) >> 6) & 0x00ff00ffu;
    st[ty][c] = __byte_perm (vlo, vhi, 0x6420);
  };
#pragma unroll
  for (int ty = warp; ty < D2_H; ty += 8) vert_word (ty, lane);
  if (threadIdx.x < 3 * D2_H) vert_word (threadIdx.x / 3, 32 + threadIdx.x % 3);
  __syncthreads ();

  // ---- horizontal: output pixels x..x+3 use b[-1..8], b[k] = intermediate[2x+k]
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const int ty = (threadIdx.x >> 4) + 16 * it, g = threadIdx.x & 15;
    const unsigned *t = &st[ty][2 * g];            // t[0] = b[-4..-1], t[1] = b[0..3], t[2] = b[4..7], t[3] = b[8..11]
    const unsigned t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3];
    const int o0 = (int) __dp4a (__funnelshift_r (t0, t1, 24), DTAPS, 32u);    // b[-1..2]
    const int o1 = (int) __dp4a (__funnelshift_r (t1, t2, 8), DTAPS, 32u);     // b[1..4]
    const int o2 = (int) __dp4a (__funnelshift_r (t1, t2, 24), DTAPS, 32u);    // b[3..6]
    const int o3 = (int) __dp4a (__funnelshift_r (t2, t3, 8), DTAPS, 32u);     // b[5..8]
    const unsigned v = pack4_sat<6> (o0, o1, o2, o3);
    if (!EDGE) *reinterpret_cast<unsigned *> (d + (ptrdiff_t) (y0 + ty) * dstr + x0 + 4 * g) = v;
    else reinterpret_cast<unsigned *> (so)[ty * (D2_W / 4) + g] = v;
  }
  if (!EDGE) return;
  __syncthreads ();
  // ---- border tiles: the words inside the plane ...
  for (int i = threadIdx.x; i < D2_H * (D2_W / 4); i += blockDim.x) {
    const int ty = i / (D2_W / 4), g = i - ty * (D2_W / 4);
    const int x = x0 + 4 * g, y = y0 + ty;
    if (x >= dw || y >= dh) continue;
    const unsigned v = reinterpret_cast<const unsigned *> (so)[i];
    uint8_t *o = d + (ptrdiff_t) y * dstr + x;
    if (x + 3 < dw) *reinterpret_cast<unsigned *> (o) = v;
    else for (int k = 0; x + k < dw; k++) o[k] = (uint8_t) (v >> (8 * k));
  }
  // ... and the fused edge extension (schro_frame_mc_edgeextend on the result): every border position
  // whose nearest plane pixel belongs to this tile, a warp per row of the tile grown by the extension
  const int e = a.dst_ext;
  if (e <= 0) return;
  for (int ry = warp; ry < D2_H + 2 * e; ry += 8) {
    const int y = y0 - e + ry;
    if (y >= dh + e) break;
    const int cy = clampi (y, 0, dh - 1);
    if (cy < y0 || cy >= y0 + D2_H) continue;
    const bool row_inside = y >= 0 && y < dh;
    for (int rx = lane; rx < D2_W + 2 * e; rx += 32) {
      const int x = x0 - e + rx;
      if (x >= dw + e) break;
      if (row_inside && x >= 0 && x < dw) continue;            // a plane pixel: stored above
      const int cx = clampi (x, 0, dw - 1);
      if (cx < x0 || cx >= x0 + D2_W) continue;                // another tile's pixel
      d[(ptrdiff_t) y * dstr + x] = so[(cy - y0) * D2_W + cx - x0];
    }
  }
}

__global__ void __launch_bounds__ (256)
downsample_kernel_words (const DownArgs a)
{
  __shared__ unsigned st[D2_H][D2_WORDS + 1];      // vertically filtered source words
  __shared__ __align__ (4) uint8_t so[D2_H * D2_W];   // border tiles: the tile's output before the store pass

  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int sw = a.sw[comp], sh = a.sh[comp], dw = a.dw[comp], dh = a.dh[comp];
  const int x0 = tp.bx * D2_W, y0 = tp.by * D2_H;
  if (x0 >= dw || y0 >= dh) return;
  const uint8_t *s = reinterpret_cast<const uint8_t *> (plane_ptr (a.src, pic, comp));
  uint8_t *d = reinterpret_cast<uint8_t *> (plane_ptr (a.dst, pic, comp));
  const int ss = a.src.stride[comp], dstr = a.dst.stride[comp];
  // interior: every tap inside the source, every output inside the destination and away from its edges
  const bool interior = 2 * x0 - 4 >= 0 && 2 * x0 + 2 * D2_W + 8 <= sw &&
      2 * y0 - 1 >= 0 && 2 * (y0 + D2_H - 1) + 2 <= sh - 1 && x0 + D2_W < dw && y0 + D2_H < dh && x0 > 0 && y0 > 0;
  if (interior) downsample_tile_words<false> (a, st, so, s, d, ss, dstr, sw, sh, dw, dh, x0, y0);
  else downsample_tile_words<true> (a, st, so, s, d, ss, dstr, sw, sh, dw, dh, x0, y0);
}

// ---- downsample, any alignment: one pixel at a time ----------------------------------------
__global__ void __launch_bounds__ (256)
downsample_kernel_pixel (const DownArgs a)
{
  __shared__ unsigned st[D2_H][D2_WORDS + 1];

  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int sw = a.sw[comp], sh = a.sh[comp], dw = a.dw[comp], dh = a.dh[comp];
  const int x0 = tp.bx * D2_W, y0 = tp.by * D2_H;
  if (x0 >= dw || y0 >= dh) return;
  const uint8_t *s = reinterpret_cast<const uint8_t *> (plane_ptr (a.src, pic, comp));
  uint8_t *d = reinterpret_cast<uint8_t *> (plane_ptr (a.dst, pic, comp));
  const int ss = a.src.stride[comp], dstr = a.dst.stride[comp];

  // per-pixel path with clamped indices and the fused border replication
  uint8_t *sm = reinterpret_cast<uint8_t *> (&st[0][0]);
  static_assert (D2_H * D2_BW <= D2_H * (D2_WORDS + 1) * 4, "byte tile fits");
  for (int i = threadIdx.x; i < D2_H * D2_BW; i += blockDim.x) {
    const int ty = i / D2_BW, tx = i - ty * D2_BW;
    const int y = y0 + ty;
    const int xx = clampi (2 * x0 - 1 + tx, 0, sw - 1);
    const int r0 = s[(ptrdiff_t) clampi (2 * y - 1, 0, sh - 1) * ss + xx];
    const int r1 = s[(ptrdiff_t) clampi (2 * y, 0, sh - 1) * ss + xx];
    const int r2 = s[(ptrdiff_t) clampi (2 * y + 1, 0, sh - 1) * ss + xx];
    const int r3 = s[(ptrdiff_t) clampi (2 * y + 2, 0, sh - 1) * ss + xx];
    sm[ty * D2_BW + tx] = (uint8_t) ((6 * (r0 + r3) + 26 * (r1 + r2) + 32) >> 6);
  }
  __syncthreads ();
  for (int i = threadIdx.x; i < D2_H * D2_W; i += blockDim.x) {
    const int ty = i / D2_W, tx = i - ty * D2_W;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= dw || y >= dh) continue;
    const uint8_t *t = &sm[ty * D2_BW + 2 * tx];
    const uint8_t v = (uint8_t) clamp255 ((6 * ((int) t[0] + t[3]) + 26 * ((int) t[1] + t[2]) + 32) >> 6);
    d[(ptrdiff_t) y * dstr + x] = v;
    if (a.dst_ext > 0 && (x == 0 || x == dw - 1 || y == 0 || y == dh - 1)) {
      const int e = a.dst_ext;
      const int qx0 = x == 0 ? -e : 0, qx1 = x == dw - 1 ? e : 0;
      const int qy0 = y == 0 ? -e : 0, qy1 = y == dh - 1 ? e : 0;
      for (int qy = qy0; qy <= qy1; qy++)
        for (int qx = qx0; qx <= qx1; qx++)
          if (qx || qy) d[(ptrdiff_t) (y + qy) * dstr + x + qx] = v;
    }
  }
}

// ---- standalone one-direction half-pel filters (schro_frame_upsample_horiz / _vert) ----
__global__ void __launch_bounds__ (256)
upsample_1d_kernel (uint8_t *dst, int ds, const uint8_t *src, int ss, int w, int h, int vertical)
{
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  int acc = 16;
  int v;
  if (vertical) {
    // schroframe.c:1612-1645: rows 0..h-2 filtered, last row copied
    if (y == h - 1) v = src[(ptrdiff_t) y * ss + x];
    else {
      const int t[8] = { -1, 3, -7, 21, 21, -7, 3, -1 };
#pragma unroll
      for (int j = 0; j < 8; j++) acc += t[j] * src[(ptrdiff_t) clampi (y + j - 3, 0, h - 1) * ss + x];
      v = clamp255 (acc >> 5);
    }
  } else {
    // schroframe.c:1515-1555: last column copied only when the row is longer than 8
    if (x == w - 1 && w > 8) v = src[(ptrdiff_t) y * ss + x];
    else {
      const int t[8] = { -1, 3, -7, 21, 21, -7, 3, -1 };
#pragma unroll
      for (int j = 0; j < 8; j++) acc += t[j] * src[(ptrdiff_t) y * ss + clampi (x + j - 3, 0, w - 1)];
      v = clamp255 (acc >> 5);
    }
  }
  dst[(ptrdiff_t) y * ds + x] = (uint8_t) v;
}

static int check_frame_slab (const sb2_slab *s, const char *who)
{
  if (!s || !s->base) return set_error (SB2_ERR_ARG, "%s: null slab", who);
  if (s->ncomp < 1 || s->ncomp > SB2_MAX_COMPONENTS || s->count < 1)
    return set_error (SB2_ERR_ARG, "%s: bad slab (ncomp %d, count %d)", who, s->ncomp, s->count);
  for (int c = 0; c < s->ncomp; c++)
    if (s->width[c] <= 0 || s->height[c] <= 0 || s->stride[c] <= 0)
      return set_error (SB2_ERR_ARG, "%s: bad component %d", who, c);
  return SB2_OK;
}

static FrameArgs frame_args (const sb2_slab *s, int ext)
{
  FrameArgs a;
  a.planes = planeset_from_slab (s);
  a.ncomp = s->ncomp;
  a.ext = ext;
  a.fuse_edge = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    a.w[c] = c < s->ncomp ? s->width[c] : 0;
    a.h[c] = c < s->ncomp ? s->height[c] : 0;
  }
  return a;
}

}  // namespace sb2

using namespace sb2;

extern "C" int
sb2_mc_edgeextend (const sb2_slab *frames, int extension, int phase, void *stream)
{
  int rc = check_frame_slab (frames, "sb2_mc_edgeextend");
  if (rc) return rc;
  if (extension <= 0) return SB2_OK;
  FrameArgs a = frame_args (frames, extension);
  if (phase) {
    for (int c = 0; c < frames->ncomp; c++) a.planes.off[c] += (size_t) (frames->stride[c] >> 2) * phase;
  }
  int maxn = 0;
  for (int c = 0; c < frames->ncomp; c++)
    maxn = max (maxn, 2 * extension * (frames->height[c] + 2 * extension) + 2 * extension * frames->width[c]);
  // items are words on the usual (aligned) path; both paths are grid-stride loops
  dim3 grid (min (ceil_div (maxn / 8 + 1, 256), 1024), 1, frames->ncomp * frames->count);
  double bytes = 0;
  for (int c = 0; c < frames->ncomp; c++)
    bytes += 2.0 * (2 * extension * (frames->height[c] + 2 * extension) + 2 * extension * frames->width[c]) * frames->count;
  {
    LaunchScope scope ("mc_edgeextend", bytes, as_stream (stream));
    edgeextend_kernel<<<grid, 256, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "edgeextend_kernel launch");
}

static int upsample_impl (const sb2_slab *frames, int extension, int fuse_edge, void *stream);

// 0: by alignment; 1 / 2: force the word (dp4a) / per-pixel kernel; SB2_UPSAMPLE_KERNEL sets the initial value
static int g_upsample_variant = -1;
static thread_local int g_upsample_last = 0;
extern "C" void sb2_upsample_force_kernel (int which) { g_upsample_variant = which < 0 || which > 2 ? 0 : which; }
extern "C" int sb2_upsample_last_kernel (void) { return g_upsample_last; }
static int upsample_forced ()
{
  if (g_upsample_variant < 0) {
    const char *v = getenv ("SB2_UPSAMPLE_KERNEL");
    g_upsample_variant = v ? atoi (v) : 0;
    if (g_upsample_variant < 0 || g_upsample_variant > 2) g_upsample_variant = 0;
  }
  return g_upsample_variant;
}

extern "C" int
sb2_upsample (const sb2_slab *frames, int extension, void *stream)
{
  return upsample_impl (frames, extension, 0, stream);
}

extern "C" int
sb2_edgeextend_upsample (const sb2_slab *frames, int extension, void *stream)
{
  return upsample_impl (frames, extension, 1, stream);
}

static int
upsample_impl (const sb2_slab *frames, int extension, int fuse_edge, void *stream)
{
  int rc = check_frame_slab (frames, "sb2_upsample");
  if (rc) return rc;
  for (int c = 0; c < frames->ncomp; c++)
    if (frames->stride[c] % 4)
      return set_error (SB2_ERR_ARG, "sb2_upsample: stride %d of component %d is not 4-phase", frames->stride[c], c);
  FrameArgs a = frame_args (frames, extension);
  a.fuse_edge = fuse_edge;
  int maxw = 0, maxh = 0;
  double bytes = 0;
  for (int c = 0; c < frames->ncomp; c++) {
    maxw = max (maxw, frames->width[c] + 2 * extension);
    maxh = max (maxh, frames->height[c] + 2 * extension);
    // algorithmic bytes: read phase 0 once, write three phases (with their borders)
    bytes += ((double) frames->width[c] * frames->height[c]
        + 3.0 * (frames->width[c] + 2 * extension) * (frames->height[c] + 2 * extension)) * frames->count;
  }
  static_assert ((U2_H + 7) * (U2_W + 8) <= (U2_H + 7) * U2_PITCH * 4 && U2_H * (U2_W + 8) <= U2_H * U2_PITCH * 4,
      "byte tiles of the pixel kernel fit the word tiles");
  (void) maxw; (void) maxh;
  int ew[SB2_MAX_COMPONENTS], eh[SB2_MAX_COMPONENTS];
  for (int c = 0; c < frames->ncomp; c++) { ew[c] = frames->width[c] + 2 * extension; eh[c] = frames->height[c] + 2 * extension; }
  if (frames->count > 65535) return set_error (SB2_ERR_ARG, "sb2_upsample: at most 65535 pictures per call");
  // the word kernel needs every row of every phase plane 4-byte aligned, and tiles that start on a word
  bool words = (extension & 3) == 0 && (((size_t) frames->base | frames->picture_pitch) & 3) == 0;
  for (int c = 0; c < frames->ncomp; c++)
    if ((frames->offset[c] | (size_t) frames->stride[c] | (size_t) (frames->stride[c] >> 2)) & 3) words = false;
  const int force = upsample_forced ();
  if (force == 1 && !words) return set_error (SB2_ERR_UNSUPPORTED, "sb2_upsample: the word kernel needs 4-byte aligned phase rows");
  if (force == 2) words = false;
  g_upsample_last = words ? 1 : 2;
  if (words) {
    const dim3 grid = make_tile_grid (a.tiles, frames->ncomp, ew, eh, U3_W, U3_H, frames->count);
    LaunchScope scope ("upsample", bytes, as_stream (stream));
    upsample_kernel_words<<<grid, 256, 0, as_stream (stream)>>> (a);
  } else {
    const dim3 grid = make_tile_grid (a.tiles, frames->ncomp, ew, eh, U2_W, U2_H, frames->count);
    LaunchScope scope ("upsample_pixel", bytes, as_stream (stream));
    upsample_kernel_pixel<<<grid, 256, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "upsample_kernel launch");
}

static int downsample_impl (const sb2_slab *src, const sb2_slab *dst, int dst_ext, void *stream);

// 0: by alignment; 1 / 2: force the word / per-pixel kernel; SB2_DOWNSAMPLE_KERNEL sets the initial value
static int g_downsample_variant = -1;
static thread_local int g_downsample_last = 0;
extern "C" void sb2_downsample_force_kernel (int which) { g_downsample_variant = which < 0 || which > 2 ? 0 : which; }
extern "C" int sb2_downsample_last_kernel (void) { return g_downsample_last; }
static int downsample_forced ()
{
  if (g_downsample_variant < 0) {
    const char *v = getenv ("SB2_DOWNSAMPLE_KERNEL");
    g_downsample_variant = v ? atoi (v) : 0;
    if (g_downsample_variant < 0 || g_downsample_variant > 2) g_downsample_variant = 0;
  }
  return g_downsample_variant;
}

extern "C" int
sb2_downsample (const sb2_slab *src, const sb2_slab *dst, void *stream)
{
  return downsample_impl (src, dst, 0, stream);
}

extern "C" int
sb2_downsample_edgeextend (const sb2_slab *src, const sb2_slab *dst, int dst_extension, void *stream)
{
  return downsample_impl (src, dst, dst_extension, stream);
}

static int
downsample_impl (const sb2_slab *src, const sb2_slab *dst, int dst_ext, void *stream)
{
  int rc = check_frame_slab (src, "sb2_downsample(src)");
  if (rc) return rc;
  rc = check_frame_slab (dst, "sb2_downsample(dst)");
  if (rc) return rc;
  if (src->ncomp != dst->ncomp || src->count != dst->count)
    return set_error (SB2_ERR_ARG, "sb2_downsample: slab shapes differ");
  DownArgs a;
  a.src = planeset_from_slab (src);
  a.dst = planeset_from_slab (dst);
  a.ncomp = src->ncomp;
  a.dst_ext = dst_ext;
  int maxw = 0, maxh = 0;
  double bytes = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    a.sw[c] = c < src->ncomp ? src->width[c] : 0;
    a.sh[c] = c < src->ncomp ? src->height[c] : 0;
    a.dw[c] = c < src->ncomp ? dst->width[c] : 0;
    a.dh[c] = c < src->ncomp ? dst->height[c] : 0;
    if (c < src->ncomp) {
      if (dst->width[c] != (src->width[c] + 1) / 2 || dst->height[c] != (src->height[c] + 1) / 2)
        return set_error (SB2_ERR_ARG, "sb2_downsample: component %d: %dx%d is not half of %dx%d", c,
            dst->width[c], dst->height[c], src->width[c], src->height[c]);
      maxw = max (maxw, dst->width[c]);
      maxh = max (maxh, dst->height[c]);
      bytes += ((double) src->width[c] * src->height[c] + (double) dst->width[c] * dst->height[c]) * src->count;
    }
  }
  (void) maxw; (void) maxh;
  if (src->count > 65535) return set_error (SB2_ERR_ARG, "sb2_downsample: at most 65535 pictures per call");
  const dim3 grid = make_tile_grid (a.tiles, src->ncomp, a.dw, a.dh, D2_W, D2_H, src->count);
  // the word kernel needs 4-byte aligned rows on both sides
  bool words = (((size_t) src->base | src->picture_pitch | (size_t) dst->base | dst->picture_pitch) & 3) == 0;
  for (int c = 0; c < src->ncomp; c++)
    if ((src->offset[c] | (size_t) src->stride[c] | dst->offset[c] | (size_t) dst->stride[c]) & 3) words = false;
  const int force = downsample_forced ();
  if (force == 1 && !words) return set_error (SB2_ERR_UNSUPPORTED, "sb2_downsample: the word kernel needs 4-byte aligned rows");
  if (force == 2) words = false;
  g_downsample_last = words ? 1 : 2;
  {
    LaunchScope scope (words ? "downsample" : "downsample_pixel", bytes, as_stream (stream));
    if (words) downsample_kernel_words<<<grid, 256, 0, as_stream (stream)>>> (a);
    else downsample_kernel_pixel<<<grid, 256, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "downsample_kernel launch");
}

extern "C" int
sb2_upsample_plane_1d (uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride, int width,
    int height, int vertical, void *stream)
{
  if (!dst || !src || width < 1 || height < 1) return set_error (SB2_ERR_ARG, "sb2_upsample_plane_1d: bad argument");
  dim3 grid (ceil_div (width, 32), ceil_div (height, 8));
  {
    LaunchScope scope (vertical ? "upsample_vert" : "upsample_horiz", 2.0 * width * height, as_stream (stream));
    upsample_1d_kernel<<<grid, 256, 0, as_stream (stream)>>> (dst, dst_stride, src, src_stride, width, height, vertical);
  }
  return check_cuda (cudaGetLastError (), "upsample_1d_kernel launch");
}
