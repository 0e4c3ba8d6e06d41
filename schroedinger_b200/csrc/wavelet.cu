// wavelet.cu -- Dirac integer lifting wavelets for sm_100a.
//
// Replaces, bit-exactly, the reference's one-level 2-D transforms
//   schro_wavelet_transform_2d          (schroedinger/schrowaveletorc.c:60-117)
//   schro_wavelet_inverse_transform_2d  (schroedinger/schrowaveletorc.c:121-188)
// for all seven filters x {s16, s32} and the multi-level drivers around them
// (schroedinger/schroframe.c:1192-1228, schroedinger/schrodecoder.c:1809-1853).
//
// Design (see DESIGN.md "wavelets"):
//  * one CTA = one output tile of one level of one component of one picture;
//    horizontal and vertical lifting are fused in shared memory so a level is a
//    single read + single write of its sub-plane;
//  * the arithmetic is plain lifting on polyphase arrays E[k]=x[2k], O[k]=x[2k+1]
//    with taps clamped to [0,n-1]; the 16-bit variants reproduce Orc's wrap
//    points (addw/subw/convlw wrap, mulswl/avgsw are wide) exactly;
//  * levels never run in place: level l reads LL from a compact scratch plane and
//    the three detail bands from the source plane, so tiles have no halo hazards
//    and the only extra traffic is the (1/4 + 1/16 + ...) LL ping-pong.
//
// No tensor cores: nothing here is a contraction.  The kernel is HBM/LSU bound.

#include "common.cuh"
#include <cstdio>

namespace sb2 {

enum { K_A22, K_A11, K_M4, K_M2, K_F8A, K_F8B, K_COPY, K_HALF };

struct Step {
  int target;   // 0: update E from O, 1: update O from E
  int kind;
  int sign;     // forward direction
  int tap0;     // offset of first tap in the source polyphase array
  int p1, p2, p3;
};

// Forward lifting steps (inverse = reverse order, flipped sign).
// Sources: schroedinger/schrowaveletorc.c:287-301 (DD9/7), :376-390 (LeGall),
// :459-473 (DD13/7), :545-567 (Haar), :606-648 (Fidelity), :729-753 (Daub 9/7).
__host__ __device__ constexpr Step step_of (int f, int s)
{
  switch (f) {
    case 0: return s == 0 ? Step{1, K_M4, -1, -1, 8, 4, 0} : Step{0, K_A22, +1, -1, 0, 0, 0};
    case 1: return s == 0 ? Step{1, K_A11, -1, 0, 0, 0, 0} : Step{0, K_A22, +1, -1, 0, 0, 0};
    case 2: return s == 0 ? Step{1, K_M4, -1, -1, 8, 4, 0} : Step{0, K_M4, +1, -2, 16, 5, 0};
    case 3:
    case 4: return s == 0 ? Step{1, K_COPY, -1, 0, 0, 0, 0} : Step{0, K_HALF, +1, 0, 0, 0, 0};
    case 5: return s == 0 ? Step{0, K_F8A, +1, -4, 128, 0, 0} : Step{1, K_F8B, +1, -3, 127, 0, 0};
    default:
      return s == 0 ? Step{1, K_M2, -1, 0, 6497, 2048, 12}
           : s == 1 ? Step{0, K_M2, -1, -1, 217, 2048, 12}
           : s == 2 ? Step{1, K_M2, +1, 0, 3616, 2048, 12}
                    : Step{0, K_M2, +1, -1, 1817, 2048, 12};
  }
}
__host__ __device__ constexpr int num_steps (int f) { return f == 6 ? 4 : 2; }
// pre/post shift (schroedinger/schroorc.orc:770-807)
__host__ __device__ constexpr int filter_shift (int f) { return (f == 3 || f == 5) ? 0 : 1; }
// halo in polyphase samples = sum of the per-step reaches
__host__ __device__ constexpr int filter_halo (int f)
{
  return f == 0 ? 3 : f == 1 ? 2 : f == 2 ? 4 : f == 5 ? 8 : f == 6 ? 4 : 0;
}
__host__ __device__ constexpr int kind_taps (int kind)
{
  return (kind == K_M4) ? 4 : (kind == K_F8A || kind == K_F8B) ? 8
       : (kind == K_COPY || kind == K_HALF) ? 1 : 2;
}

// ---- exact arithmetic ------------------------------------------------------
template <typename T> struct Ar;
template <> struct Ar<int16_t> {
  // storage wraps at 16 bit; products / mas accumulators are 32 bit wide
  static __device__ __forceinline__ int wr (int x) { return (int) (short) x; }
  static __device__ __forceinline__ int add (int a, int b) { return wr (a + b); }
  static __device__ __forceinline__ int sub (int a, int b) { return wr (a - b); }
  static __device__ __forceinline__ int wadd (int a, int b) { return a + b; }
  static __device__ __forceinline__ int wsub (int a, int b) { return a - b; }
  static __device__ __forceinline__ int wmul (int a, int b) { return a * b; }
  static __device__ __forceinline__ int avg (int a, int b) { return (a + b + 1) >> 1; }
};
template <> struct Ar<int32_t> {
  // everything wraps at 32 bit; avgsl is exact (64-bit in Orc's C emulation)
  static __device__ __forceinline__ int wr (int x) { return x; }
  static __device__ __forceinline__ int add (int a, int b) { return (int) ((unsigned) a + (unsigned) b); }
  static __device__ __forceinline__ int sub (int a, int b) { return (int) ((unsigned) a - (unsigned) b); }
  static __device__ __forceinline__ int wadd (int a, int b) { return add (a, b); }
  static __device__ __forceinline__ int wsub (int a, int b) { return sub (a, b); }
  static __device__ __forceinline__ int wmul (int a, int b) { return (int) ((unsigned) a * (unsigned) b); }
  static __device__ __forceinline__ int avg (int a, int b) { return __rhadd (a, b); }
};

// term of one lifting step from its taps v[0..ntaps)
template <typename T, int KIND>
__device__ __forceinline__ int lift_term (const int *v, int p1, int p2, int p3)
{
  typedef Ar<T> A;
  if (KIND == K_A22) {            // schroorc.orc:4-73
    int t = A::add (v[0], v[1]);
    t = A::add (t, 2);
    return t >> 2;
  } else if (KIND == K_A11) {     // avgsw / avgsl, schroorc.orc:76-134
    return A::avg (v[0], v[1]);
  } else if (KIND == K_M4) {      // mas4 1-9-9-1, schroorc.orc:295-411
    int t = A::add (v[1], v[2]);
    int u = A::add (v[0], v[3]);
    int acc = A::wadd (A::wsub (A::wmul (t, 9), u), p1);
    return acc >> p2;
  } else if (KIND == K_M2) {      // mas2, schroorc.orc:221-292
    int t = A::add (v[0], v[1]);
    int acc = A::wadd (A::wmul (t, p1), p2);
    return acc >> p3;
  } else if (KIND == K_F8A || KIND == K_F8B) {   // plain-C mas8, schrowaveletorc.c:606-663
    int acc = p1;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int w = (KIND == K_F8A)
          ? (j == 0 || j == 7 ? -8 : j == 1 || j == 6 ? 21 : j == 2 || j == 5 ? -46 : 161)
          : (j == 0 || j == 7 ? 2 : j == 1 || j == 6 ? -10 : j == 2 || j == 5 ? 25 : -81);
      acc = A::wadd (acc, A::wmul (v[j], w));
    }
    return acc >> 8;
  } else if (KIND == K_COPY) {
    return v[0];
  } else {                        // K_HALF: avgsw(x, 0)
    return A::avg (v[0], 0);
  }
}

// ---- kernel arguments ------------------------------------------------------
struct LevelArgs {
  PlaneSet dense;   // forward: input X;  inverse: output Y      (w x h, interleaved)
  PlaneSet bands;   // the coefficient plane at this level's stride ([L|H] split rows)
  PlaneSet ll;      // LL band (w/2 x h/2): forward output / inverse input
  int w[SB2_MAX_COMPONENTS];
  int h[SB2_MAX_COMPONENTS];
  int ncomp;
  int comp_map[SB2_MAX_COMPONENTS];   // launch slot -> component (a launch may cover a subset)
  // compact grid: blockIdx.x runs over the tiles of every selected component (no CTA is launched
  // just to find itself outside a smaller chroma plane), blockIdx.y = picture
  int tile_start[SB2_MAX_COMPONENTS + 1];
  int tiles_x[SB2_MAX_COMPONENTS];
  int ncomp_total;                    // components per picture in the plane sets
  // fused combine (inverse level 0 only, OUT8 kernels): `dense` is a u8 plane set of out_w x out_h pixels (the
  // picture: a crop of the transform's area); every sample is shifted right by out_shift with rounding
  // (schro_frame_shift_right) and converted to 8 bits (schro_frame_convert) on its way out
  int out_w[SB2_MAX_COMPONENTS], out_h[SB2_MAX_COMPONENTS];
  int out_shift;
};

// one sample through schro_frame_shift_right (schroedinger/schroframe.c:1265-1291: add wraps at the sample
// width) and the s16 / s32 -> u8 converters the library runs (schroorc-dist.c: orc_offsetconvert_u8_s16 /
// _u8_s32, glue.cu convert_sample) -- the non-reference intra branch of schro_decoder_x_combine
// (schroedinger/schrodecoder.c:2054-2061)
template <typename T>
__device__ __forceinline__ unsigned combine_to_u8 (int v, int shift)
{
  if (shift) {
    const int rnd = (1 << shift) >> 1;
    v = sizeof (T) == 2 ? ((int) (short) (v + rnd)) >> shift : ((int) ((unsigned) v + (unsigned) rnd)) >> shift;
  }
  int t;
  if (sizeof (T) == 2) t = (int) (short) (v + 128);
  else t = (int) (short) min (max ((int) ((unsigned) v + 128u), 0), 65535);
  return (unsigned) min (max (t, 0), 255);
}

struct TileId { int comp, bx, by; };
__device__ __forceinline__ TileId level_tile (const LevelArgs &a)
{
  int t = blockIdx.x, ci = 0;
#pragma unroll
  for (int k = 1; k < SB2_MAX_COMPONENTS; k++)
    if (k < a.ncomp && t >= a.tile_start[k]) ci = k;
  t -= a.tile_start[ci];
  TileId id;
  id.comp = a.comp_map[ci];
  id.by = t / a.tiles_x[ci];
  id.bx = t - id.by * a.tiles_x[ci];
  return id;
}

constexpr int TWH = 64;    // tile width  in polyphase samples (128 output columns)
constexpr int THH = 32;    // tile height in polyphase samples (64 output rows)
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

template <typename T, int F> struct TileGeom {
  static constexpr int VEC = 16 / (int) sizeof (T);
  static constexpr int HP = filter_halo (F);
  static constexpr int HK = ((HP + VEC - 1) / VEC) * VEC;     // horizontal halo (aligned)
  static constexpr int PAD = VEC;                              // replicated border cells (>= 4)
  static constexpr int NK = TWH + 2 * HK;
  static constexpr int SEG = NK + 2 * PAD;
  static constexpr int PITCH = 2 * SEG;
  static constexpr int NR = 2 * (THH + 2 * HP);
  static constexpr size_t SMEM = (size_t) NR * PITCH * sizeof (T);
};

// One vertical lifting step on the smem window.
// Columns [c0, c1) of every row; polyphase rows ky in [ky_lo, ky_hi).
template <typename T, int F, int S, bool INV>
__device__ __forceinline__ void vert_step (T *sm, int pitch, int ky_lo, int ky_hi,
    int c0, int c1, int warp, int lane)
{
  constexpr Step st = step_of (F, S);
  constexpr int NT = kind_taps (st.kind);
  constexpr int sign = INV ? -st.sign : st.sign;
  for (int ky = ky_lo + warp; ky < ky_hi; ky += NWARPS) {
    const T *src[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) {
      int k = ky + st.tap0 + t;
      k = min (max (k, ky_lo), ky_hi - 1);   // window == picture range at picture edges
      src[t] = sm + (size_t) (2 * (k - ky_lo) + (1 - st.target)) * pitch;
    }
    T *dst = sm + (size_t) (2 * (ky - ky_lo) + st.target) * pitch;
    for (int c = c0 + lane; c < c1; c += 32) {
      int v[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) v[t] = src[t][c];
      int term = lift_term<T, st.kind> (v, st.p1, st.p2, st.p3);
      int d = dst[c];
      dst[c] = (T) (sign > 0 ? Ar<T>::add (d, term) : Ar<T>::sub (d, term));
    }
  }
}

// One horizontal lifting step: rows [r0, r1) of the window, nk samples per segment.
// Segment cells [-PAD,0) and [nk, nk+PAD) hold the replicated border (extend_*).
template <typename T, int F, int S, bool INV>
__device__ __forceinline__ void horiz_step (T *sm, int r0, int r1, int nk,
    int warp, int lane)
{
  typedef TileGeom<T, F> G;
  constexpr Step st = step_of (F, S);
  constexpr int NT = kind_taps (st.kind);
  constexpr int sign = INV ? -st.sign : st.sign;
  for (int r = r0 + warp; r < r1; r += NWARPS) {
    T *dst = sm + (size_t) r * G::PITCH + st.target * G::SEG + G::PAD;
    const T *src = sm + (size_t) r * G::PITCH + (1 - st.target) * G::SEG + G::PAD;
    for (int j = lane; j < nk; j += 32) {
      int v[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) v[t] = src[j + st.tap0 + t];
      int term = lift_term<T, st.kind> (v, st.p1, st.p2, st.p3);
      int d = dst[j];
      T nv = (T) (sign > 0 ? Ar<T>::add (d, term) : Ar<T>::sub (d, term));
      dst[j] = nv;
      if (j == 0) {
#pragma unroll
        for (int q = 1; q <= G::PAD; q++) dst[-q] = nv;
      }
      if (j == nk - 1) {
#pragma unroll
        for (int q = 0; q < G::PAD; q++) dst[nk + q] = nv;
      }
    }
  }
}

template <typename T, int F, bool INV, int S>
struct StepSeq {
  // run all steps of the filter in lifting order for this direction
  static __device__ __forceinline__ void vert (T *sm, int pitch, int ky_lo, int ky_hi,
      int c0, int c1, int warp, int lane)
  {
    constexpr int NS = num_steps (F);
    constexpr int s = INV ? NS - 1 - S : S;
    vert_step<T, F, s, INV> (sm, pitch, ky_lo, ky_hi, c0, c1, warp, lane);
    __syncthreads ();
    if constexpr (S + 1 < NS)
      StepSeq<T, F, INV, S + 1>::vert (sm, pitch, ky_lo, ky_hi, c0, c1, warp, lane);
  }
  static __device__ __forceinline__ void horiz (T *sm, int r0, int r1, int nk, int warp, int lane)
  {
    constexpr int NS = num_steps (F);
    constexpr int s = INV ? NS - 1 - S : S;
    horiz_step<T, F, s, INV> (sm, r0, r1, nk, warp, lane);
    __syncthreads ();
    if constexpr (S + 1 < NS)
      StepSeq<T, F, INV, S + 1>::horiz (sm, r0, r1, nk, warp, lane);
  }
};

template <typename T, int F, bool INV>
__global__ void __launch_bounds__ (NTHREADS)
wavelet_level_kernel (const LevelArgs a)
{
  typedef TileGeom<T, F> G;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *> (smem_raw);

  const TileId tile = level_tile (a);
  const int comp = tile.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp];
  const int n = w >> 1, m = h >> 1;
  const int kx0 = tile.bx * TWH, ky0 = tile.by * THH;
  if (kx0 >= n || ky0 >= m) return;

  const int kx_lo = max (0, kx0 - G::HK), kx_hi = min (n, kx0 + TWH + G::HK);
  const int ky_lo = max (0, ky0 - G::HP), ky_hi = min (m, ky0 + THH + G::HP);
  const int nk = kx_hi - kx_lo;
  const int nr = 2 * (ky_hi - ky_lo);
  const int twh = min (TWH, n - kx0), thh = min (THH, m - ky0);
  const int jo = kx0 - kx_lo;                // first output sample inside the segment
  const int ro = 2 * (ky0 - ky_lo);          // first output row inside the window

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int SH = filter_shift (F);

  T *dense = reinterpret_cast<T *> (plane_ptr (a.dense, pic, comp));
  T *bands = reinterpret_cast<T *> (plane_ptr (a.bands, pic, comp));
  T *ll = reinterpret_cast<T *> (plane_ptr (a.ll, pic, comp));
  const size_t ds = a.dense.stride[comp] / sizeof (T);
  const size_t bs = a.bands.stride[comp] / sizeof (T);
  const size_t ls = a.ll.stride[comp] / sizeof (T);

  if (INV) {
    // ---- load [L|H] split rows: LL from `ll`, the three detail bands from `bands`
    for (int lr = warp; lr < nr; lr += NWARPS) {
      const int r = 2 * ky_lo + lr;
      T *row = sm + (size_t) lr * G::PITCH + G::PAD;
      const T *gl = (r & 1) ? bands + (size_t) r * bs + kx_lo : ll + (size_t) (r >> 1) * ls + kx_lo;
      const T *gh = bands + (size_t) r * bs + n + kx_lo;
      for (int j = lane; j < nk; j += 32) {
        T vl = gl[j], vh = gh[j];
        row[j] = vl;
        row[G::SEG + j] = vh;
        if (j == 0) {
#pragma unroll
          for (int q = 1; q <= G::PAD; q++) { row[-q] = vl; row[G::SEG - q] = vh; }
        }
        if (j == nk - 1) {
#pragma unroll
          for (int q = 0; q < G::PAD; q++) { row[nk + q] = vl; row[G::SEG + nk + q] = vh; }
        }
      }
    }
    __syncthreads ();
    // ---- vertical un-lift on every column of the window (incl. replicated border cells)
    StepSeq<T, F, true, 0>::vert (sm, G::PITCH, ky_lo, ky_hi, 0, G::PITCH, warp, lane);
    // ---- horizontal un-lift on the output rows
    StepSeq<T, F, true, 0>::horiz (sm, ro, ro + 2 * thh, nk, warp, lane);
    // ---- interleave, (x+1)>>1, store
    for (int lr = warp; lr < 2 * thh; lr += NWARPS) {
      const int r = 2 * ky0 + lr;
      const T *row = sm + (size_t) (ro + lr) * G::PITCH + G::PAD + jo;
      T *out = dense + (size_t) r * ds + 2 * kx0;
      for (int x = lane; x < 2 * twh; x += 32) {
        int v = row[(x & 1) * G::SEG + (x >> 1)];
        if (SH) v = Ar<T>::add (v, 1) >> 1;
        out[x] = (T) v;
      }
    }
  } else {
    // ---- load interleaved rows, x<<1, de-interleave into [E|O] segments
    for (int lr = warp; lr < nr; lr += NWARPS) {
      const int r = 2 * ky_lo + lr;
      T *row = sm + (size_t) lr * G::PITCH + G::PAD;
      const T *in = dense + (size_t) r * ds + 2 * kx_lo;
      for (int x = lane; x < 2 * nk; x += 32) {
        int v = in[x];
        if (SH) v = Ar<T>::add (v, v);
        const int j = x >> 1, seg = x & 1;
        T *cell = row + seg * G::SEG + j;
        *cell = (T) v;
        if (j == 0) {
#pragma unroll
          for (int q = 1; q <= G::PAD; q++) cell[-q] = (T) v;
        }
        if (j == nk - 1) {
#pragma unroll
          for (int q = 1; q <= G::PAD; q++) cell[q] = (T) v;
        }
      }
    }
    __syncthreads ();
    // ---- horizontal lift on every row of the window
    StepSeq<T, F, false, 0>::horiz (sm, 0, nr, nk, warp, lane);
    // ---- vertical lift on every column of the window
    StepSeq<T, F, false, 0>::vert (sm, G::PITCH, ky_lo, ky_hi, 0, G::PITCH, warp, lane);
    // ---- store: LL (even rows, L half) to `ll`, everything else to `bands`
    for (int lr = warp; lr < 2 * thh; lr += NWARPS) {
      const int r = 2 * ky0 + lr;
      const T *row = sm + (size_t) (ro + lr) * G::PITCH + G::PAD + jo;
      T *ol = (r & 1) ? bands + (size_t) r * bs + kx0 : ll + (size_t) (r >> 1) * ls + kx0;
      T *oh = bands + (size_t) r * bs + n + kx0;
      for (int j = lane; j < twh; j += 32) {
        ol[j] = row[j];
        oh[j] = row[G::SEG + j];
      }
    }
  }
}

// ---- fast path: register-chunk lifting ----------------------------------------
// The generic kernel above pays ~6 warp instructions per sample, almost all of them
// shared-memory traffic of the step-by-step lifting.  The fast inverse kernel keeps every
// lifting step in registers: a thread owns a chunk of C polyphase pairs (+HP halo pairs
// each side), applies all steps of the filter to its private arrays E[] / O[] with
// compile-time indices, and only the finished pairs are written out.  Vertical chunks are
// loaded straight from global memory (lanes = consecutive columns, coalesced), their
// results go to shared memory once; horizontal chunks are read from there (lanes =
// consecutive rows, odd pitch => conflict-free), lifted, interleaved in place and copied
// out with coalesced stores.  Picture edges re-extend (replicate) the arrays after every
// step, exactly like the reference's extend_* helpers (schrowaveletorc.c:192-269).

// reach of one step to the left / right in the source array
__host__ __device__ constexpr int step_reach_l (int f, int s) { return step_of (f, s).tap0 < 0 ? -step_of (f, s).tap0 : 0; }
__host__ __device__ constexpr int step_reach_r (int f, int s)
{
  return step_of (f, s).tap0 + kind_taps (step_of (f, s).kind) - 1 > 0 ? step_of (f, s).tap0 + kind_taps (step_of (f, s).kind) - 1 : 0;
}
// total reach of the steps that still follow position `pos` (0-based, in execution order)
__host__ __device__ constexpr int reach_after (int f, bool inv, int pos, bool left)
{
  int r = 0;
  for (int p = pos + 1; p < num_steps (f); p++) {
    const int s = inv ? num_steps (f) - 1 - p : p;
    r += left ? step_reach_l (f, s) : step_reach_r (f, s);
  }
  return r;
}

// One lifting step on the private arrays.  Only the positions that later steps (and the
// final C outputs at [HP, N-HP)) still depend on are computed.
template <typename T, int F, bool INV, int POS, int N, int HP>
__device__ __forceinline__ void chunk_step (int (&E)[N], int (&O)[N])
{
  constexpr int S = INV ? num_steps (F) - 1 - POS : POS;
  constexpr Step st = step_of (F, S);
  constexpr int NT = kind_taps (st.kind);
  constexpr int sign = INV ? -st.sign : st.sign;
  constexpr int K0 = HP - reach_after (F, INV, POS, true) < 0 ? 0 : HP - reach_after (F, INV, POS, true);
  constexpr int K1 = N - HP + reach_after (F, INV, POS, false) > N ? N : N - HP + reach_after (F, INV, POS, false);
#pragma unroll
  for (int k = K0; k < K1; k++) {
    int v[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const int idx = (k + st.tap0 + t) < 0 ? 0 : ((k + st.tap0 + t) > N - 1 ? N - 1 : (k + st.tap0 + t));
      v[t] = st.target ? E[idx] : O[idx];
    }
    const int term = lift_term<T, st.kind> (v, st.p1, st.p2, st.p3);
    if (st.target) O[k] = sign > 0 ? Ar<T>::add (O[k], term) : Ar<T>::sub (O[k], term);
    else E[k] = sign > 0 ? Ar<T>::add (E[k], term) : Ar<T>::sub (E[k], term);
  }
}

template <int N, int HP>
__device__ __forceinline__ void chunk_extend (int (&X)[N], bool lo, bool hi)
{
  if (HP > 0) {
    if (lo) {
#pragma unroll
      for (int h = 0; h < HP; h++) X[h] = X[HP];
    }
    if (hi) {
#pragma unroll
      for (int h = 0; h < HP; h++) X[N - 1 - h] = X[N - 1 - HP];
    }
  }
}

template <typename T, int F, bool INV, int N, int HP, int S>
struct ChunkSeq {
  static __device__ __forceinline__ void run (int (&E)[N], int (&O)[N], bool lo, bool hi)
  {
    constexpr int NS = num_steps (F);
    constexpr int s = INV ? NS - 1 - S : S;
    chunk_step<T, F, INV, S, N, HP> (E, O);
    if (step_of (F, s).target) chunk_extend<N, HP> (O, lo, hi);
    else chunk_extend<N, HP> (E, lo, hi);
    if constexpr (S + 1 < NS) ChunkSeq<T, F, INV, N, HP, S + 1>::run (E, O, lo, hi);
  }
};

template <typename T, int F, bool INV, int N, int HP>
__device__ __forceinline__ void chunk_lift (int (&E)[N], int (&O)[N], bool lo, bool hi)
{
  chunk_extend<N, HP> (E, lo, hi);
  chunk_extend<N, HP> (O, lo, hi);
  ChunkSeq<T, F, INV, N, HP, 0>::run (E, O, lo, hi);
}

template <typename T, int F, int CS> struct FastGeom {
  static constexpr int VEC = 16 / (int) sizeof (T);
  static constexpr int HP = filter_halo (F);
  static constexpr int HK = ((HP + VEC - 1) / VEC) * VEC;   // horizontal halo, 16-byte granular
  static constexpr int C = CS;                 // vertical chunk (polyphase rows)
  static constexpr int CH = CS;                // horizontal chunk (polyphase columns)
  static constexpr int NKV = TWH + 2 * HK;     // columns per segment held in shared memory
  static constexpr int NV = C + 2 * HP;
  static constexpr int NH = CH + 2 * HK;
  // row pitch: a multiple of 16 bytes that is 4 (mod 32) in 32-bit words, so that lanes =
  // rows reading 128 bits each are conflict-free (a quarter warp covers all 32 banks)
  static constexpr int MINW = 2 * NKV * (int) sizeof (T) / 4;
  static constexpr int PITCHW = ((MINW - 4 + 31) / 32) * 32 + 4;
  static constexpr int PITCH = PITCHW * 4 / (int) sizeof (T);
  static constexpr int VITEMS = 2 * NKV * (THH / C);
  static constexpr int HITEMS = 2 * THH * (TWH / CH);
  static constexpr int NT = ((VITEMS > HITEMS ? VITEMS : HITEMS) + 31) / 32 * 32;
  static constexpr size_t SMEM = (size_t) 2 * THH * PITCH * sizeof (T);
  static_assert (PITCHW >= MINW && PITCHW % 32 == 4, "bad pitch");
};

// 16 bytes of shared memory <-> VEC ints
template <typename T> struct Vec16;
template <> struct Vec16<int32_t> {
  static __device__ __forceinline__ void load (const int32_t *p, int *v)
  {
    const int4 q = *reinterpret_cast<const int4 *> (p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  // 4 interleaved outputs (e0 o0 e1 o1)
  static __device__ __forceinline__ void store_pairs (int32_t *p, const int *e, const int *o)
  {
    *reinterpret_cast<int4 *> (p) = make_int4 (e[0], o[0], e[1], o[1]);
  }
  // 4 interleaved inputs (e0 o0 e1 o1)
  static __device__ __forceinline__ void load_pairs (const int32_t *p, int *e, int *o)
  {
    const int4 q = *reinterpret_cast<const int4 *> (p);
    e[0] = q.x; o[0] = q.y; e[1] = q.z; o[1] = q.w;
  }
  static __device__ __forceinline__ void store (int32_t *p, const int *v)
  {
    *reinterpret_cast<int4 *> (p) = make_int4 (v[0], v[1], v[2], v[3]);
  }
  static constexpr int PAIRS = 2;
};
template <> struct Vec16<int16_t> {
  static __device__ __forceinline__ void load (const int16_t *p, int *v)
  {
    const int4 q = *reinterpret_cast<const int4 *> (p);
    v[0] = (q.x << 16) >> 16; v[1] = q.x >> 16; v[2] = (q.y << 16) >> 16; v[3] = q.y >> 16;
    v[4] = (q.z << 16) >> 16; v[5] = q.z >> 16; v[6] = (q.w << 16) >> 16; v[7] = q.w >> 16;
  }
  static __device__ __forceinline__ void store_pairs (int16_t *p, const int *e, const int *o)
  {
    *reinterpret_cast<int4 *> (p) = make_int4 ((e[0] & 0xffff) | (o[0] << 16), (e[1] & 0xffff) | (o[1] << 16),
        (e[2] & 0xffff) | (o[2] << 16), (e[3] & 0xffff) | (o[3] << 16));
  }
  static __device__ __forceinline__ void load_pairs (const int16_t *p, int *e, int *o)
  {
    const int4 q = *reinterpret_cast<const int4 *> (p);
    e[0] = (q.x << 16) >> 16; o[0] = q.x >> 16; e[1] = (q.y << 16) >> 16; o[1] = q.y >> 16;
    e[2] = (q.z << 16) >> 16; o[2] = q.z >> 16; e[3] = (q.w << 16) >> 16; o[3] = q.w >> 16;
  }
  static __device__ __forceinline__ void store (int16_t *p, const int *v)
  {
    *reinterpret_cast<int4 *> (p) = make_int4 ((v[0] & 0xffff) | (v[1] << 16), (v[2] & 0xffff) | (v[3] << 16),
        (v[4] & 0xffff) | (v[5] << 16), (v[6] & 0xffff) | (v[7] << 16));
  }
  static constexpr int PAIRS = 4;
};

template <typename T, int F, int CS, bool OUT8 = false>
__global__ void __launch_bounds__ (FastGeom<T, F, CS>::NT)
wavelet_inv_fast_kernel (const LevelArgs a)
{
  typedef FastGeom<T, F, CS> G;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *> (smem_raw);

  const TileId tile = level_tile (a);
  const int comp = tile.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp];
  const int n = w >> 1, m = h >> 1;
  const int kx0 = tile.bx * TWH, ky0 = tile.by * THH;
  if (kx0 >= n || ky0 >= m) return;
  constexpr int SH = filter_shift (F);

  T *dense = reinterpret_cast<T *> (plane_ptr (a.dense, pic, comp));
  const T *bands = reinterpret_cast<const T *> (plane_ptr (a.bands, pic, comp));
  const T *ll = reinterpret_cast<const T *> (plane_ptr (a.ll, pic, comp));
  const size_t ds = a.dense.stride[comp] / sizeof (T);
  const size_t bs = a.bands.stride[comp] / sizeof (T);
  const size_t ls = a.ll.stride[comp] / sizeof (T);
  const int tid = threadIdx.x;

  // ---- vertical: one thread = one column of one segment x one chunk of C row pairs ----
  if (tid < G::VITEMS) {
    const int chunk = tid / (2 * G::NKV), col = tid - chunk * (2 * G::NKV);
    const int seg = col >= G::NKV, cseg = col - seg * G::NKV;
    const int kx = kx0 - G::HK + cseg;
    const int kyc = ky0 + chunk * G::C;                 // first pair row of this chunk
    if (kx >= 0 && kx < n && kyc < m) {
      int E[G::NV], O[G::NV];
      const T *pe = seg ? bands + n + kx : ll + kx;      // even rows: LL (seg 0) or HL (seg 1)
      const size_t es = seg ? 2 * bs : ls;
      const T *po = bands + bs + seg * n + kx;           // odd rows: LH / HH
      const bool lo = kyc == 0, hi = kyc + G::C >= m;
      if (!lo && !hi) {
        pe += (size_t) (kyc - G::HP) * es;
        po += (size_t) (kyc - G::HP) * 2 * bs;
#pragma unroll
        for (int r = 0; r < G::NV; r++) {
          E[r] = pe[(size_t) r * es];
          O[r] = po[(size_t) r * 2 * bs];
        }
      } else {
#pragma unroll
        for (int r = 0; r < G::NV; r++) {
          const int ky = kyc - G::HP + r;
          const bool ok = ky >= 0 && ky < m;
          E[r] = ok ? (int) pe[(size_t) ky * es] : 0;
          O[r] = ok ? (int) po[(size_t) ky * 2 * bs] : 0;
        }
      }
      chunk_lift<T, F, true, G::NV, G::HP> (E, O, lo, hi);
      T *out = sm + (size_t) (chunk * 2 * G::C) * G::PITCH + col;
#pragma unroll
      for (int r = 0; r < G::C; r++) {
        out[(size_t) (2 * r) * G::PITCH] = (T) E[G::HP + r];
        out[(size_t) (2 * r + 1) * G::PITCH] = (T) O[G::HP + r];
      }
    }
  }
  __syncthreads ();

  // ---- horizontal: one thread = one row x one chunk of CH column pairs; lanes = rows ----
  const int hw = tid >> 5, lane = tid & 31;
  const int q = hw % (TWH / G::CH);                      // chunk within the tile
  const int row = (hw / (TWH / G::CH)) * 32 + lane;      // tile row
  const int k0 = kx0 + q * G::CH;
  const bool hact = tid < G::HITEMS && k0 < n && (2 * ky0 + row) < h;
  int E[G::NH], O[G::NH];
  if (hact) {
    const T *rowp = sm + (size_t) row * G::PITCH + q * G::CH;     // = column kx0 - HK + q*CH
#pragma unroll
    for (int i = 0; i < G::NH; i += G::VEC) {
      Vec16<T>::load (rowp + i, &E[i]);
      Vec16<T>::load (rowp + G::NKV + i, &O[i]);
    }
    chunk_lift<T, F, true, G::NH, G::HK> (E, O, k0 == 0, k0 + G::CH >= n);
  }
  __syncthreads ();
  if (hact) {
    // interleave + (x+1)>>1 in place: output column x of the tile at row*PITCH + x
    T *orow = sm + (size_t) row * G::PITCH + 2 * q * G::CH;
    if (SH) {
#pragma unroll
      for (int i = 0; i < G::CH; i++) {
        E[G::HK + i] = Ar<T>::add (E[G::HK + i], 1) >> 1;
        O[G::HK + i] = Ar<T>::add (O[G::HK + i], 1) >> 1;
      }
    }
#pragma unroll
    for (int i = 0; i < G::CH; i += Vec16<T>::PAIRS)
      Vec16<T>::store_pairs (orow + 2 * i, &E[G::HK + i], &O[G::HK + i]);
  }
  __syncthreads ();

  // ---- coalesced copy-out, 128 bits per lane ----
  {
    const int tw = min (2 * TWH, w - 2 * kx0);           // output columns of this tile (multiple of 32)
    const int th = min (2 * THH, h - 2 * ky0);
    const int vpr = tw / G::VEC;                         // 16-byte vectors per row
    const int total = th * vpr;
    if (OUT8) {
      // fused combine: the picture leaves as 8-bit samples (a quarter / half of the bytes), cropped to its size
      uint8_t *out = reinterpret_cast<uint8_t *> (plane_ptr (a.dense, pic, comp));
      const size_t os = (size_t) a.dense.stride[comp];
      const int ow = a.out_w[comp], oh = a.out_h[comp];
      for (int i = tid; i < total; i += G::NT) {
        const int r = i / vpr, x = i - r * vpr;
        const int py = 2 * ky0 + r, px = 2 * kx0 + x * G::VEC;
        if (py >= oh || px >= ow) continue;
        int v[G::VEC];
        Vec16<T>::load (sm + (size_t) r * G::PITCH + x * G::VEC, v);
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < G::VEC; k++) {
          const unsigned b = combine_to_u8<T> (v[k], a.out_shift);
          if (k < 4) lo |= b << (8 * k); else hi |= b << (8 * (k - 4));
        }
        uint8_t *o = out + (size_t) py * os + px;
        if (px + G::VEC <= ow && (((size_t) o) & (G::VEC - 1)) == 0) {
          if (G::VEC == 4) *reinterpret_cast<unsigned *> (o) = lo;
          else *reinterpret_cast<uint2 *> (o) = make_uint2 (lo, hi);
        } else {
          for (int k = 0; k < G::VEC && px + k < ow; k++) o[k] = (uint8_t) ((k < 4 ? lo >> (8 * k) : hi >> (8 * (k - 4))) & 0xff);
        }
      }
    } else {
      for (int i = tid; i < total; i += G::NT) {
        const int r = i / vpr, x = i - r * vpr;
        const int4 v = *reinterpret_cast<const int4 *> (sm + (size_t) r * G::PITCH + x * G::VEC);
        *reinterpret_cast<int4 *> (dense + (size_t) (2 * ky0 + r) * ds + 2 * kx0 + x * G::VEC) = v;
      }
    }
  }
}


// ---- fused inverse kernel: levels 1 and 0 in one launch -------------------------------------
// BASELINE.json's "multi-level fused tiles": the CTA of a level-0 output tile first computes the part
// of level 1's output (= level 0's LL band) its tile needs -- the tile's LL window plus the filter's
// halo -- into shared memory, then runs the level-0 tile exactly like wavelet_inv_fast_kernel with
// the LL rows coming from there.  Level 1's dense plane is never written to or read from HBM (2/3 of
// a plane per picture saved); the price is the halo: a 64 x 32 LL window needs 72 x 40 level-1
// outputs (1.41 x level 1's arithmetic) and 88 x 56 level-1 inputs per band pair.
// Stage 1 uses the same register-chunk lifting: two vertical chunks of C1 = 10 pair rows, then
// NPX / CH1 horizontal chunks.  The window is clamped INTO the plane (not cut at its edge), so that a
// chunk that reaches a picture edge ends exactly there -- what chunk_lift's edge extension assumes.
struct FusedArgs {
  LevelArgs l0;     // level 0: dense = output, bands = level-0 bands; l0.ll is unused
  PlaneSet bands1;  // the coefficient plane at level 1's stride
  PlaneSet ll1;     // level 1's LL input (w/4 x h/4)
};

template <typename T, int F, int CS> struct FusedGeom {
  typedef FastGeom<T, F, CS> B;
  static constexpr int HP = B::HP;
  static constexpr int HPE = (HP + 1) & ~1;                   // even halo: the window starts on a level-1 pair
  static constexpr int C1 = (THH + 2 * HPE) / 4;              // pair rows per vertical chunk (two chunks)
  static constexpr int NPY = 2 * C1;                          // level-1 pair rows of the window
  static constexpr int NPX = (TWH + 2 * B::HK) / 2;           // level-1 pair columns of the window
  static constexpr int CH1 = NPX % 12 == 0 ? 12 : 10;         // horizontal chunk
  static constexpr int NQ = NPX / CH1;
  static constexpr int NC1 = NPX + 2 * HP;                    // columns per segment after the vertical pass
  static constexpr int NV1 = C1 + 2 * HP;
  static constexpr int NH1 = CH1 + 2 * HP;
  static constexpr int ROWS = 2 * NPY;                        // level-1 output rows of the window
  static constexpr int WPE = 4 / (int) sizeof (T);            // elements per 32-bit word
  static constexpr int S1P = ((2 * NC1 + WPE - 1) / WPE | 1) * WPE;     // odd number of words: lanes = rows conflict-free
  static constexpr int LLP = ((2 * NPX + WPE - 1) / WPE | 1) * WPE;
  static constexpr int V1ITEMS = 2 * 2 * NC1;
  static constexpr int H1ITEMS = ROWS * NQ;
  static constexpr size_t LL_BYTES = ((size_t) ROWS * LLP * sizeof (T) + 15) / 16 * 16;
  static constexpr size_t SMEM = B::SMEM + LL_BYTES;
  static_assert ((THH + 2 * HPE) % 4 == 0 && NPX % CH1 == 0, "window does not split into chunks");
  static_assert (V1ITEMS <= B::NT && H1ITEMS <= B::NT, "stage 1 needs more threads than the tile kernel has");
  static_assert ((size_t) ROWS * S1P * sizeof (T) <= B::SMEM, "stage-1 buffer must fit the tile buffer it aliases");
};

template <typename T, int F, int CS>
__global__ void __launch_bounds__ (FastGeom<T, F, CS>::NT)
wavelet_inv_fused2_kernel (const FusedArgs fa)
{
  typedef FastGeom<T, F, CS> G;
  typedef FusedGeom<T, F, CS> U;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *> (smem_raw);
  T *lls = reinterpret_cast<T *> (smem_raw + G::SMEM);
  const LevelArgs &a = fa.l0;

  const TileId tile = level_tile (a);
  const int comp = tile.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp];
  const int n = w >> 1, m = h >> 1;
  const int kx0 = tile.bx * TWH, ky0 = tile.by * THH;
  if (kx0 >= n || ky0 >= m) return;
  constexpr int SH = filter_shift (F);
  const int tid = threadIdx.x;

  // ---- stage 1: level 1's output rows [2 sy1, 2 sy1 + ROWS) x columns [2 sx1, 2 sx1 + 2 NPX) -> lls ----
  const int n1 = n >> 1, m1 = m >> 1;
  const int sx1 = max (0, min ((kx0 - G::HK) >> 1, n1 - U::NPX));
  const int sy1 = max (0, min ((ky0 - U::HPE) >> 1, m1 - U::NPY));
  {
    const T *bands1 = reinterpret_cast<const T *> (plane_ptr (fa.bands1, pic, comp));
    const T *ll1 = reinterpret_cast<const T *> (plane_ptr (fa.ll1, pic, comp));
    const size_t bs1 = fa.bands1.stride[comp] / sizeof (T), ls1 = fa.ll1.stride[comp] / sizeof (T);
    if (tid < U::V1ITEMS) {
      const int chunk = tid / (2 * U::NC1), col = tid - chunk * (2 * U::NC1);
      const int seg = col >= U::NC1, j = col - seg * U::NC1;
      const int kx = sx1 - U::HP + j;
      const int kyc = sy1 + chunk * U::C1;
      if (kx >= 0 && kx < n1) {
        int E[U::NV1], O[U::NV1];
        const T *pe = seg ? bands1 + n1 + kx : ll1 + kx;
        const size_t es = seg ? 2 * bs1 : ls1;
        const T *po = bands1 + bs1 + seg * n1 + kx;
        const bool lo = kyc == 0, hi = kyc + U::C1 >= m1;
        if (!lo && !hi) {
          pe += (size_t) (kyc - U::HP) * es;
          po += (size_t) (kyc - U::HP) * 2 * bs1;
#pragma unroll
          for (int r = 0; r < U::NV1; r++) {
            E[r] = pe[(size_t) r * es];
            O[r] = po[(size_t) r * 2 * bs1];
          }
        } else {
#pragma unroll
          for (int r = 0; r < U::NV1; r++) {
            const int ky = kyc - U::HP + r;
            const bool ok = ky >= 0 && ky < m1;
            E[r] = ok ? (int) pe[(size_t) ky * es] : 0;
            O[r] = ok ? (int) po[(size_t) ky * 2 * bs1] : 0;
          }
        }
        chunk_lift<T, F, true, U::NV1, U::HP> (E, O, lo, hi);
        T *out = sm + (size_t) (chunk * 2 * U::C1) * U::S1P + col;
#pragma unroll
        for (int r = 0; r < U::C1; r++) {
          out[(size_t) (2 * r) * U::S1P] = (T) E[U::HP + r];
          out[(size_t) (2 * r + 1) * U::S1P] = (T) O[U::HP + r];
        }
      }
    }
    __syncthreads ();
    if (tid < U::H1ITEMS) {
      const int q = tid / U::ROWS, row = tid - q * U::ROWS;     // lanes = rows
      const int k0 = sx1 + q * U::CH1;
      int E[U::NH1], O[U::NH1];
      const T *rowp = sm + (size_t) row * U::S1P + q * U::CH1;
#pragma unroll
      for (int i = 0; i < U::NH1; i++) { E[i] = rowp[i]; O[i] = rowp[U::NC1 + i]; }
      chunk_lift<T, F, true, U::NH1, U::HP> (E, O, k0 == 0, k0 + U::CH1 >= n1);
      T *orow = lls + (size_t) row * U::LLP + 2 * q * U::CH1;
#pragma unroll
      for (int i = 0; i < U::CH1; i++) {
        int e = E[U::HP + i], o = O[U::HP + i];
        if (SH) { e = Ar<T>::add (e, 1) >> 1; o = Ar<T>::add (o, 1) >> 1; }
        orow[2 * i] = (T) e;
        orow[2 * i + 1] = (T) o;
      }
    }
    __syncthreads ();
  }

  // ---- level 0, as wavelet_inv_fast_kernel, the LL rows read from lls ----
  T *dense = reinterpret_cast<T *> (plane_ptr (a.dense, pic, comp));
  const T *bands = reinterpret_cast<const T *> (plane_ptr (a.bands, pic, comp));
  const size_t ds = a.dense.stride[comp] / sizeof (T);
  const size_t bs = a.bands.stride[comp] / sizeof (T);
  if (tid < G::VITEMS) {
    const int chunk = tid / (2 * G::NKV), col = tid - chunk * (2 * G::NKV);
    const int seg = col >= G::NKV, cseg = col - seg * G::NKV;
    const int kx = kx0 - G::HK + cseg;
    const int kyc = ky0 + chunk * G::C;
    if (kx >= 0 && kx < n && kyc < m) {
      int E[G::NV], O[G::NV];
      const T *po = bands + bs + seg * n + kx;
      const bool lo = kyc == 0, hi = kyc + G::C >= m;
      if (seg) {
        const T *pe = bands + n + kx;
#pragma unroll
        for (int r = 0; r < G::NV; r++) {
          const int ky = kyc - G::HP + r;
          const bool ok = ky >= 0 && ky < m;
          E[r] = ok ? (int) pe[(size_t) ky * 2 * bs] : 0;
          O[r] = ok ? (int) po[(size_t) ky * 2 * bs] : 0;
        }
      } else {
        const T *pe = lls + (kx - 2 * sx1) - (ptrdiff_t) (2 * sy1) * U::LLP;
#pragma unroll
        for (int r = 0; r < G::NV; r++) {
          const int ky = kyc - G::HP + r;
          const bool ok = ky >= 0 && ky < m;
          E[r] = ok ? (int) pe[(ptrdiff_t) ky * U::LLP] : 0;
          O[r] = ok ? (int) po[(size_t) ky * 2 * bs] : 0;
        }
      }
      chunk_lift<T, F, true, G::NV, G::HP> (E, O, lo, hi);
      T *out = sm + (size_t) (chunk * 2 * G::C) * G::PITCH + col;
#pragma unroll
      for (int r = 0; r < G::C; r++) {
        out[(size_t) (2 * r) * G::PITCH] = (T) E[G::HP + r];
        out[(size_t) (2 * r + 1) * G::PITCH] = (T) O[G::HP + r];
      }
    }
  }
  __syncthreads ();

  const int hw = tid >> 5, lane = tid & 31;
  const int q = hw % (TWH / G::CH);
  const int row = (hw / (TWH / G::CH)) * 32 + lane;
  const int k0 = kx0 + q * G::CH;
  const bool hact = tid < G::HITEMS && k0 < n && (2 * ky0 + row) < h;
  int E[G::NH], O[G::NH];
  if (hact) {
    const T *rowp = sm + (size_t) row * G::PITCH + q * G::CH;
#pragma unroll
    for (int i = 0; i < G::NH; i += G::VEC) {
      Vec16<T>::load (rowp + i, &E[i]);
      Vec16<T>::load (rowp + G::NKV + i, &O[i]);
    }
    chunk_lift<T, F, true, G::NH, G::HK> (E, O, k0 == 0, k0 + G::CH >= n);
  }
  __syncthreads ();
  if (hact) {
    T *orow = sm + (size_t) row * G::PITCH + 2 * q * G::CH;
    if (SH) {
#pragma unroll
      for (int i = 0; i < G::CH; i++) {
        E[G::HK + i] = Ar<T>::add (E[G::HK + i], 1) >> 1;
        O[G::HK + i] = Ar<T>::add (O[G::HK + i], 1) >> 1;
      }
    }
#pragma unroll
    for (int i = 0; i < G::CH; i += Vec16<T>::PAIRS)
      Vec16<T>::store_pairs (orow + 2 * i, &E[G::HK + i], &O[G::HK + i]);
  }
  __syncthreads ();
  {
    const int tw = min (2 * TWH, w - 2 * kx0);
    const int th = min (2 * THH, h - 2 * ky0);
    const int vpr = tw / G::VEC;
    const int total = th * vpr;
    for (int i = tid; i < total; i += G::NT) {
      const int r = i / vpr, x = i - r * vpr;
      const int4 v = *reinterpret_cast<const int4 *> (sm + (size_t) r * G::PITCH + x * G::VEC);
      *reinterpret_cast<int4 *> (dense + (size_t) (2 * ky0 + r) * ds + 2 * kx0 + x * G::VEC) = v;
    }
  }
}


// ---- fast forward kernel ---------------------------------------------------------------
// Mirror image of the fast inverse kernel: the tile (with its halo) is copied in with
// coalesced 128-bit loads, each thread lifts one row x one chunk of column pairs out of
// shared memory (lanes = rows), writes the [L|H] split back, then one thread lifts one column
// x one chunk of row pairs (lanes = columns) and stores the four bands straight to global
// memory with coalesced scalar stores.
template <typename T, int F, int CS> struct FwdGeom {
  typedef FastGeom<T, F, CS> B;
  static constexpr int ROWS = 2 * (THH + 2 * B::HP);          // input rows held in shared memory
  static constexpr int HITEMS = (ROWS + 31) / 32 * 32 * (TWH / B::CH);     // lanes = rows, whole warps
  static constexpr int VITEMS = 2 * TWH * (THH / B::C);
  static constexpr int NT = ((VITEMS > HITEMS ? VITEMS : HITEMS) + 31) / 32 * 32;
  static constexpr size_t SMEM = (size_t) ROWS * B::PITCH * sizeof (T);
  static_assert (NT <= 1024, "too many items per tile");
};

template <typename T, int F, int CS>
__global__ void __launch_bounds__ (FwdGeom<T, F, CS>::NT)
wavelet_fwd_fast_kernel (const LevelArgs a)
{
  typedef FastGeom<T, F, CS> G;
  typedef FwdGeom<T, F, CS> FG;
  extern __shared__ __align__ (16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *> (smem_raw);

  const TileId tile = level_tile (a);
  const int comp = tile.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp];
  const int n = w >> 1, m = h >> 1;
  const int kx0 = tile.bx * TWH, ky0 = tile.by * THH;
  if (kx0 >= n || ky0 >= m) return;
  constexpr int SH = filter_shift (F);

  const T *dense = reinterpret_cast<const T *> (plane_ptr (a.dense, pic, comp));
  T *bands = reinterpret_cast<T *> (plane_ptr (a.bands, pic, comp));
  T *ll = reinterpret_cast<T *> (plane_ptr (a.ll, pic, comp));
  const size_t ds = a.dense.stride[comp] / sizeof (T);
  const size_t bs = a.bands.stride[comp] / sizeof (T);
  const size_t ls = a.ll.stride[comp] / sizeof (T);
  const int tid = threadIdx.x;

  // shared-memory row r <-> picture row 2 (ky0 - HP) + r; column c <-> picture column 2 (kx0 - HK) + c
  const int py0 = 2 * (ky0 - G::HP), px0 = 2 * (kx0 - G::HK);
  // ---- coalesced copy-in, 128 bits per lane ----
  {
    const int r_lo = max (0, -py0), r_hi = min (FG::ROWS, h - py0);
    const int c_lo = max (0, -px0), c_hi = min (2 * G::NKV, w - px0);      // multiples of VEC
    const int vpr = (c_hi - c_lo) / G::VEC;
    const int total = (r_hi - r_lo) * vpr;
    for (int i = tid; i < total; i += FG::NT) {
      const int r = r_lo + i / vpr, c = c_lo + (i % vpr) * G::VEC;
      const int4 v = *reinterpret_cast<const int4 *> (dense + (size_t) (py0 + r) * ds + px0 + c);
      *reinterpret_cast<int4 *> (sm + (size_t) r * G::PITCH + c) = v;
    }
  }
  __syncthreads ();

  // ---- horizontal: one thread = one row x one chunk of CH column pairs; lanes = rows ----
  {
    const int hw = tid >> 5, lane = tid & 31;
    constexpr int CPR = TWH / G::CH;                       // chunks per row
    const int q = hw % CPR;
    const int row = (hw / CPR) * 32 + lane;                // shared-memory row
    const int k0 = kx0 + q * G::CH;
    const bool hact = tid < FG::HITEMS && row < FG::ROWS && k0 < n && py0 + row >= 0 && py0 + row < h;
    int E[G::NH], O[G::NH];
    if (hact) {
      const T *rowp = sm + (size_t) row * G::PITCH + 2 * q * G::CH;     // pair k0 - HK
#pragma unroll
      for (int i = 0; i < G::NH; i += Vec16<T>::PAIRS)
        Vec16<T>::load_pairs (rowp + 2 * i, &E[i], &O[i]);
      if (SH) {
#pragma unroll
        for (int i = 0; i < G::NH; i++) { E[i] = Ar<T>::add (E[i], E[i]); O[i] = Ar<T>::add (O[i], O[i]); }
      }
      chunk_lift<T, F, false, G::NH, G::HK> (E, O, k0 == 0, k0 + G::CH >= n);
    }
    __syncthreads ();
    if (hact) {
      // split layout: L at column (pair - (kx0 - HK)), H at NKV + the same
      T *orow = sm + (size_t) row * G::PITCH + q * G::CH + G::HK;
#pragma unroll
      for (int i = 0; i < G::CH; i += G::VEC) {
        Vec16<T>::store (orow + i, &E[G::HK + i]);
        Vec16<T>::store (orow + G::NKV + i, &O[G::HK + i]);
      }
    }
  }
  __syncthreads ();

  // ---- vertical: one thread = one column x one chunk of C row pairs; lanes = columns ----
  if (tid < FG::VITEMS) {
    const int chunk = tid / (2 * TWH), col = tid - chunk * (2 * TWH);
    const int seg = col >= TWH, cc = col - seg * TWH;
    const int kx = kx0 + cc;
    const int kyc = ky0 + chunk * G::C;
    if (kx < n && kyc < m) {
      int E[G::NV], O[G::NV];
      const T *colp = sm + (size_t) (2 * chunk * G::C) * G::PITCH + seg * G::NKV + G::HK + cc;
#pragma unroll
      for (int r = 0; r < G::NV; r++) {
        E[r] = colp[(size_t) (2 * r) * G::PITCH];
        O[r] = colp[(size_t) (2 * r + 1) * G::PITCH];
      }
      chunk_lift<T, F, false, G::NV, G::HP> (E, O, kyc == 0, kyc + G::C >= m);
      // even rows: LL (seg 0) to `ll`, HL (seg 1) to the band plane; odd rows: LH / HH
      T *pe = seg ? bands + n + kx : ll + kx;
      const size_t es = seg ? 2 * bs : ls;
      T *po = bands + bs + seg * n + kx;
#pragma unroll
      for (int r = 0; r < G::C; r++) {
        pe[(size_t) (kyc + r) * es] = (T) E[G::HP + r];
        po[(size_t) (kyc + r) * 2 * bs] = (T) O[G::HP + r];
      }
    }
  }
}

// chunk size (16, 8) usable by the fast inverse kernel for component c, or 0
template <typename T, int F>
static int fast_inverse_chunk (const LevelArgs &a, int c)
{
  // 128-bit copy-out: rows of the dense plane must be 16-byte aligned
  if ((a.dense.stride[c] % 16) || (a.dense.off[c] % 16)) return 0;
  if (((size_t) a.dense.base % 16) || (a.dense.pic_pitch % 16)) return 0;
  const int n = a.w[c] >> 1, m = a.h[c] >> 1;
  // Fidelity (halo 8) only with chunks of 8: with 16 the per-thread register arrays spill
  if (n % 16 == 0 && m % 16 == 0 && filter_halo (F) <= 4) return 16;
  if (n % 8 == 0 && m % 8 == 0) return 8;
  return 0;
}

// ---- host side ---------------------------------------------------------------

// selects components `comps[0..nsel)` for a launch and lays their tiles out on blockIdx.x
static dim3 level_grid (LevelArgs &a, const int *comps, int nsel, int count)
{
  a.ncomp = nsel;
  int total = 0;
  for (int i = 0; i < SB2_MAX_COMPONENTS; i++) { a.tile_start[i] = 0; a.tiles_x[i] = 1; }
  for (int i = 0; i < nsel; i++) {
    a.comp_map[i] = comps[i];
    a.tile_start[i] = total;
    a.tiles_x[i] = ceil_div (a.w[comps[i]] >> 1, TWH);
    total += a.tiles_x[i] * ceil_div (a.h[comps[i]] >> 1, THH);
  }
  a.tile_start[nsel < SB2_MAX_COMPONENTS ? nsel : SB2_MAX_COMPONENTS] = total;
  return dim3 (total, count, 1);
}


template <typename T, int F, int CS>
static int launch_fast_inverse (LevelArgs a, const int *comps, int nsel, int count, cudaStream_t stream,
    const char *tag, double bytes)
{
  typedef FastGeom<T, F, CS> FG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute (wavelet_inv_fast_kernel<T, F, CS>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FG::SMEM);
    if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet fast)");
    attr_set = true;
  }
  const dim3 grid = level_grid (a, comps, nsel, count);
  {
    LaunchScope scope (tag, bytes, stream);
    wavelet_inv_fast_kernel<T, F, CS><<<grid, FG::NT, FG::SMEM, stream>>> (a);
  }
  return check_cuda (cudaGetLastError (), "wavelet_inv_fast_kernel launch");
}

// 0: one launch per level (default -- measured faster), 1: levels 1 + 0 fused whenever the
// shapes allow (sb2_iwt_enable_fused, or SB2_IWT_FUSED=1 in the environment)
static int g_iwt_fused = -1;
static bool iwt_fused_enabled ()
{
  if (g_iwt_fused < 0) {
    const char *v = getenv ("SB2_IWT_FUSED");
    g_iwt_fused = (v && atoi (v)) ? 1 : 0;
  }
  return g_iwt_fused != 0;
}

// can levels 1 + 0 of the inverse transform run fused for every component?  (a0 = level 0's arguments)
template <typename T, int F>
static bool fused2_supported (const LevelArgs &a0)
{
  if (!(F == 0 || F == 2 || F == 6)) return false;
  if constexpr (F == 0 || F == 2 || F == 6) {
    typedef FusedGeom<T, F, 16> U;
    for (int c = 0; c < a0.ncomp; c++) {
      if (fast_inverse_chunk<T, F> (a0, c) != 16) return false;
      if ((a0.w[c] >> 2) < U::NPX || (a0.h[c] >> 2) < U::NPY || (a0.w[c] & 3) || (a0.h[c] & 3)) return false;
    }
    return true;
  }
  return false;
}

template <typename T, int F>
static int launch_fused2 (const LevelArgs &a0, const PlaneSet &bands1, const PlaneSet &ll1, int count, cudaStream_t stream)
{
  if constexpr (F == 0 || F == 2 || F == 6) {
    typedef FastGeom<T, F, 16> FG;
    typedef FusedGeom<T, F, 16> U;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute (wavelet_inv_fused2_kernel<T, F, 16>,
          cudaFuncAttributeMaxDynamicSharedMemorySize, (int) U::SMEM);
      if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet fused)");
      attr_set = true;
    }
    FusedArgs fa;
    fa.l0 = a0;
    fa.l0.ncomp_total = a0.ncomp;
    fa.bands1 = bands1;
    fa.ll1 = ll1;
    int comps[SB2_MAX_COMPONENTS];
    for (int c = 0; c < a0.ncomp; c++) comps[c] = c;
    const dim3 grid = level_grid (fa.l0, comps, a0.ncomp, count);
    char tag[48] = "";
    double bytes = 0;
    if (profiling ()) {
      snprintf (tag, sizeof (tag), "wavelet_inv_%s_f%d_w%d_fused2", sizeof (T) == 4 ? "s32" : "s16", F, a0.w[0]);
      // levels 1 and 0 together: every coefficient of the two levels read once, the picture written once
      for (int c = 0; c < a0.ncomp; c++) bytes += 2.0 * a0.w[c] * a0.h[c] * sizeof (T) * count;
    }
    {
      LaunchScope scope (tag, bytes, stream);
      wavelet_inv_fused2_kernel<T, F, 16><<<grid, FG::NT, U::SMEM, stream>>> (fa);
    }
    return check_cuda (cudaGetLastError (), "wavelet_inv_fused2_kernel launch");
  }
  return set_error (SB2_ERR_UNSUPPORTED, "no fused kernel for filter %d", F);
}

// level 0 of the inverse with the combine fused in (u8 out); every component must take the register-chunk kernel
template <typename T, int F>
static int launch_inverse_level0_u8 (LevelArgs a, int count, cudaStream_t stream)
{
  // the chunk choice does not look at `dense` alignment here (the u8 stores check their own), only at the sizes
  int sel[3][SB2_MAX_COMPONENTS], nsel[3] = { 0, 0, 0 };
  for (int c = 0; c < a.ncomp; c++) {
    const int n = a.w[c] >> 1, m = a.h[c] >> 1;
    const int k = (n % 16 == 0 && m % 16 == 0 && filter_halo (F) <= 4) ? 1 : (n % 8 == 0 && m % 8 == 0) ? 2 : 0;
    if (!k) return set_error (SB2_ERR_UNSUPPORTED, "fused inverse + convert needs planes whose half sizes are multiples of 8");
    sel[k][nsel[k]++] = c;
  }
  a.ncomp_total = a.ncomp;
  for (int k = 1; k < 3; k++) {
    if (!nsel[k]) continue;
    LevelArgs b = a;
    const dim3 grid = level_grid (b, sel[k], nsel[k], count);
    char tag[48] = "";
    double bytes = 0;
    if (profiling ()) {
      snprintf (tag, sizeof (tag), "wavelet_inv_%s_f%d_w%d_u8", sizeof (T) == 4 ? "s32" : "s16", F, a.w[sel[k][0]]);
      for (int i = 0; i < nsel[k]; i++) bytes += (double) a.w[sel[k][i]] * a.h[sel[k][i]] * (sizeof (T) + 1) * count;
    }
    cudaError_t e;
    if (k == 1) {
      typedef FastGeom<T, F, 16> FG;
      e = cudaFuncSetAttribute (wavelet_inv_fast_kernel<T, F, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FG::SMEM);
      if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet u8)");
      LaunchScope scope (tag, bytes, stream);
      wavelet_inv_fast_kernel<T, F, 16, true><<<grid, FG::NT, FG::SMEM, stream>>> (b);
    } else {
      typedef FastGeom<T, F, 8> FG;
      e = cudaFuncSetAttribute (wavelet_inv_fast_kernel<T, F, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FG::SMEM);
      if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet u8)");
      LaunchScope scope (tag, bytes, stream);
      wavelet_inv_fast_kernel<T, F, 8, true><<<grid, FG::NT, FG::SMEM, stream>>> (b);
    }
    e = cudaGetLastError ();
    if (e != cudaSuccess) return check_cuda (e, "wavelet_inv_fast_kernel<u8> launch");
  }
  return SB2_OK;
}

template <typename T>
static int launch_inverse_level0_u8_f (int filter, const LevelArgs &a, int count, cudaStream_t s)
{
  switch (filter) {
    case 0: return launch_inverse_level0_u8<T, 0> (a, count, s);
    case 1: return launch_inverse_level0_u8<T, 1> (a, count, s);
    case 2: return launch_inverse_level0_u8<T, 2> (a, count, s);
    case 3: return launch_inverse_level0_u8<T, 3> (a, count, s);
    case 4: return launch_inverse_level0_u8<T, 4> (a, count, s);
    case 5: return launch_inverse_level0_u8<T, 5> (a, count, s);
    default: return launch_inverse_level0_u8<T, 6> (a, count, s);
  }
}

template <typename T, int F, int CS>
static int launch_fast_forward (LevelArgs a, const int *comps, int nsel, int count, cudaStream_t stream,
    const char *tag, double bytes)
{
  typedef FwdGeom<T, F, CS> FG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute (wavelet_fwd_fast_kernel<T, F, CS>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FG::SMEM);
    if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet fast forward)");
    attr_set = true;
  }
  const dim3 grid = level_grid (a, comps, nsel, count);
  {
    LaunchScope scope (tag, bytes, stream);
    wavelet_fwd_fast_kernel<T, F, CS><<<grid, FG::NT, FG::SMEM, stream>>> (a);
  }
  return check_cuda (cudaGetLastError (), "wavelet_fwd_fast_kernel launch");
}

static int g_iwt_generic = -1;
static bool iwt_generic_forced ()
{
  if (g_iwt_generic < 0) {
    const char *v = getenv ("SB2_IWT_GENERIC");
    g_iwt_generic = (v && atoi (v)) ? 1 : 0;
  }
  return g_iwt_generic != 0;
}

template <typename T, int F, bool INV>
static int launch_level (const LevelArgs &a_in, int count, cudaStream_t stream)
{
  typedef TileGeom<T, F> G;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute (wavelet_level_kernel<T, F, INV>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int) G::SMEM);
    if (e != cudaSuccess) return check_cuda (e, "cudaFuncSetAttribute(wavelet)");
    attr_set = true;
  }
  LevelArgs a = a_in;
  a.ncomp_total = a_in.ncomp;
  // components are routed to the fastest kernel their size allows
  int sel[3][SB2_MAX_COMPONENTS], nsel[3] = { 0, 0, 0 };     // 0: generic, 1: chunk 16, 2: chunk 8
  for (int c = 0; c < a_in.ncomp; c++) {
    if ((a.w[c] >> 1) == 0 || (a.h[c] >> 1) == 0) continue;
    int cs = 0;
    if (!iwt_generic_forced ()) cs = fast_inverse_chunk<T, F> (a, c);      // the same constraints serve the forward kernel (dense = its input)
    const int k = cs == 16 ? 1 : cs == 8 ? 2 : 0;
    sel[k][nsel[k]++] = c;
  }
  for (int k = 0; k < 3; k++) {
    if (!nsel[k]) continue;
    char tag[48] = "";
    double bytes = 0;
    if (profiling ()) {
      snprintf (tag, sizeof (tag), "wavelet_%s_%s_f%d_w%d%s", INV ? "inv" : "fwd",
          sizeof (T) == 4 ? "s32" : "s16", F, a.w[sel[k][0]], k == 0 ? "_generic" : "");
      for (int i = 0; i < nsel[k]; i++) bytes += 2.0 * a.w[sel[k][i]] * a.h[sel[k][i]] * sizeof (T) * count;
    }
    int rc = SB2_OK;
    if constexpr (INV) {
      if (k == 1) rc = launch_fast_inverse<T, F, 16> (a, sel[k], nsel[k], count, stream, tag, bytes);
      if (k == 2) rc = launch_fast_inverse<T, F, 8> (a, sel[k], nsel[k], count, stream, tag, bytes);
    } else {
      if (k == 1) rc = launch_fast_forward<T, F, 16> (a, sel[k], nsel[k], count, stream, tag, bytes);
      if (k == 2) rc = launch_fast_forward<T, F, 8> (a, sel[k], nsel[k], count, stream, tag, bytes);
    }
    if (k == 0) {
      const dim3 grid = level_grid (a, sel[0], nsel[0], count);
      {
        LaunchScope scope (tag, bytes, stream);
        wavelet_level_kernel<T, F, INV><<<grid, NTHREADS, G::SMEM, stream>>> (a);
      }
      rc = check_cuda (cudaGetLastError (), "wavelet_level_kernel launch");
    }
    if (rc) return rc;
  }
  return SB2_OK;
}

template <typename T, bool INV>
static int launch_level_f (int filter, const LevelArgs &a, int count, cudaStream_t s)
{
  switch (filter) {
    case 0: return launch_level<T, 0, INV> (a, count, s);
    case 1: return launch_level<T, 1, INV> (a, count, s);
    case 2: return launch_level<T, 2, INV> (a, count, s);
    case 3: return launch_level<T, 3, INV> (a, count, s);
    case 4: return launch_level<T, 4, INV> (a, count, s);
    case 5: return launch_level<T, 5, INV> (a, count, s);
    case 6: return launch_level<T, 6, INV> (a, count, s);
    default: return set_error (SB2_ERR_ARG, "bad wavelet filter index %d", filter);
  }
}

static int launch_level_any (bool inv, int is_s32, int filter, const LevelArgs &a,
    int count, cudaStream_t s)
{
  if (is_s32)
    return inv ? launch_level_f<int32_t, true> (filter, a, count, s)
               : launch_level_f<int32_t, false> (filter, a, count, s);
  return inv ? launch_level_f<int16_t, true> (filter, a, count, s)
             : launch_level_f<int16_t, false> (filter, a, count, s);
}

template <typename T>
static bool fused2_supported_f (int filter, const LevelArgs &a0)
{
  switch (filter) {
    case 0: return fused2_supported<T, 0> (a0);
    case 2: return fused2_supported<T, 2> (a0);
    case 6: return fused2_supported<T, 6> (a0);
    default: return false;
  }
}
template <typename T>
static int launch_fused2_f (int filter, const LevelArgs &a0, const PlaneSet &bands1, const PlaneSet &ll1, int count, cudaStream_t s)
{
  switch (filter) {
    case 0: return launch_fused2<T, 0> (a0, bands1, ll1, count, s);
    case 2: return launch_fused2<T, 2> (a0, bands1, ll1, count, s);
    default: return launch_fused2<T, 6> (a0, bands1, ll1, count, s);
  }
}

// Workspace: per picture, per component: T1 (w/2 x h/2), T0 (w/4 x h/4) and, for
// in-place calls, a full-size plane F.  Strides are rounded up to 16 bytes.
struct WsLayout {
  size_t pic_pitch;
  size_t off_t1[SB2_MAX_COMPONENTS], off_t0[SB2_MAX_COMPONENTS], off_full[SB2_MAX_COMPONENTS];
  int stride_t1[SB2_MAX_COMPONENTS], stride_t0[SB2_MAX_COMPONENTS], stride_full[SB2_MAX_COMPONENTS];
};

static inline size_t round_up (size_t x, size_t a) { return (x + a - 1) / a * a; }

static WsLayout ws_layout (const sb2_slab *s, int bpp, int depth, int in_place)
{
  WsLayout L;
  size_t off = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    L.off_t1[c] = L.off_t0[c] = L.off_full[c] = 0;
    L.stride_t1[c] = L.stride_t0[c] = L.stride_full[c] = 0;
    if (c >= s->ncomp) continue;
    const int w = s->width[c], h = s->height[c];
    if (depth > 1) {
      L.stride_t1[c] = (int) round_up ((size_t) (w / 2) * bpp, 16);
      L.off_t1[c] = off;
      off += round_up ((size_t) L.stride_t1[c] * (h / 2), 256);
    }
    if (depth > 2) {
      L.stride_t0[c] = (int) round_up ((size_t) (w / 4) * bpp, 16);
      L.off_t0[c] = off;
      off += round_up ((size_t) L.stride_t0[c] * (h / 4), 256);
    }
    if (in_place) {
      L.stride_full[c] = (int) round_up ((size_t) w * bpp, 16);
      L.off_full[c] = off;
      off += round_up ((size_t) L.stride_full[c] * h, 256);
    }
  }
  L.pic_pitch = round_up (off, 256);
  return L;
}

static int validate (const sb2_slab *src, const sb2_slab *dst, int bpp, int depth)
{
  if (!src || !dst || !src->base || !dst->base)
    return set_error (SB2_ERR_ARG, "null slab");
  if (src->ncomp < 1 || src->ncomp > SB2_MAX_COMPONENTS || src->ncomp != dst->ncomp ||
      src->count != dst->count || src->count < 1)
    return set_error (SB2_ERR_ARG, "slab shapes differ (ncomp %d/%d count %d/%d)",
        src->ncomp, dst->ncomp, src->count, dst->count);
  if (depth < 1 || depth > 8) return set_error (SB2_ERR_ARG, "bad transform depth %d", depth);
  if (src->count > 65535) return set_error (SB2_ERR_ARG, "at most 65535 pictures per call (%d)", src->count);
  for (int c = 0; c < src->ncomp; c++) {
    if (src->width[c] != dst->width[c] || src->height[c] != dst->height[c])
      return set_error (SB2_ERR_ARG, "component %d size differs", c);
    if (src->width[c] <= 0 || src->height[c] <= 0 ||
        (src->width[c] & ((1 << depth) - 1)) || (src->height[c] & ((1 << depth) - 1)))
      return set_error (SB2_ERR_ARG, "component %d size %dx%d is not a multiple of 1<<%d",
          c, src->width[c], src->height[c], depth);
    if ((src->stride[c] % bpp) || (dst->stride[c] % bpp) || (src->offset[c] % bpp) ||
        (dst->offset[c] % bpp))
      return set_error (SB2_ERR_ARG, "component %d stride/offset not a multiple of the sample size", c);
  }
  return SB2_OK;
}

static bool slabs_alias (const sb2_slab *a, const sb2_slab *b)
{
  // conservative: any overlap of the two address ranges counts as aliasing
  const char *a0 = (const char *) a->base, *b0 = (const char *) b->base;
  const char *a1 = a0 + a->picture_pitch * (size_t) a->count;
  const char *b1 = b0 + b->picture_pitch * (size_t) b->count;
  if (a->count == 1) {
    size_t ext = 0;
    for (int c = 0; c < a->ncomp; c++)
      ext = max (ext, a->offset[c] + (size_t) a->stride[c] * a->height[c]);
    a1 = a0 + ext;
  }
  if (b->count == 1) {
    size_t ext = 0;
    for (int c = 0; c < b->ncomp; c++)
      ext = max (ext, b->offset[c] + (size_t) b->stride[c] * b->height[c]);
    b1 = b0 + ext;
  }
  return a0 < b1 && b0 < a1;
}

static PlaneSet ws_planeset (void *ws, const WsLayout &L, const size_t *off, const int *stride)
{
  PlaneSet p;
  p.base = (char *) ws;
  p.pic_pitch = L.pic_pitch;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) { p.off[c] = off[c]; p.stride[c] = stride[c]; }
  return p;
}

static PlaneSet scaled (const PlaneSet &p, int shift)
{
  PlaneSet q = p;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) q.stride[c] = p.stride[c] << shift;
  return q;
}

static int copy_back (const sb2_slab *dst, void *ws, const WsLayout &L, int bpp, cudaStream_t st)
{
  for (int p = 0; p < dst->count; p++) {
    for (int c = 0; c < dst->ncomp; c++) {
      cudaError_t e = cudaMemcpy2DAsync ((char *) dst->base + (size_t) p * dst->picture_pitch + dst->offset[c],
          dst->stride[c], (char *) ws + (size_t) p * L.pic_pitch + L.off_full[c], L.stride_full[c],
          (size_t) dst->width[c] * bpp, dst->height[c], cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda (e, "cudaMemcpy2DAsync(copy back)");
    }
  }
  return SB2_OK;
}

static int iwt_run (bool inv, const sb2_slab *src, const sb2_slab *dst, int is_s32, int filter,
    int depth, void *workspace, size_t workspace_bytes, void *stream)
{
  const int bpp = is_s32 ? 4 : 2;
  int rc = validate (src, dst, bpp, depth);
  if (rc) return rc;
  if (filter < 0 || filter > 6) return set_error (SB2_ERR_ARG, "bad wavelet filter index %d", filter);
  const bool in_place = slabs_alias (src, dst);
  const WsLayout L = ws_layout (src, bpp, depth, in_place);
  const size_t need = L.pic_pitch * (size_t) src->count;
  if (need > 0 && (!workspace || workspace_bytes < need))
    return set_error (SB2_ERR_WORKSPACE, "wavelet workspace too small: need %zu bytes, have %zu",
        need, workspace ? workspace_bytes : (size_t) 0);
  cudaStream_t st = as_stream (stream);

  const PlaneSet S = planeset_from_slab (src);
  const PlaneSet D = planeset_from_slab (dst);
  const PlaneSet T1 = ws_planeset (workspace, L, L.off_t1, L.stride_t1);
  const PlaneSet T0 = ws_planeset (workspace, L, L.off_t0, L.stride_t0);
  const PlaneSet FULL = ws_planeset (workspace, L, L.off_full, L.stride_full);

  LevelArgs a;
  a.ncomp = src->ncomp;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) a.w[c] = a.h[c] = 0;

  if (!inv) {
    // level l: X = (l==0 ? src : T[l&1]);  bands -> dst (stride<<l);
    // LL -> (l==depth-1 ? dst (stride<<(l+1)) : T[(l+1)&1])
    PlaneSet in0 = S;
    if (in_place) {
      // stage the input so level 0 can write its bands into dst == src
      for (int p = 0; p < src->count; p++)
        for (int c = 0; c < src->ncomp; c++) {
          cudaError_t e = cudaMemcpy2DAsync ((char *) workspace + (size_t) p * L.pic_pitch + L.off_full[c],
              L.stride_full[c], (char *) src->base + (size_t) p * src->picture_pitch + src->offset[c],
              src->stride[c], (size_t) src->width[c] * bpp, src->height[c], cudaMemcpyDeviceToDevice, st);
          if (e != cudaSuccess) return check_cuda (e, "cudaMemcpy2DAsync(stage in)");
        }
      in0 = FULL;
    }
    for (int l = 0; l < depth; l++) {
      for (int c = 0; c < src->ncomp; c++) { a.w[c] = src->width[c] >> l; a.h[c] = src->height[c] >> l; }
      a.dense = (l == 0) ? in0 : ((l & 1) ? T1 : T0);
      a.bands = scaled (D, l);
      a.ll = (l == depth - 1) ? scaled (D, l + 1) : (((l + 1) & 1) ? T1 : T0);
      rc = launch_level_any (false, is_s32, filter, a, src->count, st);
      if (rc) return rc;
    }
    return SB2_OK;
  }

  // inverse: level l = depth-1 .. 0
  //   LL <- (l==depth-1 ? src (stride<<(l+1)) : T[(l+1)&1]);  bands <- src (stride<<l);
  //   Y  -> (l==0 ? dst (or FULL when in place) : T[l&1])
  // levels 1 and 0 run as one fused launch when every component's shape allows (level 1's output then
  // never goes through HBM); the deeper levels, and everything else, run one launch per level
  bool fused = false;
  if (depth >= 2 && iwt_fused_enabled () && !iwt_generic_forced ()) {
    for (int c = 0; c < src->ncomp; c++) { a.w[c] = src->width[c]; a.h[c] = src->height[c]; }
    a.bands = S;
    a.ll = T1;
    a.dense = in_place ? FULL : D;
    fused = is_s32 ? fused2_supported_f<int32_t> (filter, a) : fused2_supported_f<int16_t> (filter, a);
  }
  for (int l = depth - 1; l >= (fused ? 2 : 0); l--) {
    for (int c = 0; c < src->ncomp; c++) { a.w[c] = src->width[c] >> l; a.h[c] = src->height[c] >> l; }
    a.ll = (l == depth - 1) ? scaled (S, l + 1) : (((l + 1) & 1) ? T1 : T0);
    a.bands = scaled (S, l);
    a.dense = (l == 0) ? (in_place ? FULL : D) : ((l & 1) ? T1 : T0);
    rc = launch_level_any (true, is_s32, filter, a, src->count, st);
    if (rc) return rc;
  }
  if (fused) {
    for (int c = 0; c < src->ncomp; c++) { a.w[c] = src->width[c]; a.h[c] = src->height[c]; }
    a.bands = S;
    a.ll = T1;                                     // unused by the fused kernel
    a.dense = in_place ? FULL : D;
    const PlaneSet ll1 = depth == 2 ? scaled (S, 2) : T0;
    rc = is_s32 ? launch_fused2_f<int32_t> (filter, a, scaled (S, 1), ll1, src->count, st)
                : launch_fused2_f<int16_t> (filter, a, scaled (S, 1), ll1, src->count, st);
    if (rc) return rc;
  }
  if (in_place) return copy_back (dst, workspace, L, bpp, st);
  return SB2_OK;
}

// inverse transform with the combine fused into its last level: levels depth-1 .. 1 as iwt_run does them, level 0
// through the OUT8 kernels straight into the u8 picture
static int iwt_inverse_convert_run (const sb2_slab *src, const sb2_slab *dst, int is_s32, int filter, int depth, int shift,
    void *workspace, size_t workspace_bytes, void *stream)
{
  const int bpp = is_s32 ? 4 : 2;
  if (!src || !dst || !src->base || !dst->base) return set_error (SB2_ERR_ARG, "null slab");
  if (src->ncomp < 1 || src->ncomp > SB2_MAX_COMPONENTS || src->ncomp != dst->ncomp || src->count != dst->count || src->count < 1 ||
      src->count > 65535)
    return set_error (SB2_ERR_ARG, "sb2_iwt_inverse_convert: slab shapes differ");
  if (depth < 1 || depth > 8 || filter < 0 || filter > 6 || shift < 0 || shift > (is_s32 ? 31 : 15))
    return set_error (SB2_ERR_ARG, "sb2_iwt_inverse_convert: depth %d / filter %d / shift %d", depth, filter, shift);
  for (int c = 0; c < src->ncomp; c++) {
    if (src->width[c] <= 0 || src->height[c] <= 0 || (src->width[c] & ((1 << depth) - 1)) || (src->height[c] & ((1 << depth) - 1)))
      return set_error (SB2_ERR_ARG, "component %d size %dx%d is not a multiple of 1<<%d", c, src->width[c], src->height[c], depth);
    if (dst->width[c] < 1 || dst->height[c] < 1 || dst->width[c] > src->width[c] || dst->height[c] > src->height[c])
      return set_error (SB2_ERR_ARG, "component %d: the picture (%dx%d) must lie inside the transform's area (%dx%d)", c,
          dst->width[c], dst->height[c], src->width[c], src->height[c]);
    if ((src->stride[c] % bpp) || (src->offset[c] % bpp)) return set_error (SB2_ERR_ARG, "component %d stride/offset not a multiple of the sample size", c);
  }
  const WsLayout L = ws_layout (src, bpp, depth, 0);
  const size_t need = L.pic_pitch * (size_t) src->count;
  if (need > 0 && (!workspace || workspace_bytes < need))
    return set_error (SB2_ERR_WORKSPACE, "wavelet workspace too small: need %zu bytes, have %zu", need, workspace ? workspace_bytes : (size_t) 0);
  cudaStream_t st = as_stream (stream);
  const PlaneSet S = planeset_from_slab (src);
  const PlaneSet T1 = ws_planeset (workspace, L, L.off_t1, L.stride_t1);
  const PlaneSet T0 = ws_planeset (workspace, L, L.off_t0, L.stride_t0);
  LevelArgs a;
  a.ncomp = src->ncomp;
  a.out_shift = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) a.w[c] = a.h[c] = a.out_w[c] = a.out_h[c] = 0;
  for (int l = depth - 1; l >= 1; l--) {
    for (int c = 0; c < src->ncomp; c++) { a.w[c] = src->width[c] >> l; a.h[c] = src->height[c] >> l; }
    a.ll = (l == depth - 1) ? scaled (S, l + 1) : (((l + 1) & 1) ? T1 : T0);
    a.bands = scaled (S, l);
    a.dense = (l & 1) ? T1 : T0;
    const int rc = launch_level_any (true, is_s32, filter, a, src->count, st);
    if (rc) return rc;
  }
  for (int c = 0; c < src->ncomp; c++) {
    a.w[c] = src->width[c]; a.h[c] = src->height[c];
    a.out_w[c] = dst->width[c]; a.out_h[c] = dst->height[c];
  }
  a.ncomp = src->ncomp;
  a.ll = depth == 1 ? scaled (S, 1) : T1;
  a.bands = S;
  a.dense = planeset_from_slab (dst);
  a.out_shift = shift;
  return is_s32 ? launch_inverse_level0_u8_f<int32_t> (filter, a, src->count, st)
                : launch_inverse_level0_u8_f<int16_t> (filter, a, src->count, st);
}

}  // namespace sb2

extern "C" int
sb2_iwt_inverse_convert (const sb2_slab *src, const sb2_slab *dst_u8, int is_s32, int filter, int depth, int shift,
    void *workspace, size_t workspace_bytes, void *stream)
{
  return sb2::iwt_inverse_convert_run (src, dst_u8, is_s32, filter, depth, shift, workspace, workspace_bytes, stream);
}

extern "C" size_t
sb2_iwt_workspace_bytes (const sb2_slab *slab, int is_s32, int depth, int in_place)
{
  if (!slab || slab->ncomp < 1 || slab->ncomp > SB2_MAX_COMPONENTS) return 0;
  const sb2::WsLayout L = sb2::ws_layout (slab, is_s32 ? 4 : 2, depth, in_place);
  return L.pic_pitch * (size_t) slab->count;
}

extern "C" int
sb2_iwt_forward (const sb2_slab *src, const sb2_slab *dst, int is_s32, int filter, int depth,
    void *workspace, size_t workspace_bytes, void *stream)
{
  return sb2::iwt_run (false, src, dst, is_s32, filter, depth, workspace, workspace_bytes, stream);
}

extern "C" int
sb2_iwt_inverse (const sb2_slab *src, const sb2_slab *dst, int is_s32, int filter, int depth,
    void *workspace, size_t workspace_bytes, void *stream)
{
  return sb2::iwt_run (true, src, dst, is_s32, filter, depth, workspace, workspace_bytes, stream);
}

extern "C" void sb2_iwt_force_generic (int on) { sb2::g_iwt_generic = on ? 1 : 0; }
extern "C" void sb2_iwt_enable_fused (int on) { sb2::g_iwt_fused = on ? 1 : 0; }
