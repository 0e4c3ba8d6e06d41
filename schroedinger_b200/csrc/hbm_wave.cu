// hbm_wave.cu -- hierarchical block matching for the codec's usual geometry (8x8 blocks,
// 4:2:0, luma-only scan): a skewed multi-row wavefront with register hand-off.
//
// Same result, bit for bit, as one level of schro_hierarchical_bm_scan_hint
// (schroedinger/schrohierbm.c:174-383) with schro_metric_fast_block / _scan_setup / _do_scan /
// _get_min (schroedinger/schrometric.c:31-214, 332-414); hbm.cu holds the generic kernel for
// every other geometry and the C entry point.
//
// A block depends on its left, upper and upper-left neighbours of the same level.  Here one
// WARP owns RPW = 32 / G adjacent block rows of one (picture, reference) pair, G lanes per
// row, and walks them skewed by one block: at step t row q works on column t - q.  The
// neighbour vectors of rows 1..RPW-1 are the registers of the G lanes above, one shuffle away;
// only a warp's top row polls a word published by the warp above, and that word is fetched
// one step ahead.  A launch is one warp per CTA; CTAs take (row group, picture) from an
// atomic ticket, so a CTA only ever waits on CTAs that have already started.
//
// Per block: lanes 0..5 hold the static candidates (zero vector + five parents) whose ranking
// SADs a dependency-free pre-pass (hbm_static_kernel) computed for every block of the level;
// lanes 6 / 7 hold left / up, up-left is group-uniform.  Duplicates are found with one MATCH;
// a neighbour that repeats a static candidate re-uses its SAD, otherwise the group's first
// eight lanes compute it one block row per lane.  The scan is split into tasks of one window
// column x up to seven window rows: a lane streams the 14 reference rows of its column once
// (three aligned words + two funnel shifts per row) and feeds seven running SADs with
// VABSDIFF4.ACC -- 26 instructions per position instead of 80 when every position loads its
// own block.  Partial blocks at the right / bottom picture edge use byte masks and row counts
// in the same code.

#include "hbm_common.cuh"

namespace sb2 {

#define SB2_FULL 0xffffffffu
static constexpr unsigned STAT_INVALID = 0xffffffffu;

#ifndef SB2_WAVE_POLL_NS
#define SB2_WAVE_POLL_NS 32
#endif

#ifdef SB2_HBM_TRACE
__device__ long long g_wave_trace[1024 * 8];
__device__ int g_wave_trace_sel[2] = { 0, 10 };          // level, row group
#define WTRACE(k) do { if (wtrace && lane == 0 && t < 1024) g_wave_trace[t * 8 + (k)] = clock64 (); } while (0)
#else
#define WTRACE(k) do { } while (0)
#endif

struct WavePlanes {
  const uint8_t *sp[3], *rp[3];
  int ss[3], rs[3];
};

struct WaveBlock {
  int x0, y0;
  int bw0, hl, hc;                  // luma width / rows, chroma rows of this block
  unsigned mlo, mhi, cm;            // byte masks: luma bytes 0-3, 4-7, chroma bytes 0-3
};

__device__ __forceinline__ WavePlanes wave_planes (const HbmArgs &A, int pic)
{
  WavePlanes P;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    P.sp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (A.src, pic, k));
    P.rp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref, pic, k));
    P.ss[k] = A.src.stride[k];
    P.rs[k] = A.ref.stride[k];
  }
  return P;
}

__device__ __forceinline__ unsigned low_bytes_mask (int n)      // n bytes set from the low end, n in 0..4
{
  return n >= 4 ? 0xffffffffu : n <= 0 ? 0u : ((1u << (8 * n)) - 1u);
}

__device__ __forceinline__ WaveBlock wave_block (const HbmArgs &A, int x0, int y0)
{
  WaveBlock B;
  B.x0 = x0;
  B.y0 = y0;
  B.bw0 = min (A.width - x0, 8);
  B.hl = max (0, min (A.height - y0, 8));
  B.hc = max (0, min (A.ch - (y0 >> 1), 4));
  B.mlo = low_bytes_mask (B.bw0);
  B.mhi = low_bytes_mask (B.bw0 - 4);
  B.cm = low_bytes_mask (min (A.cw - (x0 >> 1), 4));
  return B;
}

// Ranking SADs (schro_metric_block_sad_slow, schrometric.c:332-375) of up to NV candidate vectors
// at once by the first eight lanes of a group: lane y takes luma row y, lanes 0-3 also a U row,
// lanes 4-7 a V row.  Straight-line code: a load that is switched off reads a harmless address
// instead of sitting behind a branch, every load is issued before the first use (the vectors
// share one memory round trip), and the warp barrier keeps the compiler from moving a consumer
// between the loads.  Every lane of the warp calls it; want[k] says whether the group needs
// vector k.  The sums come back in all eight lanes (two 16-bit sums ride in one shuffle chain:
// a three-component 8x8 SAD is at most 24480).
template <int NV>
__device__ __forceinline__ void wave_rank_sads (const HbmArgs &A, const WavePlanes &P, const WaveBlock &B,
    const int (&vec)[NV], const bool (&want)[NV], int sub, unsigned (&out)[NV])
{
  const int s = A.shift;
  bool any = false;
#pragma unroll
  for (int k = 0; k < NV; k++) any = any || want[k];
  const bool lrow = any && sub < B.hl;
  const int cr = sub & 3;
  const bool crow = any && sub < 8 && cr < B.hc;
  const bool isv = (sub >> 2) != 0;          // no dynamic indexing: the plane table stays in registers
  const uint8_t *spk = isv ? P.sp[2] : P.sp[1], *rpk = isv ? P.rp[2] : P.rp[1];
  const int ssk = isv ? P.ss[2] : P.ss[1], rsk = isv ? P.rs[2] : P.rs[1];
  const uint8_t *safe = P.rp[0];             // pixel (0,0) of the reference: aligned, always readable
  const uint2 a = __ldg (reinterpret_cast<const uint2 *> (lrow ? P.sp[0] + (ptrdiff_t) (B.y0 + sub) * P.ss[0] + B.x0 : safe));
  const unsigned ac = __ldg (reinterpret_cast<const unsigned *> (crow ? spk + (ptrdiff_t) ((B.y0 >> 1) + cr) * ssk + (B.x0 >> 1) : safe));
  unsigned w[NV][5];
  unsigned sh[NV], shc[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) {
    int dx = (vec[k] >> 16) >> s, dy = ((int) (short) vec[k]) >> s;
    dx = clampi (dx + B.x0, -B.bw0, A.width) - B.x0;
    dy = clampi (dy + B.y0, -B.hl, A.height) - B.y0;
    const RowRef rl = row_ref (lrow && want[k] ? P.rp[0] + (ptrdiff_t) (B.y0 + dy + sub) * P.rs[0] + B.x0 + dx : safe);
    const RowRef rc = row_ref (crow && want[k] ? rpk + (ptrdiff_t) (((B.y0 + dy) >> 1) + cr) * rsk + ((B.x0 + dx) >> 1) : safe);
    w[k][0] = __ldg (rl.w);
    w[k][1] = __ldg (rl.w + 1);
    w[k][2] = rl.three ? __ldg (rl.w + 2) : 0u;
    w[k][3] = __ldg (rc.w);
    w[k][4] = rc.three ? __ldg (rc.w + 1) : 0u;
    sh[k] = rl.sh;
    shc[k] = rc.sh;
  }
  __syncwarp ();
  unsigned part[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const unsigned bx = __funnelshift_r (w[k][0], w[k][1], sh[k]), by = __funnelshift_r (w[k][1], w[k][2], sh[k]);
    const unsigned bc = __funnelshift_r (w[k][3], w[k][4], shc[k]);
    unsigned v = 0;
    if (lrow) v = __vsadu4 (a.x & B.mlo, bx & B.mlo) + __vsadu4 (a.y & B.mhi, by & B.mhi);
    if (crow) v += __vsadu4 (ac & B.cm, bc & B.cm);
    part[k] = want[k] ? v : 0u;          // (loads of a switched-off vector came from `safe`)
  }
#pragma unroll
  for (int k = 0; k + 1 < NV; k += 2) part[k] |= part[k + 1] << 16;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
    for (int k = 0; k < NV; k += 2) part[k] += __shfl_xor_sync (SB2_FULL, part[k], o);
  }
#pragma unroll
  for (int k = 0; k < NV; k++) out[k] = (k & 1) ? (part[k - 1] >> 16) : ((k + 1 < NV) ? (part[k] & 0xffffu) : part[k]);
}

// ---- pre-pass: the candidates that do not depend on this level's neighbours ------------------
// (0: zero vector, 1..5: parents (0,0) (-1,0) (1,0) (0,-1) (0,1), schrohierbm.c:259-277) with
// their ranking SADs (schrometric.c:332-375), for every block of the level:
// stat[block][k] = (vector, SAD | INVALID).  One thread per block, consecutive threads =
// consecutive blocks of a row, so a warp's loads of one pixel row are contiguous (source) or
// nearly so (reference, displaced by similar vectors); a vector that repeats an earlier
// candidate of the block copies its SAD.
__global__ void __launch_bounds__ (128)
hbm_static_kernel (const HbmArgs A, uint2 *stat)
{
  const long long g = (long long) blockIdx.x * 128 + threadIdx.x;
  if (g >= (long long) A.count * A.rows * A.cols) return;
  const int col = (int) (g % A.cols);
  const long long t = g / A.cols;
  const int row = (int) (t % A.rows), pic = (int) (t / A.rows);
  const int s = A.shift, skip = 1 << s, ri = A.ref_index;
  const int i = col * skip, j = row * skip;
  const WaveBlock B = wave_block (A, col * 8, row * 8);
  const WavePlanes P = wave_planes (A, pic);
  const bool act = B.x0 < A.width && B.y0 < A.height;
  const uint8_t *safe = P.rp[0];

  uint2 sl[8];
  unsigned sc[2][4];
#pragma unroll
  for (int y = 0; y < 8; y++) {
    sl[y] = __ldg (reinterpret_cast<const uint2 *> (act && y < B.hl ? P.sp[0] + (ptrdiff_t) (B.y0 + y) * P.ss[0] + B.x0 : safe));
    sl[y].x &= B.mlo;
    sl[y].y &= B.mhi;
  }
#pragma unroll
  for (int cp = 0; cp < 2; cp++)
#pragma unroll
    for (int y = 0; y < 4; y++)
      sc[cp][y] = __ldg (reinterpret_cast<const unsigned *> (act && y < B.hc
              ? P.sp[1 + cp] + (ptrdiff_t) ((B.y0 >> 1) + y) * P.ss[1 + cp] + (B.x0 >> 1) : safe)) & B.cm;

  int vec[6];
  unsigned met[6];
  bool valid[6];
  const int hint_mask = ~((1 << (s + 1)) - 1);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    vec[k] = 0;
    valid[k] = act && k == 0;
    met[k] = STAT_INVALID;
    if (k > 0 && act && A.parent) {
      const int ox = (k == 2) ? -1 : (k == 3) ? 1 : 0;
      const int oy = (k == 4) ? -1 : (k == 5) ? 1 : 0;
      const int ll = (i & hint_mask) + ox * skip * 2, kk = (j & hint_mask) + oy * skip * 2;
      if (ll >= 0 && ll < A.nbx && kk >= 0 && kk < A.nby) {
        const MotionVector *m = A.parent + (size_t) pic * A.field_pitch + (size_t) kk * A.nbx + ll;
        vec[k] = (int) (((unsigned) (unsigned short) m->v[ri] << 16) | (unsigned short) m->v[2 + ri]);
        valid[k] = true;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    bool dup = false;
#pragma unroll
    for (int e = 0; e < k; e++)
      if (valid[e] && vec[e] == vec[k]) { dup = true; met[k] = met[e]; }
    if (valid[k] && !dup) {
      int dx = (vec[k] >> 16) >> s, dy = ((int) (short) vec[k]) >> s;
      dx = clampi (dx + B.x0, -B.bw0, A.width) - B.x0;
      dy = clampi (dy + B.y0, -B.hl, A.height) - B.y0;
      unsigned m = 0;
      const RowRef rb = row_ref (P.rp[0] + (ptrdiff_t) (B.y0 + dy) * P.rs[0] + B.x0 + dx);
      const int rsw = P.rs[0] >> 2;
#pragma unroll
      for (int y = 0; y < 8; y++)
        if (y < B.hl) {
          const uint2 bv = row_load8 (rb, y * rsw);
          m += __vsadu4 (sl[y].x, bv.x & B.mlo) + __vsadu4 (sl[y].y, bv.y & B.mhi);
        }
#pragma unroll
      for (int cp = 0; cp < 2; cp++) {
        const RowRef rc = row_ref (P.rp[1 + cp] + (ptrdiff_t) ((B.y0 + dy) >> 1) * P.rs[1 + cp] + ((B.x0 + dx) >> 1));
        const int rcw = P.rs[1 + cp] >> 2;
#pragma unroll
        for (int y = 0; y < 4; y++)
          if (y < B.hc) m += __vsadu4 (sc[cp][y], row_load4 (rc, y * rcw) & B.cm);
      }
      met[k] = m;
    }
    if (!valid[k]) met[k] = STAT_INVALID;
  }
  uint4 *o = reinterpret_cast<uint4 *> (stat + g * 8);
  o[0] = make_uint4 ((unsigned) vec[0], met[0], (unsigned) vec[1], met[1]);
  o[1] = make_uint4 ((unsigned) vec[2], met[2], (unsigned) vec[3], met[3]);
  o[2] = make_uint4 ((unsigned) vec[4], met[4], (unsigned) vec[5], met[5]);
}

// ---- the wavefront ---------------------------------------------------------------------------
// One scan task = one window column x up to seven window rows: 14 reference rows of 8 bytes,
// three aligned words each.
struct WinRows { unsigned r0[14], r1[14], r2[14]; };

// request the words of a task's rows (row index clamped to `lastrow`, so every address is readable)
__device__ __forceinline__ void win_load (WinRows &W, const RowRef &rr, int rsw, int lastrow)
{
  // rows past `lastrow` repeat it: stride 0 from there on (the address is one multiply-add per row)
#pragma unroll
  for (int wr = 0; wr < 14; wr++) {
    const unsigned *wp = rr.w + (ptrdiff_t) min (wr, lastrow) * rsw;
    W.r0[wr] = __ldg (wp);
    W.r1[wr] = __ldg (wp + 1);
    W.r2[wr] = rr.three ? __ldg (wp + 2) : 0u;
  }
}

// seven running SADs of the source block against the rows of a task (FULL: all eight block rows)
template <bool FULL>
__device__ __forceinline__ void win_sads_body (const WinRows &W, unsigned sh, const uint2 (&srow)[8], const WaveBlock &B,
    unsigned (&sad)[7])
{
#pragma unroll
  for (int b = 0; b < 7; b++) sad[b] = 0;
#pragma unroll
  for (int wr = 0; wr < 14; wr++) {
    uint2 w = make_uint2 (__funnelshift_r (W.r0[wr], W.r1[wr], sh), __funnelshift_r (W.r1[wr], W.r2[wr], sh));
    if (!FULL) { w.x &= B.mlo; w.y &= B.mhi; }
#pragma unroll
    for (int y = 0; y < 8; y++) {
      const int b = wr - y;
      if (b >= 0 && b < 7) {
        if (FULL || y < B.hl) sad[b] += __vsadu4 (srow[y].x, w.x) + __vsadu4 (srow[y].y, w.y);
      }
    }
  }
}

__device__ __forceinline__ void win_sads (const WinRows &W, unsigned sh, const uint2 (&srow)[8], const WaveBlock &B,
    unsigned (&sad)[7])
{
  // (partial blocks exist only in the last block row / column: the branch is warp-uniform but
  // for a warp's steps through the last column)
  if (B.hl == 8 && B.bw0 == 8) win_sads_body<true> (W, sh, srow, B, sad);
  else win_sads_body<false> (W, sh, srow, B, sad);
}

struct WaveWin { int xmin, ymin, scan_w, scan_h, seed_a, seed_b; };

// seed clamp + scan window of a candidate vector (schrohierbm.c:349-364, schrometric.c:174-214)
__device__ __forceinline__ WaveWin wave_window (const HbmArgs &A, const WaveBlock &B, int vec)
{
  const int s = A.shift, e = A.ext, R = A.h_range;
  int dx = (vec >> 16) >> s, dy = ((int) (short) vec) >> s;
  dx = max (-B.bw0 - B.x0, min (A.width - B.x0, dx));
  dy = max (-B.hl - B.y0, min (A.height - B.y0, dy));
  WaveWin w;
  w.xmin = max (max (-B.bw0, B.x0 + dx - R), -e);
  w.ymin = max (max (-B.hl, B.y0 + dy - R), -e);
  const int xmax = min (min (A.width, B.x0 + dx + R), A.width - B.bw0 + e);
  const int ymax = min (min (A.height, B.y0 + dy + R), A.height - B.hl + e);
  w.scan_w = xmax - w.xmin + 1;
  w.scan_h = ymax - w.ymin + 1;
  w.seed_a = dx + B.x0 - w.xmin;
  w.seed_b = dy + B.y0 - w.ymin;
  return w;
}

// G lanes per block row, NT scan tasks per lane at most.  The body of a step is straight-line
// code (switched-off loads read a harmless address) apart from the warp-uniform branches around
// the neighbour SADs and the window reload, and the top row's wait for the warp above.
template <int G, int NT>
__global__ void __launch_bounds__ (32)
hbm_wave_kernel (const HbmArgs A, const uint2 *__restrict__ stat, int ngroups)
{
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x, sub = lane % G, q = lane / G, gbase = lane - sub;
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd (A.ticket, 1u);
  ticket = __shfl_sync (SB2_FULL, ticket, 0);
  const int rg = (int) (ticket / (unsigned) A.count), pic = (int) (ticket - (unsigned) rg * (unsigned) A.count);
  const int row = rg * RPW + q;
  const bool rowvalid = row < A.rows;
  const int s = A.shift, ri = A.ref_index, cols = A.cols;
  const int y0 = row * 8;
  const bool rowact = rowvalid && y0 < A.height;
  const bool has_up = row > 0;
  const WavePlanes P = wave_planes (A, pic);
  const uint8_t *safe = P.rp[0];
  const int rsw = P.rs[0] >> 2;
  MotionVector *mfrow = A.field + (size_t) pic * A.field_pitch + (size_t) (row << s) * A.nbx;
  unsigned long long *words_me = A.words + ((size_t) pic * ngroups + rg) * cols;
  const unsigned long long *words_up = rg > 0 ? words_me - cols : nullptr;
  const uint2 *statrow = stat + (((size_t) pic * A.rows + (rowvalid ? row : 0)) * cols) * 8 + (sub < 6 ? sub : 0);
  const bool statlane = rowvalid && sub < 6;

  int cur = 0, prev = 0;            // this row's results for the previous two columns
  int upw = 0;                      // top row: the vector polled for the previous column
  unsigned long long nextw = 0;
  uint2 snext = make_uint2 (0u, STAT_INVALID);
  if (words_up) nextw = ld_word (words_up);
  if (q == 0 && statlane && cols > 0) snext = __ldg (statrow);

#ifdef SB2_HBM_TRACE
  const bool wtrace = s == g_wave_trace_sel[0] && rg == g_wave_trace_sel[1] && pic == 0;
#endif
  for (int t = 0; t < cols + RPW - 1; t++) {
    WTRACE (0);
    const int c = t - q;
    const bool inrow = rowvalid && c >= 0 && c < cols;
    const WaveBlock B = wave_block (A, c * 8, y0);
    const bool act = inrow && rowact && B.x0 < A.width;
    const uint2 sv = snext;
    {
      // (written straight into the register next step reads: selecting on the loaded value here
      // would stall the step on this load)
      const int cn = c + 1;
      snext = make_uint2 (0u, STAT_INVALID);
      if (statlane && cn >= 0 && cn < cols) snext = __ldg (statrow + (size_t) cn * 8);
    }

    // ---- neighbours of this level: left is `cur`, up / up-left come from the lanes above,
    // the warp's top row takes them from the words the warp above publishes
    int up = __shfl_up_sync (SB2_FULL, cur, G);
    int upl = __shfl_up_sync (SB2_FULL, prev, G);
    if (words_up && t < cols) {
      // every lane takes part in the wait (same address: one request), so the loop is warp-uniform
      // and the code after it is known to be converged; only the top row uses the word.  The
      // sleep grows: a waiting warp is behind a slower producer, its wake-up delay is not on the
      // critical path, its polling would take issue slots and L2 requests from the producer
      unsigned long long w = nextw;
      unsigned ns = SB2_WAVE_POLL_NS;
      unsigned polls = 0;
      while (!(w >> 63)) {
        __nanosleep (ns);
        ns = min (ns * 2, 1024u);
        w = ld_word (words_up + t);
        // a producer holds a smaller ticket and is running: seconds of waiting mean a broken launch, which
        // must end as an error, not as a hung GPU
        if (++polls > (1u << 22)) __trap ();
      }
      nextw = t + 1 < cols ? ld_word (words_up + t + 1) : 0ull;
      if (q == 0) {
        upl = upw;
        up = (int) (unsigned) w;
      }
      upw = (int) (unsigned) w;
    }
    __syncwarp ();
    WTRACE (1);

    // ---- source rows: requested now, used after the ranking
    uint2 srow[8];
    {
      const uint8_t *sb = act ? P.sp[0] + (ptrdiff_t) y0 * P.ss[0] + B.x0 : safe;
      const int sstr = act ? P.ss[0] : 0, slast = B.hl - 1;
#pragma unroll
      for (int y = 0; y < 8; y++)
        srow[y] = __ldg (reinterpret_cast<const uint2 *> (sb + (ptrdiff_t) min (y, slast) * sstr));
    }

    // ---- candidates 0..8 (schrohierbm.c:255-294): lanes 0-5 static, 6 left, 7 up, 8 up-left
    int vec = 0;
    bool valid = false;
    unsigned met = STAT_INVALID;
    if (sub < 6) { vec = (int) sv.x; met = sv.y; valid = act && sv.y != STAT_INVALID; }
    else if (sub == 6) { vec = cur; valid = act && c > 0; }
    else if (sub == 7) { vec = up; valid = act && has_up; }
    const bool v8 = act && c > 0 && has_up;

    // de-duplication keeps the LAST occurrence (:298-321) and ranking the first strict minimum
    // (:323-346): the winner is the minimum over candidates of (SAD, index of the last
    // candidate with the same vector).  p = the candidates 0..7 of my group that hold my vector
    // (eight independent shuffles; a 64-bit MATCH.ANY was measured at ~500 cycles on the chain)
    unsigned p = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (__shfl_sync (SB2_FULL, vec, gbase + k) == vec) p |= 1u << k;
    p &= (__ballot_sync (SB2_FULL, valid) >> gbase) & 0xffu;
    if (!valid) p = 0;
    const bool eq8 = v8 && valid && vec == upl;
    const int lastidx = eq8 ? 8 : 31 - __clz (p | 1u);
    const unsigned b8 = (__ballot_sync (SB2_FULL, eq8) >> gbase) & 0xffu;
    const unsigned ts = p & 0x3fu;                    // static twins of my vector
    WTRACE (2);
    // SADs of left / up / up-left: a static twin's SAD is re-used, the rest are computed together
    const bool need6 = sub == 6 && valid && !ts;
    const bool need7 = sub == 7 && valid && !ts && !(p & 0x40u);
    const bool nd[3] = { __shfl_sync (SB2_FULL, (int) need6, gbase + 6) != 0,
                         __shfl_sync (SB2_FULL, (int) need7, gbase + 7) != 0, v8 && !b8 };
    const int nv[3] = { cur, up, upl };
    unsigned nm[3] = { STAT_INVALID, STAT_INVALID, STAT_INVALID };
    if (__any_sync (SB2_FULL, nd[1] || nd[2])) wave_rank_sads<3> (A, P, B, nv, nd, sub, nm);
    else if (__any_sync (SB2_FULL, nd[0])) {
      // the usual case inside coherent motion: up and up-left repeat a known vector
      const int v1[1] = { cur };
      const bool w1[1] = { nd[0] };
      unsigned m1[1];
      wave_rank_sads<1> (A, P, B, v1, w1, sub, m1);
      nm[0] = m1[0];
    }
    {
      const unsigned m2 = __shfl_sync (SB2_FULL, met, gbase + (ts ? __ffs (ts) - 1 : sub));
      if ((sub == 6 || sub == 7) && valid) met = ts ? m2 : (sub == 6 || (p & 0x40u)) ? nm[0] : nm[1];
    }
    unsigned met8 = nm[2];
    {
      const unsigned m = __shfl_sync (SB2_FULL, met, gbase + (b8 ? __ffs (b8) - 1 : 0));
      if (b8) met8 = m;
    }
    WTRACE (3);
    unsigned rkey = valid ? ((met << 4) | (unsigned) lastidx) : 0xffffffffu;
    rkey = min (rkey, __shfl_xor_sync (SB2_FULL, rkey, 1));
    rkey = min (rkey, __shfl_xor_sync (SB2_FULL, rkey, 2));
    rkey = min (rkey, __shfl_xor_sync (SB2_FULL, rkey, 4));
    if (v8) rkey = min (rkey, (met8 << 4) | 8u);      // (met8 is only good in the group's first eight lanes)
    rkey = __shfl_sync (SB2_FULL, rkey, gbase);
    const int widx = (int) (rkey & 15u);
    int wvec = __shfl_sync (SB2_FULL, vec, gbase + min (widx, 7));
    if (widx == 8) wvec = upl;

    const WaveWin ww = wave_window (A, B, wvec);
    WTRACE (4);

    // ---- scan (schrometric.c:31-71) + arg-min (:121-171): key = (SAD, not-seed, a, b)
    unsigned best = 0xffffffffu;
    {
      const int nch = NT == 1 ? 1 : max (1, (ww.scan_h + 6) / 7);
      const int ntasks = act ? ww.scan_w * nch : 0;
      const uint8_t *rwin = P.rp[0] + (ptrdiff_t) ww.ymin * P.rs[0] + ww.xmin;
#pragma unroll 1
      for (int it = 0; it < NT; it++) {
        const int task = sub + it * G;
        const bool tv = task < ntasks;
        if (NT > 1 && !__any_sync (SB2_FULL, tv)) break;
        const int a = NT == 1 ? task : task / nch;
        const int b0 = NT == 1 ? 0 : (task - a * nch) * 7;
        const int nb = min (7, ww.scan_h - b0);
        const RowRef rr = row_ref (tv ? rwin + (ptrdiff_t) b0 * P.rs[0] + a : safe);
        WinRows W;
        win_load (W, rr, rsw, tv ? nb + B.hl - 2 : 0);
        __syncwarp ();
        unsigned sad[7];
        if (it == 0 && B.bw0 < 8) {
#pragma unroll
          for (int y = 0; y < 8; y++) { srow[y].x &= B.mlo; srow[y].y &= B.mhi; }
        }
        win_sads (W, rr.sh, srow, B, sad);
        // key = (SAD, not-seed, a, b): a is the lane's, so the minimum over b of (SAD, b) decides
        // unless the seed sits in this task with the same SAD at a larger b
        unsigned kb = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 7; b++)
          if (b < nb) kb = min (kb, (sad[b] << 3) | (unsigned) b);
        unsigned seed_sad = 0xffffffffu;
        const int sbi = ww.seed_b - b0;
#pragma unroll
        for (int b = 0; b < 7; b++)
          if (b == sbi) seed_sad = sad[b];
        const bool seed_here = a == ww.seed_a && sbi >= 0 && sbi < nb && seed_sad == (kb >> 3);
        const unsigned bsel = seed_here ? (unsigned) sbi : (kb & 7u);
        const unsigned key = ((kb >> 3) << 13) | ((seed_here ? 0u : 1u) << 12) | ((unsigned) a << 6) | (unsigned) (b0 + (int) bsel);
        if (tv) best = min (best, key);
      }
    }
    WTRACE (5);
#pragma unroll
    for (int o = 1; o < G; o <<= 1) best = min (best, __shfl_xor_sync (SB2_FULL, best, o));
    WTRACE (6);

    const int ba = (int) ((best >> 6) & 63u), bb = (int) (best & 63u);
    const int rdx = (ww.xmin + ba - B.x0) << s, rdy = (ww.ymin + bb - y0) << s;
    const int res = act ? (int) (((unsigned) rdx << 16) | ((unsigned) rdy & 0xffffu)) : 0;
    if (inrow) { prev = cur; cur = res; }
    if (sub == 0) {
      if (act) {
        MotionVector *o = mfrow + (c << s);
        o->metric = best >> 13;
        o->v[ri] = (int16_t) rdx;
        o->v[2 + ri] = (int16_t) rdy;
      }
      if (inrow && q == RPW - 1) st_word (words_me + c, (1ull << 63) | (unsigned) res);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

static int wave_g (int h_range) { return h_range <= 3 ? 8 : h_range <= 5 ? 16 : 32; }

bool hbm_wave_supported (const HbmArgs &A, int h_range)
{
  if (A.bw != 8 || A.bh != 8 || A.hs != 1 || A.vs != 1 || A.use_chroma || A.ext < 8 || h_range > 20) return false;
  size_t al8 = (size_t) A.src.base | A.src.pic_pitch | A.src.off[0] | (size_t) A.src.stride[0];
  size_t al4 = (size_t) A.ref.base | A.ref.pic_pitch;
  for (int k = 0; k < 3; k++)
    al4 |= A.src.off[k] | (size_t) A.src.stride[k] | A.ref.off[k] | (size_t) A.ref.stride[k];
  return (al8 & 7) == 0 && (al4 & 3) == 0;
}

size_t hbm_wave_workspace_bytes (int rows, int cols, int count)
{
  return (size_t) rows * cols * count * sizeof (uint2) * 8;
}

int hbm_wave_launch (const HbmArgs &A, int h_range, void *stat_ws, size_t stat_bytes, cudaStream_t st, double bytes)
{
  const size_t need = hbm_wave_workspace_bytes (A.rows, A.cols, A.count);
  if (!stat_ws || stat_bytes < need)
    return set_error (SB2_ERR_WORKSPACE, "sb2_hbm_scan_hint: candidate workspace %zu < %zu", stat_bytes, need);
  uint2 *stat = static_cast<uint2 *> (stat_ws);
  const long long groups = (long long) A.count * A.rows * A.cols;
  {
    LaunchScope scope ("hbm_static", (double) groups * 48, st);
    hbm_static_kernel<<<(unsigned) ((groups + 127) / 128), 128, 0, st>>> (A, stat);
  }
  const int G = wave_g (h_range), rpw = 32 / G;
  const int ngroups = ceil_div (A.rows, rpw);
  const unsigned ctas = (unsigned) ngroups * (unsigned) A.count;
  char tag[48];
  snprintf (tag, sizeof (tag), "hbm_level_s%d_r%d", A.shift, h_range);
  LaunchScope scope (tag, bytes, st);
  if (h_range <= 3) hbm_wave_kernel<8, 1><<<ctas, 32, 0, st>>> (A, stat, ngroups);
  else if (h_range <= 5) hbm_wave_kernel<16, 2><<<ctas, 32, 0, st>>> (A, stat, ngroups);
  else if (h_range <= 10) hbm_wave_kernel<32, 2><<<ctas, 32, 0, st>>> (A, stat, ngroups);
  else hbm_wave_kernel<32, 8><<<ctas, 32, 0, st>>> (A, stat, ngroups);
  return SB2_OK;
}

}  // namespace sb2

#ifdef SB2_HBM_TRACE
extern "C" int sb2_hbm_wave_trace_select (int level, int row_group)
{
  const int v[2] = { level, row_group };
  return (int) cudaMemcpyToSymbol (sb2::g_wave_trace_sel, v, sizeof (v));
}
extern "C" int sb2_hbm_wave_trace_read (long long *host, int n)
{
  return (int) cudaMemcpyFromSymbol (host, sb2::g_wave_trace, sizeof (long long) * n);
}
#endif

