// hbm.cu -- SAD primitives and hierarchical block matching for sm_100a.
//
// Bit-exact replacement for one level of schro_hierarchical_bm_scan_hint
// (schroedinger/schrohierbm.c:174-383) on top of schro_metric_scan_setup / _do_scan /
// _get_min and schro_metric_fast_block (schroedinger/schrometric.c:31-214, 332-414).
//
// Parallel structure (DESIGN.md "block matching"): a block needs the already-computed
// vectors of its left, upper and upper-left neighbours OF THE SAME LEVEL, so blocks form a
// wavefront.  One CTA (2-16 warps) owns one block row of one (picture, reference) pair and
// walks it left to right; a finished block publishes its vector as one relaxed 64-bit word
// and the row below polls the two words it needs.  Rows take their index from an atomic ticket, so
// a row only ever waits on a CTA that has already started.  Many pairs run side by side in one
// launch to fill the machine.  Inside a block three lanes per candidate compute the ranking
// SADs, the threads split the (2r+1)^2 scan positions; byte SADs use __vsadu4, reductions
// use REDUX / warp shuffles.

#include "hbm_common.cuh"
#include <cstdlib>

namespace sb2 {

__global__ void __launch_bounds__ (128)
hbm_init_field_kernel (MotionVector *field, size_t n, uint32_t flags0, unsigned long long *words, size_t nwords,
    unsigned *ticket)
{
  if (ticket && blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0;
  // the published-result words of the level start out invalid (same launch: one API call less per level)
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nwords; i += (size_t) gridDim.x * blockDim.x)
    words[i] = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    MotionVector m;
    m.flags = flags0;
    m.metric = 0;
    m.chroma_metric = 0;
    m.v[0] = m.v[1] = m.v[2] = m.v[3] = 0;
    field[i] = m;
  }
}

#ifdef SB2_HBM_TRACE
__device__ long long g_hbm_trace[512 * 8];
#define TRACE(k) do { if (trace_on && threadIdx.x == 0 && bi < 512) g_hbm_trace[bi * 8 + (k)] = clock64 (); } while (0)
#else
#define TRACE(k) do { } while (0)
#endif

#ifndef POLL_NS
#define POLL_NS 100
#endif

// candidates that do not depend on this level's neighbours, prepared one block ahead
struct StaticCands {
  int dx[6], dy[6];
  unsigned metric[6];
  unsigned valid;
  int full;
};

struct Win { int xmin, ymin, scan_w, scan_h, seed_a, seed_b; };

struct BlockShared {
  StaticCands stc[2];
  Win win;
  int last_dx, last_dy;             // this row's previous block (the "left" candidate)
  unsigned ticket;
  unsigned long long key[16];
  unsigned luma[16], chroma[16];
};

// One CTA (NW warps) owns one block row of one (picture, reference) pair.
//
// Per block the dependent chain is: neighbour vectors (poll) -> rank candidates -> scan around
// the winner -> publish.  Everything that does not depend on the neighbours runs one block
// ahead on warp 1: the static candidates (zero + five parents) with their ranking SADs.
// (Letting warp 1 also scan speculatively around the best static candidate was measured slower
// -- 4.4 vs 3.6 ms at level 0: warp 1 becomes the longer loop -- and removed)
// 64 registers per thread (85 for the two-warp variant, measured best at 32 pictures per launch):
// a row's CTA mostly waits, so its footprint in the register file decides how many rows /
// pictures / other kernels an SM can host, while too tight a cap makes the compiler recompute
// addresses it could have kept.
#ifndef SB2_HBM_THREADS_PER_SM
#define SB2_HBM_THREADS_PER_SM 1024
#endif
template <int NW>
__global__ void __launch_bounds__ (32 * NW, NW == 2 ? 12 : SB2_HBM_THREADS_PER_SM / (32 * NW))
hbm_level_kernel (const HbmArgs A)
{
  static_assert (NW >= 2, "one warp prepares the static candidates of the next block");
  __shared__ BlockShared sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (row, picture) from an atomic ticket: the CTA of the row above holds a smaller ticket, so it
  // has started -- forward progress does not depend on the order CTAs are dispatched in
  if (threadIdx.x == 0) sh.ticket = atomicAdd (A.ticket, 1u);
  __syncthreads ();
  const int row = (int) (sh.ticket / (unsigned) A.count), pic = (int) (sh.ticket % (unsigned) A.count);
  const int skip = 1 << A.shift, s = A.shift;
  const int j = row * skip;
  const int ri = A.ref_index;

  const uint8_t *sp[3], *rp[3];
  int ss[3], rs[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    sp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (A.src, pic, k));
    rp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref, pic, k));
    ss[k] = A.src.stride[k];
    rs[k] = A.ref.stride[k];
  }
  MotionVector *mf = A.field + (size_t) pic * A.field_pitch;
  const MotionVector *pf = A.parent ? A.parent + (size_t) pic * A.field_pitch : nullptr;
  unsigned long long *words_me = A.words + ((size_t) pic * A.rows + row) * A.cols;
  const unsigned long long *words_up = row > 0 ? words_me - A.cols : nullptr;
  if (threadIdx.x == 0) { sh.last_dx = 0; sh.last_dy = 0; }
  const int hint_mask = ~((1 << (s + 1)) - 1);
  const int y0 = (j * A.bh) >> s;
  const int bh0 = min (A.height - y0, A.bh);
  const int e = A.ext;
  // every alignment assumption of the byte-SIMD paths, checked once
  const bool simd_ok = A.bw == 8 && A.bh == 8 && A.hs == 1 && A.vs == 1 &&
      ((((size_t) sp[0] | (size_t) ss[0]) & 7) == 0) && ((((size_t) sp[1] | (size_t) sp[2] | (size_t) ss[1] | (size_t) ss[2]) & 3) == 0) &&
      ((((size_t) rp[0] | (size_t) rp[1] | (size_t) rp[2] | (size_t) rs[0] | (size_t) rs[1] | (size_t) rs[2]) & 3) == 0);

  // ranking SAD of one candidate vector for the block at (x0, y0): three lanes per candidate
  // (luma rows 0-3, luma rows 4-7, both chroma blocks), result in the first of the three
  auto cand_sad = [&] (int x0, int bw0, int part, int kdx, int kdy, bool want) -> unsigned {
    int dx = kdx >> s, dy = kdy >> s;
    dx = clampi (dx + x0, -bw0, A.width) - x0;
    dy = clampi (dy + y0, -bh0, A.height) - y0;
    const bool ok = !(x0 < -e || y0 < -e || x0 + 8 > A.width + e || y0 + 8 > A.height + e) &&
        !(x0 + dx < -e || y0 + dy < -e || x0 + dx + 8 > A.width + e || y0 + dy + 8 > A.height + e);
    unsigned part_sad = 0;
    if (ok && want) {
      if (part < 2) {
        const uint2 *a = reinterpret_cast<const uint2 *> (sp[0] + (ptrdiff_t) (y0 + 4 * part) * ss[0] + x0);
        const RowRef b = row_ref (rp[0] + (ptrdiff_t) (y0 + dy + 4 * part) * rs[0] + x0 + dx);
        const int asw = ss[0] >> 3, rsw = rs[0] >> 2;
#pragma unroll
        for (int y = 0; y < 4; y++) {
          const uint2 av = __ldg (a + y * asw);
          const uint2 bv = row_load8 (b, y * rsw);
          part_sad += __vsadu4 (av.x, bv.x) + __vsadu4 (av.y, bv.y);
        }
      } else {
        const int sx = x0 >> 1, sy = y0 >> 1, rx = (x0 + dx) >> 1, ry = (y0 + dy) >> 1;
#pragma unroll
        for (int c = 1; c < 3; c++) {
          const unsigned *a = reinterpret_cast<const unsigned *> (sp[c] + (ptrdiff_t) sy * ss[c] + sx);
          const RowRef b = row_ref (rp[c] + (ptrdiff_t) ry * rs[c] + rx);
          const int asw = ss[c] >> 2, rsw = rs[c] >> 2;
#pragma unroll
          for (int y = 0; y < 4; y++)
            part_sad += __vsadu4 (__ldg (a + y * asw), row_load4 (b, y * rsw));
        }
      }
    }
    unsigned m = part_sad + __shfl_down_sync (0xffffffffu, part_sad, 1);
    m += __shfl_down_sync (0xffffffffu, part_sad, 2);
    return ok ? m : (unsigned) INT_MAX;
  };

  // seed clamp + scan window (schrohierbm.c:349-364, schrometric.c:174-214); (dx, dy) in pixels
  auto clamp_seed = [&] (int x0, int bw0, int &dx, int &dy) {
    dx = max (-bw0 - x0, min (A.width - x0, dx));
    dy = max (-bh0 - y0, min (A.height - y0, dy));
  };
  auto make_win = [&] (int x0, int bw0, int dx, int dy) -> Win {
    Win w;
    w.xmin = max (max (-bw0, x0 + dx - A.h_range), -e);
    w.ymin = max (max (-bh0, y0 + dy - A.h_range), -e);
    const int xmax = min (min (A.width, x0 + dx + A.h_range), A.width - bw0 + e);
    const int ymax = min (min (A.height, y0 + dy + A.h_range), A.height - bh0 + e);
    w.scan_w = xmax - w.xmin + 1;
    w.scan_h = ymax - w.ymin + 1;
    w.seed_a = dx + x0 - w.xmin;
    w.seed_b = dy + y0 - w.ymin;
    return w;
  };

  // full search over positions first, first+stride, ...: the calling warp's best
  // key = (metric, not-seed, a, b) -- seed wins ties, else first strict minimum in the
  // reference's a-outer / b-inner order (schrometric.c:121-171) -- with its luma / chroma SADs
  auto scan_part = [&] (int x0, int bw0, const Win &wn, int first, int stride,
      unsigned long long &wkey, unsigned &wl, unsigned &wc) {
    unsigned long long best_key = ~0ull;
    unsigned best_l = 0, best_c = 0;
    const uint8_t *sblk = sp[0] + (ptrdiff_t) y0 * ss[0] + x0;
    const int npos = wn.scan_w * wn.scan_h;
    const bool fast8 = simd_ok && bw0 == 8 && bh0 == 8;
    uint2 srow[8];
    if (fast8) {
#pragma unroll
      for (int y = 0; y < 8; y++) srow[y] = __ldg (reinterpret_cast<const uint2 *> (sblk + (ptrdiff_t) y * ss[0]));
    }
    // p / scan_w without an integer division: p < 2048 and scan_w <= 41, so the float quotient
    // of (p + 0.5) is at least 0.5 / 41 away from an integer -- far beyond the rounding error
    const float inv_w = __frcp_rn ((float) wn.scan_w);
    const uint8_t *rwin = rp[0] + (ptrdiff_t) wn.ymin * rs[0] + wn.xmin;
    const int rsw = rs[0] >> 2;
    for (int p = first; p < npos; p += stride) {
      const int b = (int) (((float) p + 0.5f) * inv_w), a = p - b * wn.scan_w;
      const uint8_t *rblk = rwin + b * rs[0] + a;
      unsigned l = 0;
      if (fast8) {
        const RowRef rr = row_ref (rblk);
#pragma unroll
        for (int y = 0; y < 8; y++) {
          const uint2 bv = row_load8 (rr, y * rsw);
          l += __vsadu4 (srow[y].x, bv.x) + __vsadu4 (srow[y].y, bv.y);
        }
      } else {
        l = block_sad (sblk, ss[0], rblk, rs[0], bw0, bh0);
      }
      unsigned c = 0;
      if (A.use_chroma) {
        // chroma_metrics[a*scan_h+b] = sum_k SAD_k at (ref_x/2 + a/2, ref_y/2 + b/2)
        // (schrometric.c:73-115; C division truncates toward zero)
        const int cx = x0 / 2, cy = y0 / 2, crx = wn.xmin / 2 + (a >> 1), cry = wn.ymin / 2 + (b >> 1);
        for (int k = 1; k < 3; k++)
          c += block_sad (sp[k] + (ptrdiff_t) cy * ss[k] + cx, ss[k], rp[k] + (ptrdiff_t) cry * rs[k] + crx, rs[k],
              bw0 / 2, bh0 / 2);
      }
      const unsigned tot = l + c;
      const unsigned notseed = (a == wn.seed_a && b == wn.seed_b) ? 0u : 1u;
      const unsigned long long key = ((unsigned long long) tot << 32) | (notseed << 24) | ((unsigned) a << 12) | (unsigned) b;
      if (key < best_key) { best_key = key; best_l = l; best_c = c; }
    }
    // min over (tot, low bits): two 32-bit REDUX instead of a 64-bit shuffle tree
    const unsigned khi = (unsigned) (best_key >> 32), klo = (unsigned) best_key;
    const unsigned mhi = __reduce_min_sync (0xffffffffu, khi);
    const unsigned mlo = __reduce_min_sync (0xffffffffu, khi == mhi ? klo : 0xffffffffu);
    wkey = ((unsigned long long) mhi << 32) | mlo;
    const unsigned owner = __ballot_sync (0xffffffffu, best_key == wkey);
    const int ol = __ffs (owner) - 1;
    wl = __shfl_sync (0xffffffffu, best_l, ol);
    wc = __shfl_sync (0xffffffffu, best_c, ol);
  };

  // one thread: publish the vector (the row below is waiting on it), then fill the output field
  auto publish = [&] (int bi, int i, int x0, const Win &wn, unsigned long long k, unsigned bl, unsigned bc) {
    const int a = (int) ((k >> 12) & 0xfff), b = (int) (k & 0xfff);
    const int rdx = (wn.xmin + a - x0) << s, rdy = (wn.ymin + b - y0) << s;
    st_word (words_me + bi, pack_word ((int16_t) rdx, (int16_t) rdy));
    sh.last_dx = (int16_t) rdx; sh.last_dy = (int16_t) rdy;
    MotionVector *o = mf + (size_t) j * A.nbx + i;
    o->metric = bl;
    o->chroma_metric = bc;
    o->v[ri] = (int16_t) rdx;
    o->v[2 + ri] = (int16_t) rdy;
    o->flags = A.flags0;
  };

  // static candidates of block `nb` (0: zero, 1..5: parents (0,0) (-1,0) (1,0) (0,-1) (0,1),
  // schrohierbm.c:259-277) and their ranking SADs into sh.stc[nb & 1].  Executed by one whole warp.
  auto do_static = [&] (int nb) {
    const int i = nb * skip;
    const int x0 = (i * A.bw) >> s;
    if (!(x0 < A.width && y0 < A.height)) return;
    const int bw0 = min (A.width - x0, A.bw);
    int cdx = 0, cdy = 0;
    bool valid = false;
    if (lane == 0) valid = true;
    else if (lane <= 5 && pf) {
      const int ox = (lane == 2) ? -1 : (lane == 3) ? 1 : 0;
      const int oy = (lane == 4) ? -1 : (lane == 5) ? 1 : 0;
      const int ll = (i & hint_mask) + ox * skip * 2, kk = (j & hint_mask) + oy * skip * 2;
      if (ll >= 0 && ll < A.nbx && kk >= 0 && kk < A.nby) {
        const MotionVector *m = pf + (size_t) kk * A.nbx + ll;
        cdx = m->v[ri]; cdy = m->v[2 + ri]; valid = true;
      }
    }
    const bool full = simd_ok && x0 + 8 <= A.width && y0 + 8 <= A.height &&
        (x0 >> 1) + 4 <= A.cw && (y0 >> 1) + 4 <= A.ch;
    const int ck = lane / 3, part = lane - 3 * ck;
    StaticCands &sc = sh.stc[nb & 1];
    unsigned m = (unsigned) INT_MAX;
    if (full) {
      const int kdx = __shfl_sync (0xffffffffu, cdx, min (ck, 5)), kdy = __shfl_sync (0xffffffffu, cdy, min (ck, 5));
      const bool kval = __shfl_sync (0xffffffffu, (int) valid, min (ck, 5)) != 0;
      m = cand_sad (x0, bw0, part, kdx, kdy, ck < 6 && kval);
      if (ck < 6 && part == 0) sc.metric[ck] = m;
    }
    const unsigned vm = __ballot_sync (0xffffffffu, valid && lane < 6);
    if (lane < 6) { sc.dx[lane] = cdx; sc.dy[lane] = cdy; }
    if (lane == 0) { sc.valid = vm; sc.full = full; }
  };

  __syncthreads ();
  if (warp == 1) do_static (0);
  __syncthreads ();

#ifdef SB2_HBM_TRACE
  const bool trace_on = (A.shift == 0 && row == 100 && pic == 0);
#endif
  for (int bi = 0; bi < A.cols; bi++) {
    TRACE (0);
    const int i = bi * skip;
    const int x0 = (i * A.bw) >> s;
    const bool active = x0 < A.width && y0 < A.height;
    const int bw0 = min (A.width - x0, A.bw);

    // ---- warp 1 runs one block ahead
    if (warp == 1 && bi + 1 < A.cols) do_static (bi + 1);

    if (warp == 0 && active) {
      const StaticCands &sc = sh.stc[bi & 1];
      int cdx = 0, cdy = 0;
      bool valid = false;
      if (lane < 6) { cdx = sc.dx[lane]; cdy = sc.dy[lane]; valid = (sc.valid >> lane) & 1; }
      const bool full = sc.full != 0;
      const int ck = lane / 3, part = lane - 3 * ck;
      unsigned metric = (part == 0 && ck < 6) ? sc.metric[ck] : (unsigned) INT_MAX;   // candidate ck's SAD in lane 3*ck

      TRACE (1);
      // ---- 6: left 7: up 8: up-left of THIS level (schrohierbm.c:279-294).  left comes from
      // this CTA's previous block; up / up-left are polled straight out of the row above's
      // published words, concurrently (one L2 round trip when both are ready), backing off
      // between polls so that waiting rows do not crowd the L2 request path
      if (lane == 6 && i > 0) {
        cdx = sh.last_dx; cdy = sh.last_dy; valid = true;
      } else if ((lane == 7 || (lane == 8 && i > 0)) && words_up) {
        const unsigned long long *wp = words_up + bi - (lane == 8 ? 1 : 0);
        unsigned long long wv = ld_word (wp);
        unsigned polls = 0;
        while (!(wv >> 63)) {
          __nanosleep (POLL_NS);
          wv = ld_word (wp);
          if (++polls > (1u << 25)) __trap ();     // seconds of waiting on a row that has started: an error, not a hang
        }
        cdx = (int) (short) (wv >> 16); cdy = (int) (short) wv; valid = true;
      }
      __syncwarp ();
      // de-duplicate keeping the LAST occurrence (schrohierbm.c:298-321): a candidate is
      // dropped when a later lane holds the same vector
      const bool isc = valid && lane < 9;
      const unsigned long long mkey = isc
          ? ((1ull << 32) | ((unsigned long long) (cdx & 0xffff) << 16) | (unsigned long long) (cdy & 0xffff))
          : ((unsigned long long) (2 + lane) << 32);
      const unsigned same = __match_any_sync (0xffffffffu, mkey);
      const bool dup = isc && (same >> (lane + 1)) != 0;
      const unsigned cmask = __ballot_sync (0xffffffffu, isc && !dup);

      TRACE (2);
      // ---- rank candidates with the 3-component SAD (schrometric.c:332-375) --------
      int best_k;
      if (full) {
        const int kdx = __shfl_sync (0xffffffffu, cdx, min (ck, 8)), kdy = __shfl_sync (0xffffffffu, cdy, min (ck, 8));
        // A neighbour whose vector equals one of the static candidates (the usual case in
        // coherent motion) has that candidate's SAD: reuse it instead of loading again.
        // `same` (from the de-duplication match) lists the lanes 0..8 holding my vector.
        const unsigned twins = __shfl_sync (0xffffffffu, same, min (ck, 8)) & 0x3fu;   // static lanes 0..5
        const bool reuse = ck >= 6 && ck < 9 && twins != 0;
        const int src_lane = 3 * (__ffs (twins | 0x80000000u) - 1);
        const unsigned reused = __shfl_sync (0xffffffffu, metric, reuse ? src_lane : 0);
        const bool need = ck >= 6 && ck < 9 && ((cmask >> min (ck, 8)) & 1) && !reuse;
        if (__any_sync (0xffffffffu, need)) {
          const unsigned m = cand_sad (x0, bw0, part, kdx, kdy, need);
          if (need) metric = m;
        }
        if (reuse) metric = reused;
        // first strict minimum in candidate order == min over (metric, k); INT_MAX never wins
        unsigned key = 0xffffffffu;
        if (ck < 9 && part == 0 && ((cmask >> ck) & 1) && metric < (unsigned) INT_MAX) key = (metric << 8) | (unsigned) ck;
        key = __reduce_min_sync (0xffffffffu, key);
        best_k = key == 0xffffffffu ? __ffs (cmask) - 1 : (int) (key & 0xff);
      } else {
        best_k = -1;
        unsigned best_metric = (unsigned) INT_MAX;
        for (int k = 0; k < 9; k++) {
          if (!((cmask >> k) & 1)) continue;
          int dx = __shfl_sync (0xffffffffu, cdx, k) >> s;
          int dy = __shfl_sync (0xffffffffu, cdy, k) >> s;
          dx = clampi (dx + x0, -bw0, A.width) - x0;
          dy = clampi (dy + y0, -bh0, A.height) - y0;
          unsigned m;
          const bool ok = !(x0 < -e || y0 < -e || x0 + A.bw > A.width + e || y0 + A.bh > A.height + e) &&
              !(x0 + dx < -e || y0 + dy < -e || x0 + dx + A.bw > A.width + e || y0 + dy + A.bh > A.height + e);
          if (!ok) {
            m = (unsigned) INT_MAX;
          } else {
            unsigned part_sum = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const int hs = c ? A.hs : 0, vs = c ? A.vs : 0;
              const int sx = x0 >> hs, sy = y0 >> vs, rx = (x0 + dx) >> hs, ry = (y0 + dy) >> vs;
              const int w = min (max (0, (c ? A.cw : A.width) - sx), A.bw >> hs);
              const int h = min (max (0, (c ? A.ch : A.height) - sy), A.bh >> vs);
              for (int p = lane; p < w * h; p += 32) {
                const int yy = p / w, xx = p - yy * w;
                part_sum += (unsigned) abs ((int) __ldg (sp[c] + (ptrdiff_t) (sy + yy) * ss[c] + sx + xx)
                    - (int) __ldg (rp[c] + (ptrdiff_t) (ry + yy) * rs[c] + rx + xx));
              }
            }
            m = warp_sum (part_sum);
          }
          if ((int) m < (int) best_metric) { best_metric = m; best_k = k; }
        }
        if (best_k < 0) best_k = __ffs (cmask) - 1;    // every candidate invalid: the reference asserts
      }

      TRACE (3);
      int dx = __shfl_sync (0xffffffffu, cdx, best_k) >> s;
      int dy = __shfl_sync (0xffffffffu, cdy, best_k) >> s;
      clamp_seed (x0, bw0, dx, dy);
      if (lane == 0) sh.win = make_win (x0, bw0, dx, dy);
    }
    __syncthreads ();
    TRACE (4);
    if (active) {
      const Win wn = sh.win;
      unsigned long long wkey;
      unsigned wl, wc;
      scan_part (x0, bw0, wn, threadIdx.x, 32 * NW, wkey, wl, wc);
      TRACE (5);
      if (lane == 0) { sh.key[warp] = wkey; sh.luma[warp] = wl; sh.chroma[warp] = wc; }
      __syncthreads ();
      if (threadIdx.x == 0) {
        unsigned long long k = sh.key[0];
        unsigned bl = sh.luma[0], bc = sh.chroma[0];
#pragma unroll
        for (int w = 1; w < NW; w++)
          if (sh.key[w] < k) { k = sh.key[w]; bl = sh.luma[w]; bc = sh.chroma[w]; }
        TRACE (6);
        publish (bi, i, x0, wn, k, bl, bc);
      }
    }
    if (!active && threadIdx.x == 0) {
      // a block outside the picture keeps the zero vector of schro_motion_field_set
      st_word (words_me + bi, pack_word (0, 0));
      sh.last_dx = 0; sh.last_dy = 0;
    }
    __syncthreads ();
  }
}


}  // namespace sb2

using namespace sb2;

__global__ void __launch_bounds__ (128)
sad_batch_kernel (const uint8_t *a, int as, const uint8_t *b, int bs, const int64_t *ao, const int64_t *bo,
    int n, int w, int h, uint32_t *out)
{
  // one warp per block pair, lanes split the pixels (orc_sad_nxm_u8, schroorc.orc:1760-1811)
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n) return;
  const uint8_t *pa = a + ao[g], *pb = b + bo[g];
  unsigned part = 0;
  for (int p = lane; p < w * h; p += 32) {
    const int y = p / w, x = p - y * w;
    part += (unsigned) abs ((int) __ldg (pa + (ptrdiff_t) y * as + x) - (int) __ldg (pb + (ptrdiff_t) y * bs + x));
  }
  part = warp_sum (part);
  if (lane == 0) out[g] = part;
}

__global__ void __launch_bounds__ (32)
sad_dc_kernel (const uint8_t *a, int as, int value, int w, int h, int *out)
{
  // schro_metric_get_dc (schroedinger/schrometric.c:253-270)
  unsigned part = 0;
  for (int p = threadIdx.x; p < w * h; p += 32) {
    const int y = p / w, x = p - y * w;
    part += (unsigned) abs (value - (int) a[(ptrdiff_t) y * as + x]);
  }
  part = warp_sum (part);
  if (threadIdx.x == 0) *out = (int) part;
}

__global__ void __launch_bounds__ (32)
sad_biref_kernel (const uint8_t *a, int as, const uint8_t *s1, int s1s, int w1, const uint8_t *s2, int s2s,
    int w2, int shift, int w, int h, int *out)
{
  // schro_metric_get_biref (schroedinger/schrometric.c:272-304)
  const int offset = 1 << (shift - 1);
  unsigned part = 0;
  for (int p = threadIdx.x; p < w * h; p += 32) {
    const int y = p / w, x = p - y * w;
    const int v = ((int) s1[(ptrdiff_t) y * s1s + x] * w1 + (int) s2[(ptrdiff_t) y * s2s + x] * w2 + offset) >> shift;
    part += (unsigned) abs ((int) a[(ptrdiff_t) y * as + x] - v);
  }
  part = warp_sum (part);
  if (threadIdx.x == 0) *out = (int) part;
}

// ---- the metric-scan entry points on their own (schrometric.c:31-214, 332-414) ---------------
// One CTA per scan: the threads split the scan positions; metrics[i * scan_h + j] as
// schro_metric_scan_do_scan lays them out.  Chroma (use_chroma) follows the reference's
// duplication scheme: the chroma SADs are computed on the sub-sampled grid and entry (i, j) of the
// luma grid reads (i >> h_shift, j >> v_shift) of it (:73-115).
struct ScanArgs {
  PlaneSet src, ref;
  const sb2_metric_scan_desc *desc;
  uint32_t *metrics, *chroma;           // n x 42*42 each (chroma may be null)
  int hs, vs, use_chroma;
};

__global__ void __launch_bounds__ (128)
metric_scan_kernel (const ScanArgs a)
{
  const sb2_metric_scan_desc d = a.desc[blockIdx.x];
  const uint8_t *sp[3], *rp[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    sp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (a.src, d.picture, k));
    rp[k] = reinterpret_cast<const uint8_t *> (plane_ptr (a.ref, d.picture, k));
  }
  uint32_t *m = a.metrics + (size_t) blockIdx.x * (42 * 42);
  uint32_t *cm = a.chroma ? a.chroma + (size_t) blockIdx.x * (42 * 42) : nullptr;
  const int npos = d.scan_width * d.scan_height;
  const int skip_h = 1 << a.hs, skip_v = 1 << a.vs;
  for (int p = threadIdx.x; p < npos; p += blockDim.x) {
    const int i = p / d.scan_height, j = p - i * d.scan_height;
    m[p] = block_sad (sp[0] + (ptrdiff_t) d.y * a.src.stride[0] + d.x, a.src.stride[0],
        rp[0] + (ptrdiff_t) (d.ref_y + j) * a.ref.stride[0] + d.ref_x + i, a.ref.stride[0], d.block_width, d.block_height);
    if (cm) {
      unsigned c = 0;
      if (a.use_chroma) {
        const int cx = d.x / skip_h, cy = d.y / skip_v, crx = d.ref_x / skip_h + (i >> a.hs), cry = d.ref_y / skip_v + (j >> a.vs);
        for (int k = 1; k < 3; k++)
          c += block_sad (sp[k] + (ptrdiff_t) cy * a.src.stride[k] + cx, a.src.stride[k],
              rp[k] + (ptrdiff_t) cry * a.ref.stride[k] + crx, a.ref.stride[k], d.block_width / skip_h, d.block_height / skip_v);
      }
      cm[p] = c;
    }
  }
}

// schro_metric_block_sad_slow (schrometric.c:332-375) for n independent (block, vector) pairs: one warp each
struct Sad3Args {
  PlaneSet src, ref;
  const sb2_metric_block_desc *desc;
  int *out;
  int n, width, height, cw, ch, ext, bw, bh, hs, vs;
};

__global__ void __launch_bounds__ (128)
metric_block_sad3_kernel (const Sad3Args a)
{
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= a.n) return;
  const sb2_metric_block_desc d = a.desc[g];
  const int e = a.ext;
  const bool ok = !(d.x < -e || d.y < -e || d.x + a.bw > a.width + e || d.y + a.bh > a.height + e) &&
      !(d.x + d.dx < -e || d.y + d.dy < -e || d.x + d.dx + a.bw > a.width + e || d.y + d.dy + a.bh > a.height + e);
  unsigned part = 0;
  if (ok) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int hs = c ? a.hs : 0, vs = c ? a.vs : 0;
      const int sx = d.x >> hs, sy = d.y >> vs, rx = (d.x + d.dx) >> hs, ry = (d.y + d.dy) >> vs;
      const int w = min (max (0, (c ? a.cw : a.width) - sx), a.bw >> hs);
      const int h = min (max (0, (c ? a.ch : a.height) - sy), a.bh >> vs);
      const uint8_t *s = reinterpret_cast<const uint8_t *> (plane_ptr (a.src, d.picture, c));
      const uint8_t *r = reinterpret_cast<const uint8_t *> (plane_ptr (a.ref, d.picture, c));
      for (int p = lane; p < w * h; p += 32) {
        const int yy = p / w, xx = p - yy * w;
        part += (unsigned) abs ((int) __ldg (s + (ptrdiff_t) (sy + yy) * a.src.stride[c] + sx + xx)
            - (int) __ldg (r + (ptrdiff_t) (ry + yy) * a.ref.stride[c] + rx + xx));
      }
    }
  }
  part = warp_sum (part);
  if (lane == 0) a.out[g] = ok ? (int) part : INT_MAX;
}

extern "C" int
sb2_metric_scan (const sb2_slab *src, const sb2_slab *ref, int chroma_h_shift, int chroma_v_shift, int use_chroma,
    const sb2_metric_scan_desc *descs, int n, uint32_t *metrics, uint32_t *chroma_metrics, void *stream)
{
  if (!src || !ref || !descs || !metrics || n < 0 || src->ncomp != 3 || ref->ncomp != 3)
    return set_error (SB2_ERR_ARG, "sb2_metric_scan: bad argument");
  if (use_chroma && !chroma_metrics) return set_error (SB2_ERR_ARG, "sb2_metric_scan: chroma scan needs the chroma array");
  if (n == 0) return SB2_OK;
  ScanArgs a;
  a.src = planeset_from_slab (src);
  a.ref = planeset_from_slab (ref);
  a.desc = descs;
  a.metrics = metrics;
  a.chroma = chroma_metrics;
  a.hs = chroma_h_shift;
  a.vs = chroma_v_shift;
  a.use_chroma = use_chroma;
  {
    LaunchScope scope ("metric_scan", 0.0, as_stream (stream));
    metric_scan_kernel<<<n, 128, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "metric_scan_kernel launch");
}

extern "C" int
sb2_metric_block_sad3 (const sb2_slab *src, const sb2_slab *ref, int extension, int block_width, int block_height,
    int chroma_h_shift, int chroma_v_shift, const sb2_metric_block_desc *descs, int n, int *metric, void *stream)
{
  if (!src || !ref || !descs || !metric || n < 0 || src->ncomp != 3 || ref->ncomp != 3 || block_width < 1 || block_height < 1)
    return set_error (SB2_ERR_ARG, "sb2_metric_block_sad3: bad argument");
  if (n == 0) return SB2_OK;
  Sad3Args a;
  a.src = planeset_from_slab (src);
  a.ref = planeset_from_slab (ref);
  a.desc = descs;
  a.out = metric;
  a.n = n;
  a.width = src->width[0];
  a.height = src->height[0];
  a.cw = src->width[1];
  a.ch = src->height[1];
  a.ext = extension;
  a.bw = block_width;
  a.bh = block_height;
  a.hs = chroma_h_shift;
  a.vs = chroma_v_shift;
  {
    LaunchScope scope ("metric_block_sad3", 0.0, as_stream (stream));
    metric_block_sad3_kernel<<<ceil_div (n, 4), 128, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "metric_block_sad3_kernel launch");
}

extern "C" int
sb2_sad_dc_u8 (const uint8_t *a, int a_stride, int value, int width, int height, int *sad, void *stream)
{
  if (!a || !sad || width < 1 || height < 1) return sb2::set_error (SB2_ERR_ARG, "sb2_sad_dc_u8: bad argument");
  {
    sb2::LaunchScope scope ("sad_dc", 1.0 * width * height, sb2::as_stream (stream));
    sad_dc_kernel<<<1, 32, 0, sb2::as_stream (stream)>>> (a, a_stride, value, width, height, sad);
  }
  return sb2::check_cuda (cudaGetLastError (), "sad_dc_kernel launch");
}

extern "C" int
sb2_sad_biref_u8 (const uint8_t *a, int a_stride, const uint8_t *src1, int src1_stride, int weight1,
    const uint8_t *src2, int src2_stride, int weight2, int shift, int width, int height, int *sad, void *stream)
{
  if (!a || !src1 || !src2 || !sad || width < 1 || height < 1 || shift < 1)
    return sb2::set_error (SB2_ERR_ARG, "sb2_sad_biref_u8: bad argument");
  {
    sb2::LaunchScope scope ("sad_biref", 3.0 * width * height, sb2::as_stream (stream));
    sad_biref_kernel<<<1, 32, 0, sb2::as_stream (stream)>>> (a, a_stride, src1, src1_stride, weight1, src2,
        src2_stride, weight2, shift, width, height, sad);
  }
  return sb2::check_cuda (cudaGetLastError (), "sad_biref_kernel launch");
}

extern "C" int
sb2_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, const int64_t *a_offset,
    const int64_t *b_offset, int n, int width, int height, uint32_t *sad, void *stream)
{
  if (!a || !b || !a_offset || !b_offset || !sad || n < 0 || width < 1 || height < 1)
    return sb2::set_error (SB2_ERR_ARG, "sb2_sad_u8: bad argument");
  if (n == 0) return SB2_OK;
  {
    sb2::LaunchScope scope ("sad_batch", 2.0 * width * height * n, sb2::as_stream (stream));
    sad_batch_kernel<<<sb2::ceil_div (n, 4), 128, 0, sb2::as_stream (stream)>>> (a, a_stride, b, b_stride,
        a_offset, b_offset, n, width, height, sad);
  }
  return sb2::check_cuda (cudaGetLastError (), "sad_batch_kernel launch");
}

// workspace: [ticket, 256 B][published words, 8 B per block][static candidates, 64 B per block]
static size_t ws_words_bytes (size_t blocks) { return (blocks * sizeof (unsigned long long) + 255) & ~(size_t) 255; }

extern "C" size_t
sb2_hbm_workspace_bytes (int x_num_blocks, int y_num_blocks, int count)
{
  const size_t blocks = (size_t) y_num_blocks * (size_t) x_num_blocks * (size_t) count;
  return 256 + ws_words_bytes (blocks) + blocks * 64 + 256;
}

// 0: pick by geometry, 1: always the generic one-row-per-CTA kernel (tests force it to cover both)
static int g_force_generic = -1;
extern "C" void sb2_hbm_force_generic (int on) { g_force_generic = on ? 1 : 0; }
static bool force_generic ()
{
  if (g_force_generic < 0) {
    const char *v = getenv ("SB2_HBM_GENERIC");
    g_force_generic = (v && *v && *v != '0') ? 1 : 0;
  }
  return g_force_generic != 0;
}

extern "C" int
sb2_hbm_scan_hint (const sb2_hbm_params *p, const sb2_slab *src_level, const sb2_slab *ref_level,
    int extension, int shift, int h_range, const void *parent_field, void *out_field,
    size_t field_picture_pitch, void *workspace, size_t workspace_bytes, void *stream)
{
  if (!p || !src_level || !ref_level || !out_field)
    return set_error (SB2_ERR_ARG, "sb2_hbm_scan_hint: null argument");
  if (src_level->ncomp != 3 || ref_level->ncomp != 3 || src_level->count != ref_level->count)
    return set_error (SB2_ERR_ARG, "sb2_hbm_scan_hint: need two 3-component slabs of equal count");
  if (shift < 0 || shift > 8 || h_range < 1 || 2 * h_range + 1 > 42)
    return set_error (SB2_ERR_ARG, "sb2_hbm_scan_hint: shift %d / range %d out of range (SCHRO_LIMIT_METRIC_SCAN 42)", shift, h_range);
  if (p->use_chroma && !(p->chroma_h_shift == 1 && p->chroma_v_shift == 1))
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_hbm_scan_hint: chroma ME is defined for 4:2:0 only");
  if (p->ref_index < 0 || p->ref_index > 1) return set_error (SB2_ERR_ARG, "sb2_hbm_scan_hint: ref_index %d", p->ref_index);
  const int count = src_level->count;
  const int skip = 1 << shift;
  HbmArgs A;
  A.src = planeset_from_slab (src_level);
  A.ref = planeset_from_slab (ref_level);
  A.parent = static_cast<const MotionVector *> (parent_field);
  A.field = static_cast<MotionVector *> (out_field);
  A.field_pitch = field_picture_pitch;
  A.width = src_level->width[0];
  A.height = src_level->height[0];
  A.cw = src_level->width[1];
  A.ch = src_level->height[1];
  A.hs = p->chroma_h_shift;
  A.vs = p->chroma_v_shift;
  A.ext = extension;
  A.bw = p->xbsep;
  A.bh = p->ybsep;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.ref_index = p->ref_index;
  A.shift = shift;
  A.h_range = h_range;
  A.use_chroma = p->use_chroma;
  A.rows = ceil_div (A.nby, skip);
  A.cols = ceil_div (A.nbx, skip);
  A.count = count;
  const int split = shift > 1 ? 0 : (shift == 1 ? 1 : 2);
  A.flags0 = (uint32_t) (p->ref_index + 1) | ((uint32_t) split << 3);
  const size_t blocks = (size_t) A.rows * A.cols * count;
  const size_t nwords = blocks;
  const bool wave = !force_generic () && hbm_wave_supported (A, h_range);
  const size_t need = 256 + ws_words_bytes (blocks) + (wave ? hbm_wave_workspace_bytes (A.rows, A.cols, count) : 0);
  if (!workspace || workspace_bytes < need || ((size_t) workspace & 15) != 0)
    return set_error (SB2_ERR_WORKSPACE, "sb2_hbm_scan_hint: workspace %zu < %zu (or not 16-byte aligned)", workspace_bytes, need);
  A.ticket = static_cast<unsigned *> (workspace);
  A.words = reinterpret_cast<unsigned long long *> (static_cast<char *> (workspace) + 256);
  void *stat_ws = static_cast<char *> (workspace) + 256 + ws_words_bytes (blocks);
  cudaStream_t st = as_stream (stream);
  const size_t nfield = (size_t) A.nbx * A.nby;
  if (field_picture_pitch == nfield || count == 1) {
    // contiguous fields: one launch initialises every pair's field
    LaunchScope scope ("hbm_init_field", (double) nfield * 20 * count, st);
    hbm_init_field_kernel<<<(unsigned) min ((size_t) 2048, (nfield * count + 127) / 128), 128, 0, st>>> (
        A.field, nfield * count, A.flags0, A.words, nwords, A.ticket);
  } else {
    for (int pic = 0; pic < count; pic++) {
      LaunchScope scope ("hbm_init_field", (double) nfield * 20, st);
      hbm_init_field_kernel<<<(unsigned) min ((size_t) 1024, (nfield + 127) / 128), 128, 0, st>>> (
          A.field + (size_t) pic * field_picture_pitch, nfield, A.flags0, A.words + pic * (nwords / count),
          nwords / count, A.ticket);
    }
  }
  // algorithmic bytes: both pyramids of this level once + the fields
  double bytes = 0;
  for (int c = 0; c < 3; c++) bytes += 2.0 * src_level->width[c] * src_level->height[c] * count;
  bytes += (double) A.rows * A.cols * 20 * (parent_field ? 2 : 1) * count;
  if (wave) {
    const int rc = hbm_wave_launch (A, h_range, stat_ws, workspace_bytes - (256 + ws_words_bytes (blocks)), st, bytes);
    if (rc != SB2_OK) return rc;
    return check_cuda (cudaGetLastError (), "hbm_wave_kernel launch");
  }
  const int ctas = A.rows * count;
  const int npos = (2 * h_range + 1) * (2 * h_range + 1);
  {
    char tag[48];
    snprintf (tag, sizeof (tag), "hbm_generic_s%d_r%d", shift, h_range);
    LaunchScope scope (tag, bytes, st);
    // warps per block row: enough threads to cover the scan positions in few rounds -- unless the
    // launch has more rows than the GPU can hold at that width (148 SMs x 1024 threads at the 64
    // register cap): then half as many warps per row keeps twice as many rows resident
    int nw = npos <= 64 ? 2 : npos <= 128 ? 4 : npos <= 512 ? 8 : 16;
    while (nw > 2 && (long long) ctas * nw * 32 > 148LL * 1024) nw >>= 1;
    if (nw == 2) hbm_level_kernel<2><<<ctas, 64, 0, st>>> (A);
    else if (nw == 4) hbm_level_kernel<4><<<ctas, 128, 0, st>>> (A);
    else if (nw == 8) hbm_level_kernel<8><<<ctas, 256, 0, st>>> (A);
    else hbm_level_kernel<16><<<ctas, 512, 0, st>>> (A);
  }
  return check_cuda (cudaGetLastError (), "hbm_level_kernel launch");
}

#ifdef SB2_HBM_TRACE
extern "C" int sb2_hbm_trace_read (long long *host, int n)
{
  return (int) cudaMemcpyFromSymbol (host, sb2::g_hbm_trace, sizeof (long long) * n);
}

#endif
