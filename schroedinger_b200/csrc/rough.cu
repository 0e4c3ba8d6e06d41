// rough.cu -- the rough ("bigblock") motion search for sm_100a.
//
// Bit-exact replacement for the two level functions of schro_rough_me_heirarchical_scan
// (schroedinger/schroroughmotion.c:46-60):
//   _nohint (:62-143)  every block of the coarsest level: full search of the window
//                      schro_metric_scan_setup (0, 0, distance) gives it; first strict minimum in
//                      the reference's x-outer / y-inner order (gravity = the first position).
//   _hint   (:145-300) candidates zero / four nearest parents / left, up, up-left of the SAME
//                      level, ranked by luma SAD (first strict minimum), then a scan of
//                      `distance` around the winner in which the winner wins ties.
//
// _nohint has no dependency between blocks -- the one SAD workload of the codec that is pure
// throughput.  A warp owns a block; a lane owns a window COLUMN and walks down the reference
// rows once: every 8-byte reference row segment it assembles is matched against all eight
// source rows (kept in registers) and accumulated into the eight window positions it belongs
// to, so a position costs 16 VABSDIFF4 and the loads amortise over 8 positions.
// _hint is a wavefront like hierarchical block matching (hbm.cu): a warp owns a block row, polls
// the row above's published vectors, rows are handed out by an atomic ticket.

#include "hbm_common.cuh"

namespace sb2 {

struct RoughArgs {
  PlaneSet src, ref;
  const MotionVector *parent;
  MotionVector *field;
  size_t field_pitch;
  unsigned long long *words;        // [count][rows][cols] published vectors (hint only)
  unsigned *ticket;
  int width, height, ext;           // luma size of this pyramid level, its edge extension
  int bw, bh, nbx, nby, ref_index, shift, distance;
  int rows, cols, count;
  int staged;                       // full search: stage the window in shared memory (tests turn it off to run both paths)
};

struct RoughWin { int xmin, ymin, scan_w, scan_h; };

// schro_metric_scan_setup (schroedinger/schrometric.c:174-214)
__device__ __forceinline__ RoughWin
rough_window (const RoughArgs &A, int x, int y, int bw, int bh, int dx, int dy)
{
  RoughWin w;
  w.xmin = max (max (-bw, x + dx - A.distance), -A.ext);
  w.ymin = max (max (-bh, y + dy - A.distance), -A.ext);
  const int xmax = min (min (A.width, x + dx + A.distance), A.width - bw + A.ext);
  const int ymax = min (min (A.height, y + dy + A.distance), A.height - bh + A.ext);
  w.scan_w = xmax - w.xmin + 1;
  w.scan_h = ymax - w.ymin + 1;
  return w;
}

__device__ __forceinline__ unsigned sad_acc (unsigned a, unsigned b, unsigned c)
{
  unsigned r;
  asm ("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

__global__ void __launch_bounds__ (128)
rough_init_kernel (MotionVector *field, size_t n, unsigned long long *words, size_t nwords, unsigned *ticket)
{
  if (ticket && blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nwords; i += (size_t) gridDim.x * blockDim.x)
    words[i] = 0;
  // schro_motion_field_set (mf, 0, 1) (schroedinger/schromotionest.c:416-432)
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    MotionVector m;
    m.flags = 1;
    m.metric = 0;
    m.chroma_metric = 0;
    m.v[0] = m.v[1] = m.v[2] = m.v[3] = 0;
    field[i] = m;
  }
}

// The scan of one block by one warp: key = (SAD << 32 | a << 8 | b) of the best position, the minimum
// over keys being the reference's first strict minimum in a-outer / b-inner order.  MAXH = the
// largest window height the register path holds; taller windows and partial blocks take the
// position-per-lane path.
template <int MAXH>
__device__ __forceinline__ unsigned long long
rough_scan_warp (const uint8_t *sblk, int ss, const uint8_t *rp, int rs, const RoughWin &wn, int bw, int bh,
    bool aligned, int lane)
{
  unsigned long long best = ~0ull;
  if (aligned && bw == 8 && bh == 8 && wn.scan_h <= MAXH) {
    uint2 srow[8];
#pragma unroll
    for (int r = 0; r < 8; r++) srow[r] = __ldg (reinterpret_cast<const uint2 *> (sblk + (ptrdiff_t) r * ss));
    const int rsw = rs >> 2;
    for (int a0 = 0; a0 < wn.scan_w; a0 += 32) {
      const int a = min (a0 + lane, wn.scan_w - 1);           // idle lanes repeat the last column
      const RowRef rr = row_ref (rp + (ptrdiff_t) wn.ymin * rs + wn.xmin + a);
      unsigned acc[MAXH];
#pragma unroll
      for (int j = 0; j < MAXH; j++) acc[j] = 0;
#pragma unroll
      for (int rho = 0; rho < MAXH + 7; rho++) {
        if (rho < wn.scan_h + 7) {                             // warp-uniform: rows the window has
          const uint2 bv = row_load8 (rr, rho * rsw);
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int j = rho - r;
            if (j >= 0 && j < MAXH) acc[j] = sad_acc (srow[r].y, bv.y, sad_acc (srow[r].x, bv.x, acc[j]));
          }
        }
      }
      if (a0 + lane < wn.scan_w) {
#pragma unroll
        for (int j = 0; j < MAXH; j++) {
          const unsigned long long key = ((unsigned long long) acc[j] << 32) | ((unsigned) a << 8) | (unsigned) j;
          if (j < wn.scan_h && key < best) best = key;
        }
      }
    }
  } else {
    const int npos = wn.scan_w * wn.scan_h;
    for (int p = lane; p < npos; p += 32) {
      const int a = p / wn.scan_h, b = p - a * wn.scan_h;
      const unsigned m = block_sad (sblk, ss, rp + (ptrdiff_t) (wn.ymin + b) * rs + wn.xmin + a, rs, bw, bh);
      const unsigned long long key = ((unsigned long long) m << 32) | ((unsigned) a << 8) | (unsigned) b;
      if (key < best) best = key;
    }
  }
  return warp_min64 (best);
}

// The same scan with the reference window staged in shared memory: lane = window ROW for the copy
// (up to three aligned 16-byte pieces of its row, one memory round trip for the whole window), lane =
// window COLUMN for the arithmetic, which then reads its 8-byte row segments from shared memory (lanes
// read consecutive bytes: no bank conflicts).  For full 8 x 8 blocks and windows of at most 25 x 25
// positions on 16-byte aligned rows; key = (SAD << 16 | a << 8 | b) fits 32 bits (SAD <= 16320).
constexpr int RS_PITCH = 48;                              // bytes per staged row: 15 + 25 + 7 <= 48
template <int MAXH>
__device__ __forceinline__ unsigned
rough_scan_warp_staged (const uint8_t *sblk, int ss, const uint8_t *rp, int rs, const RoughWin &wn, unsigned *buf, int lane)
{
  uint2 srow[8];
#pragma unroll
  for (int r = 0; r < 8; r++) srow[r] = __ldg (reinterpret_cast<const uint2 *> (sblk + (ptrdiff_t) r * ss));
  const uint8_t *win = rp + (ptrdiff_t) wn.ymin * rs + wn.xmin;
  const unsigned off = (unsigned) ((size_t) win & 15);
  const int need = (int) off + wn.scan_w + 7;               // bytes of a staged row that are read below
  if (lane < wn.scan_h + 7) {
    const uint4 *g = reinterpret_cast<const uint4 *> (win - off + (ptrdiff_t) lane * rs);
    uint4 *d = reinterpret_cast<uint4 *> (buf + lane * (RS_PITCH / 4));
    const uint4 v0 = __ldg (g);
    uint4 v1 = make_uint4 (0, 0, 0, 0), v2 = v1;
    if (need > 16) v1 = __ldg (g + 1);
    if (need > 32) v2 = __ldg (g + 2);
    d[0] = v0; d[1] = v1; d[2] = v2;
  }
  __syncwarp ();
  const int a = min (lane, wn.scan_w - 1);                  // idle lanes repeat the last column
  const unsigned bo = off + (unsigned) a;
  const unsigned *w = buf + (bo >> 2);
  const unsigned sh = (bo & 3) * 8;
  unsigned acc[MAXH];
#pragma unroll
  for (int j = 0; j < MAXH; j++) acc[j] = 0;
  if (wn.scan_h == MAXH) {
    // the interior block (the whole window fits the frame): straight-line code, no row guards
#pragma unroll
    for (int rho = 0; rho < MAXH + 7; rho++) {
      const unsigned w0 = w[rho * (RS_PITCH / 4)], w1 = w[rho * (RS_PITCH / 4) + 1], w2 = w[rho * (RS_PITCH / 4) + 2];
      const unsigned bx = __funnelshift_r (w0, w1, sh), by = __funnelshift_r (w1, w2, sh);
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int j = rho - r;
        if (j >= 0 && j < MAXH) acc[j] = sad_acc (srow[r].y, by, sad_acc (srow[r].x, bx, acc[j]));
      }
    }
  } else {
#pragma unroll
    for (int rho = 0; rho < MAXH + 7; rho++) {
      if (rho < wn.scan_h + 7) {                            // warp-uniform: rows the window has
        const unsigned w0 = w[rho * (RS_PITCH / 4)], w1 = w[rho * (RS_PITCH / 4) + 1], w2 = w[rho * (RS_PITCH / 4) + 2];
        const unsigned bx = __funnelshift_r (w0, w1, sh), by = __funnelshift_r (w1, w2, sh);
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int j = rho - r;
          if (j >= 0 && j < MAXH) acc[j] = sad_acc (srow[r].y, by, sad_acc (srow[r].x, bx, acc[j]));
        }
      }
    }
  }
  unsigned best = 0xffffffffu;
  if (lane < wn.scan_w) {
    const unsigned low = (unsigned) a << 8;
#pragma unroll
    for (int j = 0; j < MAXH; j++) {
      const unsigned key = acc[j] * 65536u + (low + (unsigned) j);
      if (j < wn.scan_h) best = min (best, key);
    }
  }
  __syncwarp ();
  return __reduce_min_sync (0xffffffffu, best);
}

template <int MAXH>
__global__ void __launch_bounds__ (128)
rough_full_kernel (const RoughArgs A)
{
  __shared__ __align__ (16) unsigned stage[4][MAXH <= 25 ? 32 * RS_PITCH / 4 : 4];
  const int lane = threadIdx.x & 31;
  const long long g = (long long) blockIdx.x * 4 + (threadIdx.x >> 5);
  const int per_pic = A.rows * A.cols;
  if (g >= (long long) A.count * per_pic) return;
  const int pic = (int) (g / per_pic), rem = (int) (g - (long long) pic * per_pic);
  const int bj = rem / A.cols, bi = rem - bj * A.cols;
  const int x = bi * A.bw, y = bj * A.bh;
  const int bw = min (A.width - x, A.bw), bh = min (A.height - y, A.bh);
  const RoughWin wn = rough_window (A, x, y, bw, bh, 0, 0);
  MotionVector *o = A.field + (size_t) pic * A.field_pitch + (size_t) (bj << A.shift) * A.nbx + (bi << A.shift);
  if (wn.scan_w <= 0 || wn.scan_h <= 0) {
    // (:105-109) clears dx[0] / dy[0] whatever the reference index
    if (lane == 0) { o->v[0] = 0; o->v[2] = 0; o->metric = (uint32_t) INT_MAX; }
    return;
  }
  const uint8_t *sp = reinterpret_cast<const uint8_t *> (plane_ptr (A.src, pic, 0));
  const uint8_t *rp = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref, pic, 0));
  const int ss = A.src.stride[0], rs = A.ref.stride[0];
  const bool aligned = ((((size_t) sp | (size_t) ss | (size_t) x) & 7) == 0) && ((((size_t) rp | (size_t) rs) & 3) == 0);
  int a, b;
  unsigned metric;
  if (MAXH <= 25 && aligned && A.staged && bw == 8 && bh == 8 && wn.scan_w <= 25 && wn.scan_h <= MAXH &&
      ((((size_t) (rp - A.ext)) | (size_t) rs) & 15) == 0) {
    const unsigned k = rough_scan_warp_staged<(MAXH <= 25 ? MAXH : 1)> (sp + (ptrdiff_t) y * ss + x, ss, rp, rs, wn,
        stage[threadIdx.x >> 5], lane);
    metric = k >> 16; a = (int) ((k >> 8) & 0xff); b = (int) (k & 0xff);
  } else {
    const unsigned long long k = rough_scan_warp<MAXH> (sp + (ptrdiff_t) y * ss + x, ss, rp, rs, wn, bw, bh, aligned, lane);
    metric = (unsigned) (k >> 32); a = (int) ((k >> 8) & 0xffff); b = (int) (k & 0xff);
  }
  if (lane == 0) {
    o->metric = metric;
    o->v[A.ref_index] = (int16_t) ((wn.xmin + a - x) << A.shift);
    o->v[2 + A.ref_index] = (int16_t) ((wn.ymin + b - y) << A.shift);
  }
}

// One warp = one block row of one (picture, reference) pair.
__global__ void __launch_bounds__ (32)
rough_hint_kernel (const RoughArgs A)
{
  const int lane = threadIdx.x;
  unsigned t = 0;
  if (lane == 0) t = atomicAdd (A.ticket, 1u);
  t = __shfl_sync (0xffffffffu, t, 0);
  const int row = (int) (t / (unsigned) A.count), pic = (int) (t % (unsigned) A.count);
  const int s = A.shift, skip = 1 << s, j = row * skip, ri = A.ref_index;
  const uint8_t *sp = reinterpret_cast<const uint8_t *> (plane_ptr (A.src, pic, 0));
  const uint8_t *rp = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref, pic, 0));
  const int ss = A.src.stride[0], rs = A.ref.stride[0];
  MotionVector *mf = A.field + (size_t) pic * A.field_pitch;
  const MotionVector *pf = A.parent + (size_t) pic * A.field_pitch;
  unsigned long long *words_me = A.words + ((size_t) pic * A.rows + row) * A.cols;
  const unsigned long long *words_up = row > 0 ? words_me - A.cols : nullptr;
  const int hint_mask = ~((1 << (s + 1)) - 1);
  const int y = (j * A.bh) >> s;
  const int bh = min (A.height - y, A.bh), h = min (A.bh, max (0, A.height - y));
  const bool base_aligned = ((((size_t) sp | (size_t) ss) & 7) == 0) && ((((size_t) rp | (size_t) rs) & 3) == 0);
  int left_dx = 0, left_dy = 0;

  for (int bi = 0; bi < A.cols; bi++) {
    const int i = bi * skip;
    const int x = (i * A.bw) >> s;
    const int bw = min (A.width - x, A.bw), w = min (A.bw, max (0, A.width - x));
    // ---- candidates, one per lane: 0 zero, 1..4 parents, 5 left, 6 up, 7 up-left (:186-216)
    int cdx = 0, cdy = 0;
    bool valid = lane == 0;
    if (lane >= 1 && lane <= 4) {
      const int m = lane - 1;
      const int l = (i + skip * (-1 + 2 * (m & 1))) & hint_mask;
      const int k = (j + skip * (-1 + (m & 2))) & hint_mask;
      if (l >= 0 && l < A.nbx && k >= 0 && k < A.nby) {
        const MotionVector *p = pf + (size_t) k * A.nbx + l;
        cdx = p->v[ri]; cdy = p->v[2 + ri]; valid = true;
      }
    } else if (lane == 5 && i > 0) {
      cdx = left_dx; cdy = left_dy; valid = true;
    }
    // up / up-left: every lane polls the same two words (warp-uniform wait, see hbm_wave.cu)
    if (words_up) {
      unsigned long long wu = ld_word (words_up + bi), wl = bi > 0 ? ld_word (words_up + bi - 1) : (1ull << 63);
      unsigned ns = 32, polls = 0;
      while (!((wu & wl) >> 63)) {
        __nanosleep (ns);
        if (ns < 1024) ns <<= 1;
        if (++polls > (1u << 22)) __trap ();       // seconds of waiting on a row that has started: an error, not a hang
        wu = ld_word (words_up + bi);
        if (bi > 0) wl = ld_word (words_up + bi - 1);
      }
      if (lane == 6) { cdx = (int) (short) (wu >> 16); cdy = (int) (short) wu; valid = true; }
      if (lane == 7 && bi > 0) { cdx = (int) (short) (wl >> 16); cdy = (int) (short) wl; valid = true; }
    }
    // ---- rank by luma SAD (:222-255): a lane per candidate
    unsigned key = 0xffffffffu;
    {
      const int rx = (i * A.bw + cdx) >> s, ry = (j * A.bh + cdy) >> s;
      const bool ok = valid && lane < 8 && rx >= 0 && ry >= 0 && w != 0 && h != 0 &&
          max (0, A.width - rx) >= w && max (0, A.height - ry) >= h;
      if (ok) {
        const unsigned m = block_sad (sp + (ptrdiff_t) y * ss + x, ss, rp + (ptrdiff_t) ry * rs + rx, rs, w, h);
        if (m < (unsigned) INT_MAX) key = (m << 3) | (unsigned) lane;     // SAD <= 255 * bw * bh, far below 2^28
      }
    }
    key = __reduce_min_sync (0xffffffffu, key);
    // the surviving lane order equals the reference's list order, so min (SAD, lane) is its first strict minimum
    const int best_lane = key == 0xffffffffu ? 0 : (int) (key & 7);
    const int dx = __shfl_sync (0xffffffffu, cdx, best_lane) >> s, dy = __shfl_sync (0xffffffffu, cdy, best_lane) >> s;

    // ---- scan around the winner (:257-299)
    const RoughWin wn = rough_window (A, x, y, bw, bh, dx, dy);
    int rdx, rdy;
    unsigned metric;
    if (wn.scan_w <= 0 || wn.scan_h <= 0) {
      rdx = rdy = 0; metric = (unsigned) INT_MAX;
    } else if (bw <= 0 || bh <= 0) {
      // a block outside its frame: the reference's seed lies outside the window (oracle_rough.c)
      rdx = dx << s; rdy = dy << s; metric = 0;
    } else {
      const bool aligned = base_aligned && (x & 7) == 0;
      unsigned long long k = rough_scan_warp<9> (sp + (ptrdiff_t) y * ss + x, ss, rp, rs, wn, bw, bh, aligned, lane);
      // the seed wins ties: it is re-inserted with a smaller tie-break field than any scanned position
      const int sa = x + dx - wn.xmin, sb = y + dy - wn.ymin;
      unsigned seed_m = 0;
      if (lane == 0)
        seed_m = block_sad (sp + (ptrdiff_t) y * ss + x, ss, rp + (ptrdiff_t) (y + dy) * rs + x + dx, rs, bw, bh);
      seed_m = __shfl_sync (0xffffffffu, seed_m, 0);
      int a = (int) ((k >> 8) & 0xffff), b = (int) (k & 0xff);
      if (seed_m <= (unsigned) (k >> 32)) { a = sa; b = sb; }
      metric = min (seed_m, (unsigned) (k >> 32));
      rdx = (wn.xmin + a - x) << s; rdy = (wn.ymin + b - y) << s;
    }
    if (lane == 0) {
      st_word (words_me + bi, pack_word ((int16_t) rdx, (int16_t) rdy));
      MotionVector *o = mf + (size_t) j * A.nbx + i;
      o->metric = metric;
      o->v[ri] = (int16_t) rdx;
      o->v[2 + ri] = (int16_t) rdy;
    }
    left_dx = (int16_t) rdx; left_dy = (int16_t) rdy;
  }
}

}  // namespace sb2

using namespace sb2;

// 1: the full search stages its windows in shared memory (default), 0: every row segment from global memory
static int g_rough_staged = 1;
extern "C" void sb2_rough_force_unstaged (int on) { g_rough_staged = on ? 0 : 1; }

static size_t rough_words_bytes (size_t blocks) { return (blocks * sizeof (unsigned long long) + 255) & ~(size_t) 255; }

extern "C" size_t
sb2_rough_workspace_bytes (int x_num_blocks, int y_num_blocks, int count)
{
  return 256 + rough_words_bytes ((size_t) x_num_blocks * (size_t) y_num_blocks * (size_t) count);
}

static int
rough_args (RoughArgs &A, const char *who, const sb2_hbm_params *p, const sb2_slab *src_level, const sb2_slab *ref_level,
    int extension, int shift, int distance, const void *parent_field, void *out_field, size_t field_picture_pitch)
{
  if (!p || !src_level || !ref_level || !out_field) return set_error (SB2_ERR_ARG, "%s: null argument", who);
  if (src_level->ncomp < 1 || ref_level->ncomp < 1 || src_level->count != ref_level->count)
    return set_error (SB2_ERR_ARG, "%s: need two slabs of equal count", who);
  if (shift < 0 || shift > 8 || distance < 1 || 2 * distance + 1 > 42)
    return set_error (SB2_ERR_ARG, "%s: shift %d / distance %d out of range (SCHRO_LIMIT_METRIC_SCAN 42)", who, shift, distance);
  if (p->ref_index < 0 || p->ref_index > 1) return set_error (SB2_ERR_ARG, "%s: ref_index %d", who, p->ref_index);
  if (p->xbsep < 1 || p->ybsep < 1 || p->x_num_blocks < 1 || p->y_num_blocks < 1)
    return set_error (SB2_ERR_ARG, "%s: bad block geometry", who);
  A.src = planeset_from_slab (src_level);
  A.ref = planeset_from_slab (ref_level);
  A.parent = static_cast<const MotionVector *> (parent_field);
  A.field = static_cast<MotionVector *> (out_field);
  A.field_pitch = field_picture_pitch;
  A.words = nullptr;
  A.ticket = nullptr;
  A.width = src_level->width[0];
  A.height = src_level->height[0];
  A.ext = extension;
  A.bw = p->xbsep;
  A.bh = p->ybsep;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.ref_index = p->ref_index;
  A.shift = shift;
  A.distance = distance;
  A.rows = ceil_div (A.nby, 1 << shift);
  A.cols = ceil_div (A.nbx, 1 << shift);
  A.count = src_level->count;
  A.staged = g_rough_staged;
  return SB2_OK;
}

static void
rough_init (const RoughArgs &A, size_t nwords, cudaStream_t st)
{
  const size_t nfield = (size_t) A.nbx * A.nby;
  if (A.field_pitch == nfield || A.count == 1) {
    LaunchScope scope ("rough_init_field", (double) nfield * 20 * A.count, st);
    rough_init_kernel<<<(unsigned) min ((size_t) 2048, (nfield * A.count + 127) / 128), 128, 0, st>>> (
        A.field, nfield * A.count, A.words, nwords, A.ticket);
  } else {
    for (int pic = 0; pic < A.count; pic++) {
      LaunchScope scope ("rough_init_field", (double) nfield * 20, st);
      rough_init_kernel<<<(unsigned) min ((size_t) 1024, (nfield + 127) / 128), 128, 0, st>>> (
          A.field + (size_t) pic * A.field_pitch, nfield, A.words ? A.words + pic * (nwords / A.count) : nullptr,
          A.words ? nwords / A.count : 0, A.ticket);
    }
  }
}

static double
rough_bytes (const sb2_slab *src_level, const RoughArgs &A, bool parent)
{
  // both luma planes of this level once + the fields
  return 2.0 * src_level->width[0] * src_level->height[0] * A.count + (double) A.rows * A.cols * 20 * (parent ? 2 : 1) * A.count;
}

extern "C" int
sb2_rough_scan_nohint (const sb2_hbm_params *p, const sb2_slab *src_level, const sb2_slab *ref_level, int extension,
    int shift, int distance, void *out_field, size_t field_picture_pitch, void *stream)
{
  RoughArgs A;
  const int rc = rough_args (A, "sb2_rough_scan_nohint", p, src_level, ref_level, extension, shift, distance, nullptr,
      out_field, field_picture_pitch);
  if (rc != SB2_OK) return rc;
  cudaStream_t st = as_stream (stream);
  rough_init (A, 0, st);
  const long long warps = (long long) A.rows * A.cols * A.count;
  const unsigned ctas = (unsigned) ((warps + 3) / 4);
  {
    char tag[48];
    snprintf (tag, sizeof (tag), "rough_nohint_s%d_d%d", shift, distance);
    LaunchScope scope (tag, rough_bytes (src_level, A, false), st);
    if (distance <= 4) rough_full_kernel<9><<<ctas, 128, 0, st>>> (A);
    else if (distance <= 12) rough_full_kernel<25><<<ctas, 128, 0, st>>> (A);
    else rough_full_kernel<41><<<ctas, 128, 0, st>>> (A);
  }
  return check_cuda (cudaGetLastError (), "rough_full_kernel launch");
}

extern "C" int
sb2_rough_scan_hint (const sb2_hbm_params *p, const sb2_slab *src_level, const sb2_slab *ref_level, int extension,
    int shift, int distance, const void *parent_field, void *out_field, size_t field_picture_pitch, void *workspace,
    size_t workspace_bytes, void *stream)
{
  RoughArgs A;
  const int rc = rough_args (A, "sb2_rough_scan_hint", p, src_level, ref_level, extension, shift, distance, parent_field,
      out_field, field_picture_pitch);
  if (rc != SB2_OK) return rc;
  if (!parent_field) return set_error (SB2_ERR_ARG, "sb2_rough_scan_hint: the hint level needs the field of level shift+1");
  const size_t blocks = (size_t) A.rows * A.cols * A.count;
  const size_t need = 256 + rough_words_bytes (blocks);
  if (!workspace || workspace_bytes < need || ((size_t) workspace & 15) != 0)
    return set_error (SB2_ERR_WORKSPACE, "sb2_rough_scan_hint: workspace %zu < %zu (or not 16-byte aligned)", workspace_bytes, need);
  A.ticket = static_cast<unsigned *> (workspace);
  A.words = reinterpret_cast<unsigned long long *> (static_cast<char *> (workspace) + 256);
  cudaStream_t st = as_stream (stream);
  rough_init (A, blocks, st);
  {
    char tag[48];
    snprintf (tag, sizeof (tag), "rough_hint_s%d_d%d", shift, distance);
    LaunchScope scope (tag, rough_bytes (src_level, A, true), st);
    rough_hint_kernel<<<(unsigned) (A.rows * A.count), 32, 0, st>>> (A);
  }
  return check_cuda (cudaGetLastError (), "rough_hint_kernel launch");
}
