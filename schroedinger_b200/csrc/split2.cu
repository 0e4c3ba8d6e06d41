// split2.cu -- the split-2 pass of the encoder's mode decision for sm_100a.
//
// Bit-exact replacement for schro_do_split2 + schro_motion_copy_to (schroedinger/schromotionest.c:
// 1601-1802, 1511-1523) applied to every superblock, the first step schro_mode_decision (:2587-2685)
// takes for each of them.  Per block inside the picture the candidates are each reference's sub-pel
// vector (luma SAD from the field + chroma SADs at the halved vector, schro_get_split2_metric
// :1527-1594), both vectors together (schro_metric_get_biref, schroedinger/schrometric.c:273-304) and
// a DC block (schro_block_average, :481-516); the cost is entropy + lambda * error, the entropy taken
// against the DECIDED left / up / up-left blocks (schro_motion_block_estimate_entropy :1243-1282,
// schro_motion_vector_prediction schroedinger/schromotion.c:315-368).
//
// The reference does all of it in one loop.  Here, as for the sub-pel refinement (subpel.cu):
//   split2_candidates_kernel  every pixel sum of a block depends on the block's own field entries only:
//                             one warp per block, every block of every picture in parallel;
//   split2_decide_kernel      the decisions form a wavefront: one CTA per picture, one thread per
//                             block row, row j one block behind row j-1, the decided vectors and modes
//                             handed down through a 3-deep shared-memory ring.
// Three properties of the reference are reproduced because results depend on them (oracle_split2.c
// spells them out): a single-reference winner records the luma metric only as its error; with
// mv_precision >= 2 the bi-reference fetches of the three components share one scratch block, so the
// luma metric sees the V prediction in its top-left corner and the U metric is taken against the V
// prediction; with one reference the DC candidate is tried whenever the best error is positive.

#include "obmc_common.cuh"
#include <climits>
#include <math_constants.h>

namespace sb2 {

struct Split2Args {
  PlaneSet orig, ref[2];
  const MotionVector *field[2];
  size_t field_pitch;
  MotionVector *motion;
  size_t motion_pitch;
  int *sb_error, *sb_entropy;       // [count][nsb]
  unsigned *rec;                    // [count][nby * nbx][8]
  int pw[3], ph[3];                 // plane sizes
  int hs, vs, orig_ext, ref_ext;
  int xblen, yblen, nbx, nby, prec, num_refs, count;
  double lambda;
};

constexpr unsigned S2_INSIDE = 1u, S2_BIREF = 2u;

// one sample of the upsampled reference at half-pel (u, v) + pixel offset (a, b); coordinates are
// held inside the reference's border so that a field with wild vectors cannot read outside the slab
// (the reference does not test these fetches; fields produced by the search never need the clamp)
struct RefPlane {
  const uint8_t *p;
  int stride, w, h, ext;
};

__device__ __forceinline__ int half_sample (const RefPlane &r, int u, int v, int a, int b)
{
  const int ph = ((v & 1) << 1) | (u & 1);
  const int x = min (max ((u >> 1) + a, -r.ext), r.w + r.ext - 1), y = min (max ((v >> 1) + b, -r.ext), r.h + r.ext - 1);
  return __ldg (r.p + (ptrdiff_t) ph * (r.stride >> 2) + (ptrdiff_t) y * r.stride + x);
}

// schro_upsampled_frame_get_block_fast_precN (schroedinger/schroframe.c:2458-2482), one pixel
__device__ __forceinline__ int sample (const RefPlane &r, int prec, int x, int y, int a, int b)
{
  if (prec == 0) return half_sample (r, x << 1, y << 1, a, b);
  if (prec == 1) return half_sample (r, x, y, a, b);
  if (prec == 2) { x <<= 1; y <<= 1; }
  const int hx = x >> 2, hy = y >> 2, rx = x & 3, ry = y & 3;
  const int s00 = half_sample (r, hx, hy, a, b);
  if ((rx | ry) == 0) return s00;
  if (ry == 0 && rx == 2) return (s00 + half_sample (r, hx + 1, hy, a, b) + 1) >> 1;
  if (ry == 2 && rx == 0) return (s00 + half_sample (r, hx, hy + 1, a, b) + 1) >> 1;
  const int s01 = half_sample (r, hx + 1, hy, a, b);
  const int s10 = half_sample (r, hx, hy + 1, a, b);
  const int s11 = half_sample (r, hx + 1, hy + 1, a, b);
  return ((4 - ry) * (4 - rx) * s00 + (4 - ry) * rx * s01 + ry * (4 - rx) * s10 + ry * rx * s11 + 8) >> 4;
}

__device__ __forceinline__ unsigned warp_sum (unsigned v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
  return v;
}

// the four taps of a block fetched at (x, y) in units of 2^-prec pixels, as offsets from the plane's pixel (0, 0),
// with the unified weights of obmc_common.cuh; [x0, x1] x [y0, y1] = the pixels the taps of sample (0, 0) touch
__device__ __forceinline__ void s2_blkref (BlkRef &br, int rstride, int prec, int x, int y, int &x0, int &x1, int &y0, int &y1)
{
  const int q = rstride >> 2;
  int hx, hy, rx = 0, ry = 0;
  if (prec == 0) { hx = x << 1; hy = y << 1; }
  else if (prec == 1) { hx = x; hy = y; }
  else {
    if (prec == 2) { x <<= 1; y <<= 1; }
    hx = x >> 2; hy = y >> 2; rx = x & 3; ry = y & 3;
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int u = hx + (t & 1), v = hy + (t >> 1);
    br.o[t] = (((v & 1) << 1) | (u & 1)) * q + (v >> 1) * rstride + (u >> 1);
  }
  const unsigned w00 = (4 - ry) * (4 - rx), w01 = (4 - ry) * rx, w10 = ry * (4 - rx), w11 = ry * rx;
  br.w = w00 | (w01 << 8) | (w10 << 16) | (w11 << 24);
  x0 = hx >> 1; x1 = (hx + 1) >> 1; y0 = hy >> 1; y1 = (hy + 1) >> 1;
}

// one pixel through the unified 4-tap sum, all four taps loaded whatever their weights (the caller has checked
// that they lie inside the plane's border): no branch between the loads of consecutive fetches, so they overlap
__device__ __forceinline__ int s2_fetch (const uint8_t *ref, const BlkRef &br, int pix)
{
  const int s00 = __ldg (ref + br.o[0] + pix), s01 = __ldg (ref + br.o[1] + pix);
  const int s10 = __ldg (ref + br.o[2] + pix), s11 = __ldg (ref + br.o[3] + pix);
  const unsigned w = br.w;
  return ((int) (w & 0xff) * s00 + (int) ((w >> 8) & 0xff) * s01 + (int) ((w >> 16) & 0xff) * s10 + (int) (w >> 24) * s11 + 8) >> 4;
}

// sum within each half of the warp (lanes 0-15, lanes 16-31)
__device__ __forceinline__ unsigned half_sum (unsigned v)
{
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
  return v;
}

template <bool FAST_OK, int MINB>
__global__ void __launch_bounds__ (128, MINB)
split2_candidates_kernel (const Split2Args A)
{
  const int lane = threadIdx.x & 31;
  // (32-bit index arithmetic: the launcher refuses more than 2^31 blocks per launch)
  const unsigned g = blockIdx.x * 4u + (threadIdx.x >> 5);
  const unsigned per_pic = (unsigned) (A.nbx * A.nby);
  if (g >= (unsigned) A.count * per_pic) return;
  const int pic = (int) (g / per_pic), blk = (int) (g - (unsigned) pic * per_pic);
  const int by = blk / A.nbx, bx = blk - by * A.nbx;
  unsigned *rec = A.rec + ((size_t) pic * per_pic + blk) * 8;
  if (!(A.pw[0] > bx * A.xblen) || !(A.ph[0] > by * A.yblen)) {
    if (lane == 0) rec[0] = 0;
    return;
  }
  int cwk[3], chk[3], w[3], h[3], os[3];
  const uint8_t *op[3];
  RefPlane rp[2][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    cwk[k] = k ? A.xblen >> A.hs : A.xblen;
    chk[k] = k ? A.yblen >> A.vs : A.yblen;
    w[k] = min (cwk[k], A.pw[k] - bx * cwk[k]);
    h[k] = min (chk[k], A.ph[k] - by * chk[k]);
    os[k] = A.orig.stride[k];
    op[k] = reinterpret_cast<const uint8_t *> (plane_ptr (A.orig, pic, k)) + (ptrdiff_t) (by * chk[k]) * os[k] + bx * cwk[k];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      rp[r][k].p = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref[r], pic, k));
      rp[r][k].stride = A.ref[r].stride[k];
      rp[r][k].w = A.pw[k];
      rp[r][k].h = A.ph[k];
      rp[r][k].ext = A.ref_ext;
    }
  }
  const MotionVector *f[2] = { A.field[0] + (size_t) pic * A.field_pitch + blk, A.field[1] + (size_t) pic * A.field_pitch + blk };
  unsigned flags = S2_INSIDE, chroma[2] = { 0, 0 }, bi_luma = 0, bi_chroma = 0, dc_error = 0;
  int dc[3];
  // vectors and fetch positions: luma for the bi-reference candidate, chroma for it and for each reference alone
  int vx[2] = { 0, 0 }, vy[2] = { 0, 0 }, lx[2], ly[2], cx[2], cy[2];
  bool have[2] = { false, false };
#pragma unroll
  for (int r = 0; r < 2; r++) {
    if (r < A.num_refs) {
      vx[r] = f[r]->v[r];
      vy[r] = f[r]->v[2 + r];
      have[r] = f[r]->metric != (unsigned) INT_MAX;
    }
    lx[r] = vx[r] + bx * (cwk[0] << A.prec);
    ly[r] = vy[r] + by * (chk[0] << A.prec);
    cx[r] = (vx[r] >> A.hs) + bx * (cwk[1] << A.prec);
    cy[r] = (vy[r] >> A.vs) + by * (chk[1] << A.prec);
  }
  bool biref = A.num_refs > 1;
  if (biref) {
    const int xmin = -A.orig_ext, ymin = -A.orig_ext, xmax = (A.pw[0] << A.prec) + A.orig_ext, ymax = (A.ph[0] << A.prec) + A.orig_ext;
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (xmin > lx[r] || ymin > ly[r] || !(xmax > lx[r] + w[0] - 1) || !(ymax > ly[r] + h[0] - 1)) biref = false;
    if (biref) flags |= S2_BIREF;
  }

  // full 8 x 8 blocks (4 x 4 chroma) whose taps all lie inside the references' borders: fixed lane -> pixel map,
  // tap offsets and weights once per block, every source pixel loaded once
  bool fast = FAST_OK && w[0] == 8 && h[0] == 8 && w[1] == 4 && h[1] == 4 && A.xblen == 8 && A.yblen == 8 &&
      os[1] == os[2] && rp[0][1].stride == rp[0][2].stride && rp[1][1].stride == rp[1][2].stride;
  BlkRef bl[2], bc[2];
  if (fast) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r >= A.num_refs) continue;
      int x0, x1, y0, y1;
      s2_blkref (bc[r], rp[r][1].stride, A.prec, cx[r], cy[r], x0, x1, y0, y1);
      if (x0 < -A.ref_ext || y0 < -A.ref_ext || x1 + 3 > A.pw[1] + A.ref_ext - 1 || y1 + 3 > A.ph[1] + A.ref_ext - 1) fast = false;
      if (biref) {
        s2_blkref (bl[r], rp[r][0].stride, A.prec, lx[r], ly[r], x0, x1, y0, y1);
        if (x0 < -A.ref_ext || y0 < -A.ref_ext || x1 + 7 > A.pw[0] + A.ref_ext - 1 || y1 + 7 > A.ph[0] + A.ref_ext - 1) fast = false;
      }
    }
  }
  if (fast) {
    const int la = lane & 7, lb = lane >> 3;                       // luma: pixels (la, lb) and (la, lb + 4)
    const int ca = lane & 3, cb = (lane >> 2) & 3;                 // chroma: pixel (ca, cb) of U (lanes 0-15) or V (16-31)
    const bool is_v = lane >= 16;
    // every load of the block first ...
    const int ya = __ldg (op[0] + (ptrdiff_t) lb * os[0] + la), yb = __ldg (op[0] + (ptrdiff_t) (lb + 4) * os[0] + la);
    const int c = __ldg ((is_v ? op[2] : op[1]) + (ptrdiff_t) cb * os[1] + ca);
    int pc[2] = { 0, 0 };                                          // this lane's chroma prediction from each reference
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (r < A.num_refs) pc[r] = s2_fetch (is_v ? rp[r][2].p : rp[r][1].p, bc[r], cb * rp[r][1].stride + ca);
    // mv_precision >= 2: the reference's scratch block holds the V prediction where the chroma block lies
    const bool shared_scratch = A.prec >= 2;
    int p0a = 0, p1a = 0, p0b = 0, p1b = 0;
    if (biref) {
      const bool corner = shared_scratch && la < 4;
      p0a = s2_fetch (corner ? rp[0][2].p : rp[0][0].p, corner ? bc[0] : bl[0], corner ? lb * rp[0][1].stride + la : lb * rp[0][0].stride + la);
      p1a = s2_fetch (corner ? rp[1][2].p : rp[1][0].p, corner ? bc[1] : bl[1], corner ? lb * rp[1][1].stride + la : lb * rp[1][0].stride + la);
      p0b = s2_fetch (rp[0][0].p, bl[0], (lb + 4) * rp[0][0].stride + la);
      p1b = s2_fetch (rp[1][0].p, bl[1], (lb + 4) * rp[1][0].stride + la);
    }
    // ... then the sums
    {
      const int ave_y = ((int) warp_sum ((unsigned) (ya + yb)) + 32) >> 6;
      const int ave_c = ((int) half_sum ((unsigned) c) + 8) >> 4;
      dc_error = warp_sum ((unsigned) (abs (ave_y - ya) + abs (ave_y - yb) + abs (ave_c - c)));
      dc[0] = ave_y - 128;
      dc[1] = __shfl_sync (0xffffffffu, ave_c, 0) - 128;
      dc[2] = __shfl_sync (0xffffffffu, ave_c, 16) - 128;
    }
#pragma unroll
    for (int r = 0; r < 2; r++)
      if (r < A.num_refs) chroma[r] = warp_sum ((unsigned) abs (c - pc[r]));
    if (biref) {
      int v0 = pc[0], v1 = pc[1];
      if (shared_scratch) {
        v0 = __shfl_sync (0xffffffffu, pc[0], (lane & 15) + 16);
        v1 = __shfl_sync (0xffffffffu, pc[1], (lane & 15) + 16);
      }
      bi_chroma = warp_sum ((unsigned) abs (c - ((v0 + v1 + 1) >> 1)));
      bi_luma = warp_sum ((unsigned) (abs (ya - ((p0a + p1a + 1) >> 1)) + abs (yb - ((p0b + p1b + 1) >> 1))));
    }
  } else {
    // DC block: rounded mean and the absolute deviations from it, per component
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int n = w[k] * h[k];
      unsigned s = 0;
      for (int p = lane; p < n; p += 32) { const int b = p / w[k], a = p - b * w[k]; s += __ldg (op[k] + (ptrdiff_t) b * os[k] + a); }
      s = warp_sum (s);
      const int ave = ((int) s + n / 2) / n;
      unsigned e = 0;
      for (int p = lane; p < n; p += 32) { const int b = p / w[k], a = p - b * w[k]; e += (unsigned) abs (ave - (int) __ldg (op[k] + (ptrdiff_t) b * os[k] + a)); }
      dc_error += warp_sum (e);
      dc[k] = ave - 128;
    }
    // each reference alone: U + V SADs at the halved vector (skipped, as in the reference, when the field holds no metric)
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (r >= A.num_refs || !have[r]) continue;
      const int n = w[1] * h[1];
      unsigned e = 0;
      for (int p = lane; p < 2 * n; p += 32) {
        const bool v = p >= n;
        const int q = v ? p - n : p;
        const int b = q / w[1], a = q - b * w[1];
        const int o = __ldg ((v ? op[2] : op[1]) + (ptrdiff_t) b * (v ? os[2] : os[1]) + a);
        e += (unsigned) abs (o - (v ? sample (rp[r][2], A.prec, cx[r], cy[r], a, b) : sample (rp[r][1], A.prec, cx[r], cy[r], a, b)));
      }
      chroma[r] = warp_sum (e);
    }
    // both references
    if (biref) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int n = w[k] * h[k];
        unsigned e = 0;
        for (int p = lane; p < n; p += 32) {
          const int b = p / w[k], a = p - b * w[k];
          // the reference's shared scratch block holds the V prediction where the chroma block lies
          const bool v = k == 2 || (A.prec >= 2 && (k == 1 || (a < w[2] && b < h[2])));
          int p0, p1;
          if (v) { p0 = sample (rp[0][2], A.prec, cx[0], cy[0], a, b); p1 = sample (rp[1][2], A.prec, cx[1], cy[1], a, b); }
          else if (k == 1) { p0 = sample (rp[0][1], A.prec, cx[0], cy[0], a, b); p1 = sample (rp[1][1], A.prec, cx[1], cy[1], a, b); }
          else { p0 = sample (rp[0][0], A.prec, lx[0], ly[0], a, b); p1 = sample (rp[1][0], A.prec, lx[1], ly[1], a, b); }
          e += (unsigned) abs ((int) __ldg (op[k] + (ptrdiff_t) b * os[k] + a) - ((p0 + p1 + 1) >> 1));
        }
        e = warp_sum (e);
        if (k == 0) bi_luma = e; else bi_chroma += e;
      }
    }
  }
  if (lane == 0) {
    uint4 *o = reinterpret_cast<uint4 *> (rec);
    o[0] = make_uint4 (flags, chroma[0], chroma[1], bi_luma);
    o[1] = make_uint4 (bi_chroma, dc_error, (unsigned) (dc[0] & 0xffff) | ((unsigned) (dc[1] & 0xffff) << 16), (unsigned) (dc[2] & 0xffff));
  }
}

// schro_pack_estimate_sint (schroedinger/schropack.c:204-226)
__device__ __forceinline__ int s2_bits_sint (int v)
{
  const unsigned a = (unsigned) abs (v);
  const int n = 32 - __clz (a + 1);
  return n + n - 1 + (a ? 1 : 0);
}
__device__ __forceinline__ int s2_med3 (int a, int b, int c) { return max (min (a, b), min (max (a, b), c)); }

constexpr int S2_PF = 8;
__device__ __forceinline__ void prefetch_l1 (const void *p) { asm volatile ("prefetch.global.L1 [%0];" :: "l" (p)); }

// a decided block as its neighbours see it: vector of reference 0, vector of reference 1, pred_mode
struct Decided {
  unsigned v0, v1, mode;
};

// schro_motion_vector_prediction (schroedinger/schromotion.c:315-368) from the three decided neighbours (left, up,
// up-left; `have` says which exist): those predicted from reference `mode` count -- one: its vector, two: their
// rounded mean, three: the median
__device__ __forceinline__ void s2_predict (const Decided &l, const Decided &u, const Decided &ul, unsigned have, int mode, int &px, int &py)
{
  const bool b0 = (have & 1u) && (l.mode & mode), b1 = (have & 2u) && (u.mode & mode), b2 = (have & 4u) && (ul.mode & mode);
  const unsigned w0 = mode == 1 ? l.v0 : l.v1, w1 = mode == 1 ? u.v0 : u.v1, w2 = mode == 1 ? ul.v0 : ul.v1;
  const int x0 = (int) (short) (w0 & 0xffff), y0 = (int) (short) (w0 >> 16);
  const int x1 = (int) (short) (w1 & 0xffff), y1 = (int) (short) (w1 >> 16);
  const int x2 = (int) (short) (w2 & 0xffff), y2 = (int) (short) (w2 >> 16);
  const int n = (int) b0 + (int) b1 + (int) b2;
  const int sx = (b0 ? x0 : 0) + (b1 ? x1 : 0) + (b2 ? x2 : 0), sy = (b0 ? y0 : 0) + (b1 ? y1 : 0) + (b2 ? y2 : 0);
  if (n == 3) { px = s2_med3 (x0, x1, x2); py = s2_med3 (y0, y1, y2); }
  else if (n == 2) { px = (sx + 1) >> 1; py = (sy + 1) >> 1; }
  else { px = sx; py = sy; }
}

__global__ void __launch_bounds__ (1024)
split2_decide_kernel (const Split2Args A)
{
  __shared__ unsigned ring_v0[3][1024], ring_v1[3][1024];
  __shared__ unsigned char ring_mode[3][1024];
  const int pic = blockIdx.x, j = threadIdx.x;
  const int per_pic = A.nbx * A.nby, nsbx = A.nbx >> 2;
  const uint4 *rec = reinterpret_cast<const uint4 *> (A.rec + (size_t) pic * per_pic * 8);
  const MotionVector *f0 = A.field[0] + (size_t) pic * A.field_pitch, *f1 = A.field[1] + (size_t) pic * A.field_pitch;
  MotionVector *motion = A.motion + (size_t) pic * A.motion_pitch;
  int *sb_error = A.sb_error + (size_t) pic * nsbx * (A.nby >> 2), *sb_entropy = A.sb_entropy + (size_t) pic * nsbx * (A.nby >> 2);
  const bool row = j < A.nby;
  const int cw = A.xblen >> A.hs, ch = A.yblen >> A.vs;
  Decided left = { 0, 0, 0 };
  int acc_error = 0, acc_entropy = 0;
  const int steps = A.nbx + A.nby - 1;
  // the next block's record and field entries are in flight while the current block is decided
  uint4 n0 = make_uint4 (0, 0, 0, 0), n1 = n0;
  MotionVector nf[2];
  nf[0].flags = nf[0].metric = nf[0].chroma_metric = 0; nf[0].v[0] = nf[0].v[1] = nf[0].v[2] = nf[0].v[3] = 0;
  nf[1] = nf[0];
  if (row) {
    const int blk = j * A.nbx;
    n0 = __ldg (rec + (size_t) blk * 2); n1 = __ldg (rec + (size_t) blk * 2 + 1);
    nf[0] = f0[blk];
    if (A.num_refs > 1) nf[1] = f1[blk];
    for (int k = 1; k < S2_PF && k < A.nbx; k += 4) {     // four 32-byte records to a line; row j idles j steps before it starts
      prefetch_l1 (rec + (size_t) (blk + k) * 2);
      prefetch_l1 (f0 + blk + k);
      if (A.num_refs > 1) prefetch_l1 (f1 + blk + k);
    }
  }
  for (int s = 0; s < steps; s++) {
    const int i = s - j;
    if (row && i >= 0 && i < A.nbx) {
      const int blk = j * A.nbx + i;
      const uint4 c0 = n0, c1 = n1;
      MotionVector fr[2];
      fr[0] = nf[0];
      fr[1] = nf[1];
      if (i + 1 < A.nbx) {
        n0 = __ldg (rec + (size_t) (blk + 1) * 2); n1 = __ldg (rec + (size_t) (blk + 1) * 2 + 1);
        nf[0] = f0[blk + 1];
        if (A.num_refs > 1) nf[1] = f1[blk + 1];
      }
      // a step lasts as long as its slowest row: without this nearly every step has some row whose next record misses
      // the caches and pays a DRAM round trip; the lines a row needs S2_PF blocks from now are requested here
      if (i + S2_PF < A.nbx) {
        prefetch_l1 (rec + (size_t) (blk + S2_PF) * 2);
        prefetch_l1 (f0 + blk + S2_PF);
        if (A.num_refs > 1) prefetch_l1 (f1 + blk + S2_PF);
      }
      MotionVector best;
      best.flags = (2u << 3) | 1u; best.metric = 0; best.chroma_metric = 0; best.v[0] = best.v[1] = best.v[2] = best.v[3] = 0;
      int best_error = 0, best_entropy = 2;
      if (c0.x & S2_INSIDE) {
        Decided up = { 0, 0, 0 }, ul = { 0, 0, 0 };
        const unsigned have = (i > 0 ? 1u : 0u) | (j > 0 ? 2u : 0u) | (i > 0 && j > 0 ? 4u : 0u);
        if (j > 0) { up.v0 = ring_v0[(s + 2) % 3][j - 1]; up.v1 = ring_v1[(s + 2) % 3][j - 1]; up.mode = ring_mode[(s + 2) % 3][j - 1]; }
        if (i > 0 && j > 0) { ul.v0 = ring_v0[(s + 1) % 3][j - 1]; ul.v1 = ring_v1[(s + 1) % 3][j - 1]; ul.mode = ring_mode[(s + 1) % 3][j - 1]; }
        const int w0 = min (A.xblen, A.pw[0] - i * A.xblen), h0 = min (A.yblen, A.ph[0] - j * A.yblen);
        const int w1 = min (cw, A.pw[1] - i * cw), h1 = min (ch, A.ph[1] - j * ch);
        double min_score = CUDART_INF;
        int entropy[2] = { 0, 0 };
        best_entropy = INT_MAX;
        best_error = INT_MAX;
        MotionVector mv = best;
        const unsigned chroma[2] = { c0.y, c0.z };
#pragma unroll
        for (int r = 0; r < 2; r++) {
          if (r >= A.num_refs) continue;
          mv = fr[r];
          mv.flags = (mv.flags & ~0x1fu) | (2u << 3) | (unsigned) (r + 1);
          int px, py;
          s2_predict (left, up, ul, have, r + 1, px, py);
          entropy[r] = s2_bits_sint (mv.v[r] - px) + s2_bits_sint (mv.v[2 + r] - py);
          int error;
          if (mv.metric == (unsigned) INT_MAX) error = INT_MAX;
          else { mv.chroma_metric = chroma[r]; error = (int) (chroma[r] + mv.metric); }
          const double score = __dadd_rn ((double) entropy[r], __dmul_rn ((double) error, A.lambda));
          if (min_score > score) { min_score = score; best = mv; best_entropy = entropy[r]; best_error = (int) mv.metric; }
        }
        int width0 = 0, height0 = 0, width1 = 0, height1 = 0;
        if (A.num_refs > 1) {
          mv.v[0] = fr[0].v[0]; mv.v[2] = fr[0].v[2]; mv.v[1] = fr[1].v[1]; mv.v[3] = fr[1].v[3];
          mv.flags = (mv.flags & ~0x7u) | 3u;
          width0 = w0; height0 = h0; width1 = w1; height1 = h1;
          if (c0.x & S2_BIREF) {
            mv.metric = c0.w;
            mv.chroma_metric = c1.x;
            const double score = __dadd_rn ((double) (entropy[0] + entropy[1]), __dmul_rn ((double) (mv.metric + mv.chroma_metric), A.lambda));
            if (min_score > score) {
              best_error = (int) (mv.metric + mv.chroma_metric);
              best_entropy = entropy[0] + entropy[1];
              best = mv;
              min_score = score;
            }
          }
        }
        if (4 * (width0 * height0 + 2 * width1 * height1) < best_error) {
          mv.flags = (mv.flags & ~0x1fu) | (2u << 3);
          mv.v[0] = (int16_t) (c1.z & 0xffff); mv.v[1] = (int16_t) (c1.z >> 16); mv.v[2] = (int16_t) (c1.w & 0xffff);
          const int error = (int) c1.y;
          mv.metric = c1.y;
          const int dc_entropy = s2_bits_sint (mv.v[0]) + s2_bits_sint (mv.v[1]) + s2_bits_sint (mv.v[2]);
          if (error < best_error) { best = mv; best_error = error; best_entropy = dc_entropy; }
        }
      }
      motion[blk] = best;
      acc_error = (int) ((unsigned) acc_error + (unsigned) best_error);
      acc_entropy = (int) ((unsigned) acc_entropy + (unsigned) best_entropy);
      if ((i & 3) == 3 || i == A.nbx - 1) {
        atomicAdd (sb_error + (j >> 2) * nsbx + (i >> 2), acc_error);
        atomicAdd (sb_entropy + (j >> 2) * nsbx + (i >> 2), acc_entropy);
        acc_error = acc_entropy = 0;
      }
      const unsigned mode = best.flags & 3u;
      left.v0 = (unsigned) (best.v[0] & 0xffff) | ((unsigned) (best.v[2] & 0xffff) << 16);
      left.v1 = (unsigned) (best.v[1] & 0xffff) | ((unsigned) (best.v[3] & 0xffff) << 16);
      left.mode = mode;
      ring_v0[s % 3][j] = left.v0;
      ring_v1[s % 3][j] = left.v1;
      ring_mode[s % 3][j] = (unsigned char) mode;
    }
    __syncthreads ();
  }
}

}  // namespace sb2

using namespace sb2;

// tests run the candidates kernel both ways: the fixed lane map for full 8 x 8 blocks (default) and the per-pixel
// path every other block takes
static int g_split2_generic = 0;
extern "C" void sb2_split2_force_generic (int on) { g_split2_generic = on ? 1 : 0; }

extern "C" size_t
sb2_split2_workspace_bytes (int x_num_blocks, int y_num_blocks, int count)
{
  return (size_t) x_num_blocks * (size_t) y_num_blocks * (size_t) count * 32;
}

extern "C" int
sb2_split2_decide (const sb2_split2_params *p, const sb2_slab *orig, const sb2_slab *upref0, const sb2_slab *upref1,
    int upref_extension, const void *field0, const void *field1, size_t field_picture_pitch, void *motion,
    size_t motion_picture_pitch, int *sb_error, int *sb_entropy, void *workspace, size_t workspace_bytes, void *stream)
{
  if (!p || !orig || !upref0 || !field0 || !motion || !sb_error || !sb_entropy)
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: null argument");
  if (p->num_refs < 1 || p->num_refs > 2 || (p->num_refs == 2 && (!upref1 || !field1)))
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: num_refs %d needs that many references and fields", p->num_refs);
  if (orig->ncomp != 3 || upref0->ncomp != 3 || orig->count != upref0->count ||
      (p->num_refs == 2 && (upref1->ncomp != 3 || upref1->count != orig->count)))
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: need three-component slabs of equal count");
  if (p->xblen < 1 || p->yblen < 1 || p->x_num_blocks < 4 || p->y_num_blocks < 4 || (p->x_num_blocks & 3) || (p->y_num_blocks & 3) ||
      p->mv_precision < 0 || p->mv_precision > 3 || p->chroma_h_shift < 0 || p->chroma_h_shift > 1 ||
      p->chroma_v_shift < 0 || p->chroma_v_shift > 1 || (p->xblen >> p->chroma_h_shift) < 1 || (p->yblen >> p->chroma_v_shift) < 1)
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: bad parameters");
  if (p->y_num_blocks > 1024)
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_split2_decide: more than 1024 block rows (%d)", p->y_num_blocks);
  for (int k = 0; k < 3; k++) {
    const sb2_slab *refs[2] = { upref0, p->num_refs == 2 ? upref1 : upref0 };
    for (int r = 0; r < 2; r++)
      if (refs[r]->width[k] != orig->width[k] || refs[r]->height[k] != orig->height[k])
        return set_error (SB2_ERR_ARG, "sb2_split2_decide: picture and reference %d differ in size (component %d)", r, k);
  }
  if (orig->width[1] != orig->width[2] || orig->height[1] != orig->height[2] ||
      orig->width[1] != ((orig->width[0] + (1 << p->chroma_h_shift) - 1) >> p->chroma_h_shift) ||
      orig->height[1] != ((orig->height[0] + (1 << p->chroma_v_shift) - 1) >> p->chroma_v_shift))
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: chroma planes do not match the chroma shifts");
  if (upref_extension < 2)
    return set_error (SB2_ERR_ARG, "sb2_split2_decide: the references need a border");
  const size_t need = sb2_split2_workspace_bytes (p->x_num_blocks, p->y_num_blocks, orig->count);
  if (!workspace || workspace_bytes < need || ((size_t) workspace & 15) != 0)
    return set_error (SB2_ERR_WORKSPACE, "sb2_split2_decide: workspace %zu < %zu (or not 16-byte aligned)", workspace_bytes, need);
  Split2Args A;
  A.orig = planeset_from_slab (orig);
  A.ref[0] = planeset_from_slab (upref0);
  A.ref[1] = planeset_from_slab (p->num_refs == 2 ? upref1 : upref0);
  A.field[0] = static_cast<const MotionVector *> (field0);
  A.field[1] = static_cast<const MotionVector *> (p->num_refs == 2 ? field1 : field0);
  A.field_pitch = field_picture_pitch;
  A.motion = static_cast<MotionVector *> (motion);
  A.motion_pitch = motion_picture_pitch;
  A.sb_error = sb_error;
  A.sb_entropy = sb_entropy;
  A.rec = static_cast<unsigned *> (workspace);
  for (int k = 0; k < 3; k++) { A.pw[k] = orig->width[k]; A.ph[k] = orig->height[k]; }
  A.hs = p->chroma_h_shift;
  A.vs = p->chroma_v_shift;
  A.orig_ext = p->orig_extension;
  A.ref_ext = upref_extension;
  A.xblen = p->xblen;
  A.yblen = p->yblen;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.prec = p->mv_precision;
  A.num_refs = p->num_refs;
  A.count = orig->count;
  A.lambda = p->lambda;
  cudaStream_t st = as_stream (stream);
  const size_t nsb = (size_t) (A.nbx >> 2) * (A.nby >> 2) * A.count;
  cudaError_t e = cudaMemsetAsync (sb_error, 0, nsb * sizeof (int), st);
  if (e == cudaSuccess) e = cudaMemsetAsync (sb_entropy, 0, nsb * sizeof (int), st);
  if (e != cudaSuccess) return check_cuda (e, "sb2_split2_decide: clearing the superblock sums");
  const long long warps = (long long) A.nbx * A.nby * A.count;
  if (warps >= (1ll << 31)) return set_error (SB2_ERR_UNSUPPORTED, "sb2_split2_decide: more than 2^31 blocks in one launch");
  {
    // algorithmic bytes: the source once, each reference's four phase planes once (1.5 bytes per luma pixel each), the fields, the records
    const double px = 1.5 * A.pw[0] * A.ph[0] * A.count;
    LaunchScope scope ("split2_candidates", px * (1.0 + 4.0 * A.num_refs) + (20.0 * A.num_refs + 32.0) * A.nbx * A.nby * A.count, st);
    const unsigned grid = (unsigned) ((warps + 3) / 4);
    // the kernel waits for loads most of the time: 8 CTAs per SM at 64 registers (112 bytes spilled) beat 5 at 96 by 20 %
    if (g_split2_generic) split2_candidates_kernel<false, 5><<<grid, 128, 0, st>>> (A);
    else split2_candidates_kernel<true, 8><<<grid, 128, 0, st>>> (A);
  }
  {
    LaunchScope scope ("split2_decide", (32.0 + 20.0 * A.num_refs + 20.0) * A.nbx * A.nby * A.count, st);
    split2_decide_kernel<<<A.count, min (1024, (A.nby + 31) & ~31), 0, st>>> (A);
  }
  return check_cuda (cudaGetLastError (), "split2 kernels launch");
}
