// subpel.cu -- sub-pel refinement of a motion field for sm_100a.
//
// Bit-exact replacement for schro_encoder_motion_predict_subpel_deep
// (schroedinger/schromotionest.c:246-355) for one reference: for mvprec = 1 .. mv_precision every
// block's vector is doubled and its eight sub-pel neighbours are probed (block fetch at that
// precision, schroedinger/schroframe.c:2287-2482, luma SAD against the source block); a probe
// replaces the vector when entropy + lambda * SAD gets smaller, the entropy being
// schro_pack_estimate_sint of the difference to the median of the ALREADY REFINED left / up /
// up-left vectors (schroedinger/schromotion.c:259-312, schropack.c:204-226).
//
// The reference does this in one raster-order loop.  Here a pass is two kernels:
//   subpel_probe_kernel   the eight probe SADs of a block depend on its own vector only: one warp per
//                         block, every block of every picture in parallel (all the pixel work);
//   subpel_decide_kernel  the decisions form a wavefront (a block needs the refined left, up and
//                         up-left vectors): one CTA per picture, one thread per block row, row j one
//                         block behind row j-1, results handed down through a 3-deep shared-memory
//                         ring, one barrier per step, the next block's record prefetched.
// The score arithmetic is the reference's, in double precision, multiply and add rounded
// separately (the reference is built without FMA contraction).

#include "obmc_common.cuh"
#include <climits>

namespace sb2 {

struct SubpelArgs {
  PlaneSet orig, ref;
  MotionVector *field;
  size_t field_pitch;
  unsigned *rec;                    // [count][nby * nbx][12]: mask, vector, metric, -, err[8]
  int width, height, orig_ext;
  int xblen, yblen, nbx, nby, ref_index, mvprec, count;
  int fast;                         // full 8 x 8 blocks take the word-wide probe path (tests turn it off to run both)
  double lambda;
};

constexpr unsigned SKIP = 0x80000000u;

__device__ __forceinline__ int subpel_sample (const uint8_t *ref, int rstride, int prec, int x, int y, int a, int b)
{
  if (prec == 1) return halfpel (ref, rstride, x, y, a, b);
  if (prec == 2) { x <<= 1; y <<= 1; }
  const int hx = x >> 2, hy = y >> 2, rx = x & 3, ry = y & 3;
  const int s00 = halfpel (ref, rstride, hx, hy, a, b);
  if ((rx | ry) == 0) return s00;
  if (ry == 0 && rx == 2) return (s00 + halfpel (ref, rstride, hx + 1, hy, a, b) + 1) >> 1;
  if (ry == 2 && rx == 0) return (s00 + halfpel (ref, rstride, hx, hy + 1, a, b) + 1) >> 1;
  const int s01 = halfpel (ref, rstride, hx + 1, hy, a, b);
  const int s10 = halfpel (ref, rstride, hx, hy + 1, a, b);
  const int s11 = halfpel (ref, rstride, hx + 1, hy + 1, a, b);
  return ((4 - ry) * (4 - rx) * s00 + (4 - ry) * rx * s01 + ry * (4 - rx) * s10 + ry * rx * s11 + 8) >> 4;
}

__device__ __forceinline__ int probe_dx (int k) { return k < 3 ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6)); }
__device__ __forceinline__ int probe_dy (int k) { return k < 3 ? -1 : (k < 5 ? 0 : 1); }

__global__ void __launch_bounds__ (128)
subpel_probe_kernel (const SubpelArgs A)
{
  const int lane = threadIdx.x & 31;
  // (32-bit index arithmetic: the launcher refuses more than 2^31 blocks per launch)
  const unsigned g = blockIdx.x * 4u + (threadIdx.x >> 5);
  const unsigned per_pic = (unsigned) (A.nbx * A.nby);
  if (g >= (unsigned) A.count * per_pic) return;
  const int pic = (int) (g / per_pic), blk = (int) (g - (unsigned) pic * per_pic);
  const int j = blk / A.nbx, i = blk - j * A.nbx;
  const MotionVector *mv = A.field + (size_t) pic * A.field_pitch + blk;
  unsigned *rec = A.rec + ((size_t) pic * per_pic + blk) * 12;
  const int vx = mv->v[A.ref_index], vy = mv->v[2 + A.ref_index];
  const int w = min (A.xblen, A.width - i * A.xblen), h = min (A.yblen, A.height - j * A.yblen);
  if (w <= 0 || h <= 0) {
    // schro_frame_get_data fails: the block is skipped, its vector stays as it is (:288-291)
    if (lane == 0) { rec[0] = SKIP; rec[1] = (unsigned) (vx & 0xffff) | ((unsigned) (vy & 0xffff) << 16); rec[2] = mv->metric; }
    return;
  }
  const int dvx = (int) (short) (vx << 1), dvy = (int) (short) (vy << 1);
  const int x = i * (A.xblen << A.mvprec) + dvx, y = j * (A.yblen << A.mvprec) + dvy;
  const int x_min = -A.orig_ext, x_max = (A.width << A.mvprec) + A.orig_ext, y_max = (A.height << A.mvprec) + A.orig_ext;
  const uint8_t *op = reinterpret_cast<const uint8_t *> (plane_ptr (A.orig, pic, 0)) + (ptrdiff_t) (j * A.yblen) * A.orig.stride[0] + i * A.xblen;
  const uint8_t *rp = reinterpret_cast<const uint8_t *> (plane_ptr (A.ref, pic, 0));
  const int os = A.orig.stride[0], rs = A.ref.stride[0];
  unsigned mask = 0, my_err = 0;
  if (A.fast && w == 8 && h == 8 && A.xblen == 8 && A.yblen == 8 && ((((size_t) op | (size_t) os) & 7) == 0) &&
      ((((size_t) rp | (size_t) rs) & 3) == 0)) {
    // full 8 x 8 block: four lanes per probe, two rows of eight pixels each; every sub-pel case is the 4-tap sum
    // (w00 s00 + w01 s01 + w10 s10 + w11 s11 + 8) >> 4 on packed 16-bit pairs (obmc_common.cuh: the copy is w00 = 16,
    // the two avgub cases 8 / 8), four pixels per word load, the SAD four bytes at a time
    const int k = lane >> 2, r0 = (lane & 3) * 2;
    const int px = x + probe_dx (k), py = y + probe_dy (k);
    const bool ok = (x_min < px) && (x_max > px + A.xblen - 1) && (x_min < py) && (y_max > py + A.yblen - 1);
    unsigned e = 0;
    if (ok) {
      BlkRef br;
      {
        int qx = px, qy = py, rx = 0, ry = 0, hx = px, hy = py;
        const int q = rs >> 2;
        if (A.mvprec >= 2) {
          if (A.mvprec == 2) { qx <<= 1; qy <<= 1; }
          hx = qx >> 2; hy = qy >> 2; rx = qx & 3; ry = qy & 3;
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int u = hx + (t & 1), v = hy + (t >> 1);
          br.o[t] = (((v & 1) << 1) | (u & 1)) * q + (v >> 1) * rs + (u >> 1);
        }
        const unsigned w00 = (4 - ry) * (4 - rx), w01 = (4 - ry) * rx, w10 = ry * (4 - rx), w11 = ry * rx;
        br.w = w00 | (w01 << 8) | (w10 << 16) | (w11 << 24);
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const uint2 o8 = __ldg (reinterpret_cast<const uint2 *> (op + (ptrdiff_t) (r0 + r) * os));
        const uint2 a = fetch4x4 (rp, br, (r0 + r) * rs), b = fetch4x4 (rp, br, (r0 + r) * rs + 4);
        e += __vsadu4 (o8.x, __byte_perm (a.x, a.y, 0x6420)) + __vsadu4 (o8.y, __byte_perm (b.x, b.y, 0x6420));
      }
    }
    e += __shfl_xor_sync (0xffffffffu, e, 1);
    e += __shfl_xor_sync (0xffffffffu, e, 2);
    mask = __ballot_sync (0xffffffffu, ok && (lane & 3) == 0);
    // bit 4k of the ballot -> bit k
    mask = ((mask >> 0) & 1) | ((mask >> 3) & 2) | ((mask >> 6) & 4) | ((mask >> 9) & 8) | ((mask >> 12) & 16) |
        ((mask >> 15) & 32) | ((mask >> 18) & 64) | ((mask >> 21) & 128);
    my_err = __shfl_sync (0xffffffffu, e, (lane & 7) * 4);          // lane k < 8 takes probe k's sum
  } else {
    const int npix = w * h;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
      const int px = x + probe_dx (k), py = y + probe_dy (k);
      if (!(x_min < px) || !(x_max > px + A.xblen - 1) || !(x_min < py) || !(y_max > py + A.yblen - 1)) continue;   // warp-uniform
      unsigned e = 0;
      for (int p = lane; p < npix; p += 32) {
        const int b = p / w, a = p - b * w;
        e += (unsigned) abs ((int) __ldg (op + (ptrdiff_t) b * os + a) - subpel_sample (rp, rs, A.mvprec, px, py, a, b));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync (0xffffffffu, e, o);
      mask |= 1u << k;
      if (lane == k) my_err = e;
    }
  }
  if (lane < 8) rec[4 + lane] = my_err;
  if (lane == 0) { rec[0] = mask; rec[1] = (unsigned) (dvx & 0xffff) | ((unsigned) (dvy & 0xffff) << 16); rec[2] = mv->metric; }
}

// schro_pack_estimate_sint (schroedinger/schropack.c:204-226)
__device__ __forceinline__ int bits_sint (int v)
{
  const unsigned a = (unsigned) abs (v);
  const int n = 32 - __clz (a + 1);
  return n + n - 1 + (a ? 1 : 0);
}
__device__ __forceinline__ int med3 (int a, int b, int c) { return max (min (a, b), min (max (a, b), c)); }

__global__ void __launch_bounds__ (1024)
subpel_decide_kernel (const SubpelArgs A)
{
  __shared__ unsigned ring[3][1024];
  const int pic = blockIdx.x, j = threadIdx.x;
  const int per_pic = A.nbx * A.nby;
  MotionVector *field = A.field + (size_t) pic * A.field_pitch;
  const uint4 *rec = reinterpret_cast<const uint4 *> (A.rec + (size_t) pic * per_pic * 12);
  const bool row = j < A.nby;
  uint4 n0 = make_uint4 (SKIP, 0, 0, 0), n1 = n0, n2 = n0;
  if (row) { const uint4 *r = rec + (size_t) (j * A.nbx) * 3; n0 = __ldg (r); n1 = __ldg (r + 1); n2 = __ldg (r + 2); }
  unsigned left = 0;
  const int steps = A.nbx + A.nby - 1;
  for (int s = 0; s < steps; s++) {
    const int i = s - j;
    if (row && i >= 0 && i < A.nbx) {
      const uint4 c0 = n0, c1 = n1, c2 = n2;
      if (i + 1 < A.nbx) { const uint4 *r = rec + (size_t) (j * A.nbx + i + 1) * 3; n0 = __ldg (r); n1 = __ldg (r + 1); n2 = __ldg (r + 2); }
      // a step lasts as long as its slowest row: request the line a row needs eight blocks from now, so that the load
      // above hits the L1 instead of paying a DRAM round trip in some row at nearly every step
      if (i + 8 < A.nbx) asm volatile ("prefetch.global.L1 [%0];" :: "l" (rec + (size_t) (j * A.nbx + i + 8) * 3));
      unsigned result = c0.y;
      if (!(c0.x & SKIP)) {
        int dx = (int) (short) (c0.y & 0xffff), dy = (int) (short) (c0.y >> 16);
        int pred_x = 0, pred_y = 0;
        if (i > 0 && j > 0) {
          const unsigned up = ring[(s + 2) % 3][j - 1], ul = ring[(s + 1) % 3][j - 1];
          pred_x = med3 ((int) (short) (left & 0xffff), (int) (short) (up & 0xffff), (int) (short) (ul & 0xffff));
          pred_y = med3 ((int) (short) (left >> 16), (int) (short) (up >> 16), (int) (short) (ul >> 16));
        } else if (i > 0) {
          pred_x = (int) (short) (left & 0xffff); pred_y = (int) (short) (left >> 16);
        } else if (j > 0) {
          const unsigned up = ring[(s + 2) % 3][j - 1];
          pred_x = (int) (short) (up & 0xffff); pred_y = (int) (short) (up >> 16);
        }
        double min_score = __dadd_rn ((double) (bits_sint (dx - pred_x) + bits_sint (dy - pred_y)), __dmul_rn (A.lambda, (double) c0.z));
        int m = -1;
        unsigned min_error = 0;
        const unsigned err[8] = { c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w };
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if ((c0.x >> k) & 1) {
            const int entropy = bits_sint (dx + probe_dx (k) - pred_x) + bits_sint (dy + probe_dy (k) - pred_y);
            const double score = __dadd_rn ((double) entropy, __dmul_rn (A.lambda, (double) (int) err[k]));
            if (min_score > score) { min_score = score; min_error = err[k]; m = k; }
          }
        }
        MotionVector *o = field + j * A.nbx + i;
        if (m >= 0) { dx = (int) (short) (dx + probe_dx (m)); dy = (int) (short) (dy + probe_dy (m)); o->metric = min_error; }
        o->v[A.ref_index] = (int16_t) dx;
        o->v[2 + A.ref_index] = (int16_t) dy;
        result = (unsigned) (dx & 0xffff) | ((unsigned) (dy & 0xffff) << 16);
      }
      left = result;
      ring[s % 3][j] = result;
    }
    __syncthreads ();
  }
}

}  // namespace sb2

using namespace sb2;

// tests run the probe kernel both ways: the word-wide path for full 8 x 8 blocks (default) and the per-pixel path
static int g_subpel_generic = 0;
extern "C" void sb2_subpel_force_generic (int on) { g_subpel_generic = on ? 1 : 0; }

extern "C" size_t
sb2_subpel_workspace_bytes (int x_num_blocks, int y_num_blocks, int count)
{
  return (size_t) x_num_blocks * (size_t) y_num_blocks * (size_t) count * 48;
}

extern "C" int
sb2_subpel_refine (const sb2_subpel_params *p, const sb2_slab *orig, const sb2_slab *upref, int upref_extension,
    void *field, size_t field_picture_pitch, void *workspace, size_t workspace_bytes, void *stream)
{
  if (!p || !orig || !upref || !field) return set_error (SB2_ERR_ARG, "sb2_subpel_refine: null argument");
  if (orig->ncomp < 1 || upref->ncomp < 1 || orig->count != upref->count)
    return set_error (SB2_ERR_ARG, "sb2_subpel_refine: need two slabs of equal count");
  if (p->xblen < 1 || p->yblen < 1 || p->x_num_blocks < 1 || p->y_num_blocks < 1 || p->ref_index < 0 || p->ref_index > 1 ||
      p->mv_precision < 0 || p->mv_precision > 3)
    return set_error (SB2_ERR_ARG, "sb2_subpel_refine: bad parameters");
  if (p->y_num_blocks > 1024)
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_subpel_refine: more than 1024 block rows (%d)", p->y_num_blocks);
  if (upref->width[0] != orig->width[0] || upref->height[0] != orig->height[0])
    return set_error (SB2_ERR_ARG, "sb2_subpel_refine: picture and reference differ in size");
  // every probe that passes the reference's range test (:306-312) must stay inside the upsampled
  // reference's border: the fetch reads pixels (p >> prec) .. (p >> prec) + len (+1 when interpolating)
  for (int prec = 1; prec <= p->mv_precision; prec++) {
    const int lo = (-p->orig_extension + 1) >> prec;                                   // floor: most negative first pixel
    const int hix = ((p->orig_extension - p->xblen) >> prec) + p->xblen, hiy = ((p->orig_extension - p->yblen) >> prec) + p->yblen;
    if (lo < -upref_extension || hix > upref_extension - 1 || hiy > upref_extension - 1)
      return set_error (SB2_ERR_UNSUPPORTED, "sb2_subpel_refine: blocks of %dx%d at precision %d reach beyond the reference's %d-pixel border",
          p->xblen, p->yblen, prec, upref_extension);
  }
  const size_t need = sb2_subpel_workspace_bytes (p->x_num_blocks, p->y_num_blocks, orig->count);
  if (p->mv_precision > 0 && (!workspace || workspace_bytes < need || ((size_t) workspace & 15) != 0))
    return set_error (SB2_ERR_WORKSPACE, "sb2_subpel_refine: workspace %zu < %zu (or not 16-byte aligned)", workspace_bytes, need);
  SubpelArgs A;
  A.orig = planeset_from_slab (orig);
  A.ref = planeset_from_slab (upref);
  A.field = static_cast<MotionVector *> (field);
  A.field_pitch = field_picture_pitch;
  A.rec = static_cast<unsigned *> (workspace);
  A.width = orig->width[0];
  A.height = orig->height[0];
  A.orig_ext = p->orig_extension;
  A.xblen = p->xblen;
  A.yblen = p->yblen;
  A.nbx = p->x_num_blocks;
  A.nby = p->y_num_blocks;
  A.ref_index = p->ref_index;
  A.count = orig->count;
  A.lambda = p->lambda;
  A.fast = g_subpel_generic ? 0 : 1;
  cudaStream_t st = as_stream (stream);
  const long long warps = (long long) A.nbx * A.nby * A.count;
  if (warps >= (1ll << 31)) return set_error (SB2_ERR_UNSUPPORTED, "sb2_subpel_refine: more than 2^31 blocks in one launch");
  // algorithmic bytes of a pass: the source picture once, the reference's four phase planes once, the field twice
  const double bytes = 5.0 * A.width * A.height * A.count + 40.0 * A.nbx * A.nby * A.count;
  for (int prec = 1; prec <= p->mv_precision; prec++) {
    A.mvprec = prec;
    {
      char tag[32];
      snprintf (tag, sizeof (tag), "subpel_probe_p%d", prec);
      LaunchScope scope (tag, bytes, st);
      subpel_probe_kernel<<<(unsigned) ((warps + 3) / 4), 128, 0, st>>> (A);
    }
    {
      LaunchScope scope ("subpel_decide", 68.0 * A.nbx * A.nby * A.count, st);
      subpel_decide_kernel<<<A.count, min (1024, (A.nby + 31) & ~31), 0, st>>> (A);
    }
  }
  return check_cuda (cudaGetLastError (), "subpel kernels launch");
}
