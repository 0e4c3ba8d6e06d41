// glue.cu -- combine / convert glue between the picture-core stages.
//
// Bit-exact replacements for the planar, equal-chroma-format part of
//   schro_frame_convert             (schroedinger/schroframe.c:870-978)
//   schro_frame_add / _subtract     (schroedinger/schroframe.c:1012-1182)
// Depth conversion follows the Orc programs the library actually runs (schroorc-dist.c:
// orc_offsetconvert_u8_s16 / _u8_s32 / _s16_u8 / _s32_u8, orc_convert_s16_s32 / _s32_s16); crop and
// edge extension (schrovirtframe.c:1824-1960) collapse into one index clamp:
//   dest(x, y) = conv (src (min (x, sw-1), min (y, sh-1))).
// Pure streaming kernels: four destination samples per thread, vector loads / stores when the
// rows allow it.  HBM bound.

#include "common.cuh"

namespace sb2 {

struct GlueArgs {
  TileGrid tiles;
  PlaneSet src, dst;
  int sw[SB2_MAX_COMPONENTS], sh[SB2_MAX_COMPONENTS];
  int dw[SB2_MAX_COMPONENTS], dh[SB2_MAX_COMPONENTS];
  int ncomp;
};

template <int D> struct Sample;
template <> struct Sample<0> { typedef uint8_t T; };
template <> struct Sample<1> { typedef int16_t T; };
template <> struct Sample<2> { typedef int32_t T; };

// one sample through the reference's converter chain (schroframe.c:905-925 picks the pair)
template <int SD, int DD>
__device__ __forceinline__ int convert_sample (int v)
{
  if (SD == DD) return v;
  if (DD == 0) {
    int t;
    if (SD == 1) {
      t = (int) (short) (v + 128);                          // addw wraps, then convsuswb
    } else {
      // addl wraps; convsuslw saturates to 0..65535; convsuswb reads that word as signed
      const int w = (int) ((unsigned) v + 128u);
      t = (int) (short) min (max (w, 0), 65535);
    }
    return min (max (t, 0), 255);
  }
  if (DD == 1) return SD == 0 ? v - 128 : (int) (short) v;  // convubw + subw / convlw truncates
  return SD == 0 ? v - 128 : v;                             // convubw + subw + convswl / convswl
}

constexpr int GLUE_GROUPS = 4;      // rows per thread (four samples of each): enough bytes in flight per SM to
                                    // cover the HBM latency whatever the picture width

template <int SD>
__device__ __forceinline__ void load4 (const typename Sample<SD>::T *srow, int x, int sw, int (&v)[4])
{
  typedef typename Sample<SD>::T TS;
  // one vector load when the four samples exist and the address allows it
  if (x + 3 < sw && (((size_t) (srow + x)) & (4 * sizeof (TS) - 1)) == 0) {
    if (SD == 0) {
      const unsigned q = *reinterpret_cast<const unsigned *> (srow + x);
      v[0] = q & 0xff; v[1] = (q >> 8) & 0xff; v[2] = (q >> 16) & 0xff; v[3] = q >> 24;
    } else if (SD == 1) {
      const int2 q = *reinterpret_cast<const int2 *> (srow + x);
      v[0] = (q.x << 16) >> 16; v[1] = q.x >> 16; v[2] = (q.y << 16) >> 16; v[3] = q.y >> 16;
    } else {
      const int4 q = *reinterpret_cast<const int4 *> (srow + x);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (int) srow[min (x + k, sw - 1)];
  }
}

template <int DD>
__device__ __forceinline__ void store4 (typename Sample<DD>::T *drow, int x, int dw, const int (&v)[4])
{
  typedef typename Sample<DD>::T TD;
  if (x + 3 < dw && (((size_t) (drow + x)) & (4 * sizeof (TD) - 1)) == 0) {
    if (DD == 0)
      *reinterpret_cast<unsigned *> (drow + x) = (unsigned) v[0] | ((unsigned) v[1] << 8) | ((unsigned) v[2] << 16) | ((unsigned) v[3] << 24);
    else if (DD == 1)
      *reinterpret_cast<int2 *> (drow + x) = make_int2 ((v[0] & 0xffff) | (v[1] << 16), (v[2] & 0xffff) | (v[3] << 16));
    else
      *reinterpret_cast<int4 *> (drow + x) = make_int4 (v[0], v[1], v[2], v[3]);
  } else {
    for (int k = 0; k < 4 && x + k < dw; k++) drow[x + k] = (TD) v[k];
  }
}

template <int SD, int DD>
__global__ void __launch_bounds__ (256)
convert_kernel (const GlueArgs a)
{
  typedef typename Sample<SD>::T TS;
  typedef typename Sample<DD>::T TD;
  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int dw = a.dw[comp], dh = a.dh[comp], sw = a.sw[comp], sh = a.sh[comp];
  const int x = (tp.bx * blockDim.x + threadIdx.x) * 4, y0 = tp.by * GLUE_GROUPS;
  if (x >= dw || y0 >= dh) return;
  const char *sbase = plane_ptr (a.src, pic, comp);
  char *dbase = plane_ptr (a.dst, pic, comp);
  const size_t ss = (size_t) a.src.stride[comp], ds = (size_t) a.dst.stride[comp];
  int v[GLUE_GROUPS][4];
#pragma unroll
  for (int g = 0; g < GLUE_GROUPS; g++)
    if (y0 + g < dh) load4<SD> (reinterpret_cast<const TS *> (sbase + (size_t) min (y0 + g, sh - 1) * ss), x, sw, v[g]);
#pragma unroll
  for (int g = 0; g < GLUE_GROUPS; g++) {
    if (y0 + g < dh) {
#pragma unroll
      for (int k = 0; k < 4; k++) v[g][k] = convert_sample<SD, DD> (v[g][k]);
      store4<DD> (reinterpret_cast<TD *> (dbase + (size_t) (y0 + g) * ds), x, dw, v[g]);
    }
  }
}

// dest (s16) +-= src (u8 or s16) over the common area; addw / subw wrap at 16 bits
template <int SD>
__global__ void __launch_bounds__ (256)
add_kernel (const GlueArgs a, int subtract)
{
  typedef typename Sample<SD>::T TS;
  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int w = min (a.dw[comp], a.sw[comp]), h = min (a.dh[comp], a.sh[comp]);
  const int x = (tp.bx * blockDim.x + threadIdx.x) * 4, y0 = tp.by * GLUE_GROUPS;
  if (x >= w || y0 >= h) return;
  const char *sbase = plane_ptr (a.src, pic, comp);
  char *dbase = plane_ptr (a.dst, pic, comp);
  const size_t ss = (size_t) a.src.stride[comp], ds = (size_t) a.dst.stride[comp];
  int sv[GLUE_GROUPS][4], dv[GLUE_GROUPS][4];
#pragma unroll
  for (int g = 0; g < GLUE_GROUPS; g++) {
    if (y0 + g < h) {
      load4<SD> (reinterpret_cast<const TS *> (sbase + (size_t) (y0 + g) * ss), x, w, sv[g]);
      load4<1> (reinterpret_cast<const int16_t *> (dbase + (size_t) (y0 + g) * ds), x, w, dv[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < GLUE_GROUPS; g++) {
    if (y0 + g < h) {
#pragma unroll
      for (int k = 0; k < 4; k++) dv[g][k] = subtract ? dv[g][k] - sv[g][k] : dv[g][k] + sv[g][k];
      store4<1> (reinterpret_cast<int16_t *> (dbase + (size_t) (y0 + g) * ds), x, w, dv[g]);   // the 16-bit store wraps like addw / subw
    }
  }
}

static int
glue_args (GlueArgs &a, const sb2_slab *src, const sb2_slab *dst, const char *who, int src_bpp, int dst_bpp)
{
  if (!src || !dst || !src->base || !dst->base) return set_error (SB2_ERR_ARG, "%s: null slab", who);
  if (src->ncomp < 1 || src->ncomp > SB2_MAX_COMPONENTS || src->ncomp != dst->ncomp || src->count != dst->count ||
      src->count < 1)
    return set_error (SB2_ERR_ARG, "%s: slab shapes differ (ncomp %d/%d count %d/%d)", who, src->ncomp, dst->ncomp,
        src->count, dst->count);
  a.src = planeset_from_slab (src);
  a.dst = planeset_from_slab (dst);
  a.ncomp = src->ncomp;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    a.sw[c] = a.sh[c] = a.dw[c] = a.dh[c] = 0;
    if (c >= src->ncomp) continue;
    a.sw[c] = src->width[c]; a.sh[c] = src->height[c];
    a.dw[c] = dst->width[c]; a.dh[c] = dst->height[c];
    if (a.sw[c] < 1 || a.sh[c] < 1 || a.dw[c] < 1 || a.dh[c] < 1)
      return set_error (SB2_ERR_ARG, "%s: bad component %d", who, c);
    if ((src->stride[c] % src_bpp) || (src->offset[c] % src_bpp) || (dst->stride[c] % dst_bpp) || (dst->offset[c] % dst_bpp))
      return set_error (SB2_ERR_ARG, "%s: component %d stride/offset not a multiple of the sample size", who, c);
  }
  return SB2_OK;
}

template <int SD, int DD>
static void launch_convert (const GlueArgs &a, const dim3 &grid, cudaStream_t st) { convert_kernel<SD, DD><<<grid, 256, 0, st>>> (a); }

}  // namespace sb2

namespace sb2 {
// schro_frame_shift_left / _right (schroedinger/schroframe.c:1238-1291): in place, every sample of
// every component.  left: shlw (16-bit wrap); right: addw / addl of (1 << shift) >> 1, then an
// arithmetic shift (orc_lshift_s16_ip, orc_add_const_rshift_s16 / _s32, schroorc.orc:146-163, 207-218).
struct ShiftArgs {
  TileGrid tiles;
  PlaneSet p;
  int w[SB2_MAX_COMPONENTS], h[SB2_MAX_COMPONENTS];
  int shift, right;
};

template <typename T>
__global__ void __launch_bounds__ (256)
shift_kernel (const ShiftArgs a)
{
  const TilePos t = tile_pos (a.tiles);
  const int comp = t.comp, pic = blockIdx.y;
  const int x0 = (t.bx * 256 + threadIdx.x) * 4, y = t.by;
  if (x0 >= a.w[comp] || y >= a.h[comp]) return;
  T *row = reinterpret_cast<T *> (plane_ptr (a.p, pic, comp) + (size_t) y * a.p.stride[comp]);
  const int n = min (4, a.w[comp] - x0);
  const int rnd = (1 << a.shift) >> 1;
  for (int k = 0; k < n; k++) {
    const int v = (int) row[x0 + k];
    int r;
    if (!a.right) r = (int) ((unsigned) v << a.shift);
    else if (sizeof (T) == 2) r = (int) (short) (v + rnd) >> a.shift;
    else r = (int) ((unsigned) v + (unsigned) rnd) >> a.shift;
    row[x0 + k] = (T) r;
  }
}
}  // namespace sb2

using namespace sb2;

extern "C" int
sb2_frame_shift (const sb2_slab *frames, int depth, int shift, int right, void *stream)
{
  if (!frames || !frames->base || frames->ncomp < 1 || frames->ncomp > SB2_MAX_COMPONENTS || frames->count < 1 || frames->count > 65535)
    return set_error (SB2_ERR_ARG, "sb2_frame_shift: bad slab");
  if (depth < 1 || depth > 2 || shift < 0 || shift > (depth == 1 ? 15 : 31))
    return set_error (SB2_ERR_ARG, "sb2_frame_shift: depth %d (1: s16, 2: s32) / shift %d", depth, shift);
  if (!right && depth != 1)
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_frame_shift: the left shift is s16 only (as the reference, schroframe.c:1238)");
  ShiftArgs a;
  a.p = planeset_from_slab (frames);
  a.shift = shift;
  a.right = right;
  double bytes = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    a.w[c] = c < frames->ncomp ? frames->width[c] : 0;
    a.h[c] = c < frames->ncomp ? frames->height[c] : 0;
    bytes += 2.0 * a.w[c] * a.h[c] * (depth == 1 ? 2 : 4) * frames->count;
  }
  const dim3 grid = make_tile_grid (a.tiles, frames->ncomp, a.w, a.h, 4 * 256, 1, frames->count);
  cudaStream_t st = as_stream (stream);
  {
    LaunchScope scope (right ? "frame_shift_right" : "frame_shift_left", bytes, st);
    if (depth == 1) shift_kernel<int16_t><<<grid, 256, 0, st>>> (a);
    else shift_kernel<int32_t><<<grid, 256, 0, st>>> (a);
  }
  return check_cuda (cudaGetLastError (), "shift_kernel launch");
}

extern "C" int
sb2_frame_convert (const sb2_slab *src, int src_depth, const sb2_slab *dst, int dst_depth, void *stream)
{
  if (src_depth < 0 || src_depth > 2 || dst_depth < 0 || dst_depth > 2)
    return set_error (SB2_ERR_ARG, "sb2_frame_convert: depth codes are 0 (u8), 1 (s16), 2 (s32)");
  GlueArgs a;
  const int bpp[3] = { 1, 2, 4 };
  int rc = glue_args (a, src, dst, "sb2_frame_convert", bpp[src_depth], bpp[dst_depth]);
  if (rc) return rc;
  int maxw = 0, maxh = 0;
  double bytes = 0;
  for (int c = 0; c < a.ncomp; c++) {
    maxw = max (maxw, a.dw[c]);
    maxh = max (maxh, a.dh[c]);
    bytes += (double) a.dw[c] * a.dh[c] * (bpp[src_depth] + bpp[dst_depth]) * src->count;
  }
  (void) maxw; (void) maxh;
  if (src->count > 65535) return set_error (SB2_ERR_ARG, "sb2_frame_convert: at most 65535 pictures per call");
  const dim3 grid = make_tile_grid (a.tiles, a.ncomp, a.dw, a.dh, 4 * 256, GLUE_GROUPS, src->count);
  cudaStream_t st = as_stream (stream);
  {
    LaunchScope scope ("frame_convert", bytes, st);
    switch (src_depth * 3 + dst_depth) {
      case 0: launch_convert<0, 0> (a, grid, st); break;
      case 1: launch_convert<0, 1> (a, grid, st); break;
      case 2: launch_convert<0, 2> (a, grid, st); break;
      case 3: launch_convert<1, 0> (a, grid, st); break;
      case 4: launch_convert<1, 1> (a, grid, st); break;
      case 5: launch_convert<1, 2> (a, grid, st); break;
      case 6: launch_convert<2, 0> (a, grid, st); break;
      case 7: launch_convert<2, 1> (a, grid, st); break;
      default: launch_convert<2, 2> (a, grid, st); break;
    }
  }
  return check_cuda (cudaGetLastError (), "convert_kernel launch");
}

extern "C" int
sb2_frame_add (const sb2_slab *dst, const sb2_slab *src, int src_depth, int subtract, void *stream)
{
  if (src_depth < 0 || src_depth > 1)
    return set_error (SB2_ERR_UNSUPPORTED, "sb2_frame_add: the source is u8 (0) or s16 (1), as in the reference's tables "
        "(schroframe.c:984-1047)");
  GlueArgs a;
  int rc = glue_args (a, src, dst, "sb2_frame_add", src_depth ? 2 : 1, 2);
  if (rc) return rc;
  int maxw = 0, maxh = 0;
  double bytes = 0;
  for (int c = 0; c < a.ncomp; c++) {
    const int w = min (a.dw[c], a.sw[c]), h = min (a.dh[c], a.sh[c]);
    maxw = max (maxw, w);
    maxh = max (maxh, h);
    bytes += (double) w * h * (4 + (src_depth ? 2 : 1)) * src->count;
  }
  (void) maxw; (void) maxh;
  int cw[SB2_MAX_COMPONENTS], chh[SB2_MAX_COMPONENTS];
  for (int c = 0; c < a.ncomp; c++) { cw[c] = min (a.dw[c], a.sw[c]); chh[c] = min (a.dh[c], a.sh[c]); }
  if (src->count > 65535) return set_error (SB2_ERR_ARG, "sb2_frame_add: at most 65535 pictures per call");
  const dim3 grid = make_tile_grid (a.tiles, a.ncomp, cw, chh, 4 * 256, GLUE_GROUPS, src->count);
  cudaStream_t st = as_stream (stream);
  {
    LaunchScope scope (subtract ? "frame_subtract" : "frame_add", bytes, st);
    if (src_depth) add_kernel<1><<<grid, 256, 0, st>>> (a, subtract);
    else add_kernel<0><<<grid, 256, 0, st>>> (a, subtract);
  }
  return check_cuda (cudaGetLastError (), "add_kernel launch");
}
