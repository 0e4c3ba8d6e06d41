// dequant.cu -- dequantisation of a whole coefficient frame in place.
//
// Bit-exact replacement for what schro_decoder_decode_subband applies codeblock by codeblock
// (schroedinger/schrodecoder.c:3395-3448, 3559-3576): orc_dequantise_s16_ip_2d / _s32_ip_2d
// (schroedinger/schroorc.orc:1154-1168, 2148-2162) on every subband of the in-place layout
// (schro_subband_get_frame_data, schroedinger/schroparams.c:319-352).  With this on the device
// the host uploads QUANTISED coefficients -- what the entropy decoder produces -- and the frame
// never exists dequantised in host memory.
//
// One streaming kernel over the plane: a thread owns four consecutive samples of a row, finds
// each sample's subband from its coordinates (rows: trailing zeros of y; columns: which
// power-of-two slice of the width), its codeblock from the band-local coordinates, and the
// (factor, offset) pair of that codeblock from a small table in global memory (L1-resident).

#include "common.cuh"

namespace sb2 {

struct DequantArgs {
  TileGrid tiles;
  PlaneSet planes;                                 // input (quantised)
  PlaneSet out;                                    // output: the same planes, or an s32 slab when widening
  int w[SB2_MAX_COMPONENTS], h[SB2_MAX_COMPONENTS];
  int ncomp, depth;
  int hcb[SB2_DEQUANT_MAX_LEVELS + 1], vcb[SB2_DEQUANT_MAX_LEVELS + 1];
  int band_base[3 * SB2_DEQUANT_MAX_LEVELS + 1];   // first pair of band `index` inside a component's table
  int comp_pairs;                                  // pairs per component
  const int2 *quant;
  size_t quant_pitch;                              // pairs between pictures
};

__device__ __forceinline__ int dq16 (int v, int factor, int offset)
{
  // copyw, signw, absw, mullw, addw, shrsw 2, mullw: every step wraps at 16 bits
  const int sign = v > 0 ? 1 : v < 0 ? -1 : 0;
  int t = (int) (short) abs (v);
  t = (int) (short) (t * (int) (short) factor);
  t = (int) (short) (t + (int) (short) offset);
  t >>= 2;
  return (int) (short) (t * sign);
}

__device__ __forceinline__ int dq32 (int v, int factor, int offset)
{
  const int sign = v > 0 ? 1 : v < 0 ? -1 : 0;
  unsigned t = v < 0 ? 0u - (unsigned) v : (unsigned) v;
  t = t * (unsigned) factor + (unsigned) offset;
  return (int) ((unsigned) ((int) t >> 2) * (unsigned) sign);
}

// T = sample type of the input; WIDEN: s16 input, s32 arithmetic (orc_dequantise_s32_ip_2d on the
// sign-extended values) and s32 output in a second slab -- the host uploads half the bytes
template <typename T, bool WIDEN>
__global__ void __launch_bounds__ (256)
dequant_kernel (const DequantArgs a)
{
  const TilePos tp = tile_pos (a.tiles);
  const int comp = tp.comp, pic = blockIdx.y;
  const int w = a.w[comp], h = a.h[comp], D = a.depth;
  const int x0 = (tp.bx * blockDim.x + threadIdx.x) * 4, y = tp.by;
  if (x0 >= w || y >= h) return;
  T *row = reinterpret_cast<T *> (plane_ptr (a.planes, pic, comp) + (size_t) y * a.planes.stride[comp]);
  const int2 *q = a.quant + (size_t) pic * a.quant_pitch + (size_t) comp * a.comp_pairs;
  // vertical: a row is the odd row of a pair at shift sv = tz(y) + 1 (vertical high band of that
  // level), or belongs to the coarsest low band when y is a multiple of 2^D (sv = D + 1)
  const int tz = __ffs (y | (1 << D)) - 1;                  // min (trailing zeros of y, D)
  const int sv = tz + 1;
  int v[4];
  const bool vec = x0 + 3 < w && (((size_t) (row + x0)) & (4 * sizeof (T) - 1)) == 0;
  if (vec) {
    if (sizeof (T) == 2) {
      const int2 t = *reinterpret_cast<const int2 *> (row + x0);
      v[0] = (t.x << 16) >> 16; v[1] = t.x >> 16; v[2] = (t.y << 16) >> 16; v[3] = t.y >> 16;
    } else {
      const int4 t = *reinterpret_cast<const int4 *> (row + x0);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = x0 + k < w ? (int) row[x0 + k] : 0;
  }
  // band / codeblock of one sample
  auto locate = [&] (int x, int &pair) {
    // horizontal: columns [w >> s, w >> (s-1)) are the high band of shift s; below w >> D the low band
    int sh = D + 1;
    for (int s = 1; s <= D; s++)
      if (x >= (w >> s)) { sh = s; break; }
    const int s = min (sv, sh);                            // the finer of the two decides the level
    int index, bx, by, bw, bh, nh, nv;
    if (s > D) {
      index = 0; bw = w >> D; bh = h >> D; bx = x; by = y >> D;
      nh = a.hcb[0]; nv = a.vcb[0];
    } else {
      const int level = D - s, orient = (sh == s ? 1 : 0) | (sv == s ? 2 : 0);
      index = 1 + 3 * level + orient - 1;
      bw = w >> s; bh = h >> s;
      bx = x - (sh == s ? bw : 0);
      by = y >> s;
      nh = a.hcb[level + 1]; nv = a.vcb[level + 1];
    }
    // codeblock c covers [(size * c) / n, (size * (c + 1)) / n)
    int cx = 0, cy = 0;
    if (nh > 1) cx = ((bx + 1) * nh - 1) / bw;
    if (nv > 1) cy = ((by + 1) * nv - 1) / bh;
    pair = a.band_base[index] + cy * nh + cx;
  };
  if ((v[0] | v[1] | v[2] | v[3]) != 0) {                  // most quantised coefficients are zero
    int p0, p3;
    locate (x0, p0);
    locate (min (x0 + 3, w - 1), p3);
    if (p0 == p3) {
      // the usual case: the four samples share a codeblock (bands and codeblocks are contiguous
      // in x, so equal ends mean equal everywhere in between)
      const int2 fo = __ldg (q + p0);
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (v[k] != 0) v[k] = (sizeof (T) == 2 && !WIDEN) ? dq16 (v[k], fo.x, fo.y) : dq32 (v[k], fo.x, fo.y);
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (x0 + k >= w || v[k] == 0) continue;            // zero stays zero (the sign factor)
        int pk;
        locate (x0 + k, pk);
        const int2 fo = __ldg (q + pk);
        v[k] = (sizeof (T) == 2 && !WIDEN) ? dq16 (v[k], fo.x, fo.y) : dq32 (v[k], fo.x, fo.y);
      }
    }
  }
  if (WIDEN) {
    int *orow = reinterpret_cast<int *> (plane_ptr (a.out, pic, comp) + (size_t) y * a.out.stride[comp]);
    if (x0 + 3 < w && (((size_t) (orow + x0)) & 15) == 0)
      *reinterpret_cast<int4 *> (orow + x0) = make_int4 (v[0], v[1], v[2], v[3]);
    else
      for (int k = 0; k < 4 && x0 + k < w; k++) orow[x0 + k] = v[k];
    return;
  }
  if (vec) {
    if (sizeof (T) == 2)
      *reinterpret_cast<int2 *> (row + x0) = make_int2 ((v[0] & 0xffff) | (v[1] << 16), (v[2] & 0xffff) | (v[3] << 16));
    else
      *reinterpret_cast<int4 *> (row + x0) = make_int4 (v[0], v[1], v[2], v[3]);
  } else {
    for (int k = 0; k < 4 && x0 + k < w; k++) row[x0 + k] = (T) v[k];
  }
}

static int
dequant_layout (const sb2_dequant_params *p, int *band_base)
{
  int n = 0;
  for (int index = 0; index <= 3 * p->transform_depth; index++) {
    const int level = index == 0 ? 0 : (index - 1) / 3;
    band_base[index] = n;
    n += index == 0 ? p->horiz_codeblocks[0] * p->vert_codeblocks[0]
                    : p->horiz_codeblocks[level + 1] * p->vert_codeblocks[level + 1];
  }
  return n;
}

}  // namespace sb2

using namespace sb2;

extern "C" size_t
sb2_dequant_table_pairs (const sb2_dequant_params *p, int ncomp)
{
  if (!p || p->transform_depth < 1 || p->transform_depth > SB2_DEQUANT_MAX_LEVELS || ncomp < 1) return 0;
  int base[3 * SB2_DEQUANT_MAX_LEVELS + 1];
  return (size_t) dequant_layout (p, base) * (size_t) ncomp;
}

static int
dequantise_launch (const char *who, const sb2_slab *coeffs, const sb2_slab *out, int is_s32, const sb2_dequant_params *p,
    const int32_t *quant, size_t quant_picture_pitch, void *stream)
{
  if (!coeffs || !coeffs->base || !p || !quant) return set_error (SB2_ERR_ARG, "%s: null argument", who);
  if (coeffs->ncomp < 1 || coeffs->ncomp > SB2_MAX_COMPONENTS || coeffs->count < 1)
    return set_error (SB2_ERR_ARG, "%s: bad slab", who);
  if (p->transform_depth < 1 || p->transform_depth > SB2_DEQUANT_MAX_LEVELS)
    return set_error (SB2_ERR_ARG, "%s: transform depth %d", who, p->transform_depth);
  if (out && (!out->base || out->ncomp != coeffs->ncomp || out->count != coeffs->count))
    return set_error (SB2_ERR_ARG, "%s: output slab does not match the input", who);
  const int bpp = is_s32 ? 4 : 2;
  DequantArgs a;
  a.planes = planeset_from_slab (coeffs);
  a.out = planeset_from_slab (out ? out : coeffs);
  a.ncomp = coeffs->ncomp;
  a.depth = p->transform_depth;
  double bytes = 0;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) {
    a.w[c] = a.h[c] = 0;
    if (c >= coeffs->ncomp) continue;
    a.w[c] = coeffs->width[c];
    a.h[c] = coeffs->height[c];
    if (a.w[c] < 1 || a.h[c] < 1 || (a.w[c] & ((1 << a.depth) - 1)) || (a.h[c] & ((1 << a.depth) - 1)))
      return set_error (SB2_ERR_ARG, "%s: component %d size %dx%d is not a multiple of 1<<%d", who, c, a.w[c],
          a.h[c], a.depth);
    if ((coeffs->stride[c] % bpp) || (coeffs->offset[c] % bpp))
      return set_error (SB2_ERR_ARG, "%s: component %d stride/offset not a multiple of the sample size", who, c);
    if (out && (out->width[c] != a.w[c] || out->height[c] != a.h[c] || (out->stride[c] % 4) || (out->offset[c] % 4)))
      return set_error (SB2_ERR_ARG, "%s: output component %d differs in size or is not 4-byte aligned", who, c);
    bytes += (double) a.w[c] * a.h[c] * (out ? 2 + 4 : 2 * bpp) * coeffs->count;
  }
  for (int l = 0; l <= SB2_DEQUANT_MAX_LEVELS; l++) {
    a.hcb[l] = l <= a.depth ? p->horiz_codeblocks[l] : 1;
    a.vcb[l] = l <= a.depth ? p->vert_codeblocks[l] : 1;
    if (a.hcb[l] < 1 || a.vcb[l] < 1) return set_error (SB2_ERR_ARG, "%s: codeblock counts must be >= 1", who);
  }
  a.comp_pairs = dequant_layout (p, a.band_base);
  if (quant_picture_pitch < (size_t) a.comp_pairs * a.ncomp)
    return set_error (SB2_ERR_ARG, "%s: table pitch %zu < %d pairs", who, quant_picture_pitch, a.comp_pairs * a.ncomp);
  a.quant = reinterpret_cast<const int2 *> (quant);
  a.quant_pitch = quant_picture_pitch;
  if (coeffs->count > 65535) return set_error (SB2_ERR_ARG, "%s: at most 65535 pictures per call", who);
  const dim3 grid = make_tile_grid (a.tiles, a.ncomp, a.w, a.h, 4 * 256, 1, coeffs->count);
  cudaStream_t st = as_stream (stream);
  {
    LaunchScope scope (out ? "dequantise_s16_to_s32" : is_s32 ? "dequantise_s32" : "dequantise_s16", bytes, st);
    if (out) dequant_kernel<int16_t, true><<<grid, 256, 0, st>>> (a);
    else if (is_s32) dequant_kernel<int32_t, false><<<grid, 256, 0, st>>> (a);
    else dequant_kernel<int16_t, false><<<grid, 256, 0, st>>> (a);
  }
  return check_cuda (cudaGetLastError (), "dequant_kernel launch");
}

extern "C" int
sb2_dequantise (const sb2_slab *coeffs, int is_s32, const sb2_dequant_params *p, const int32_t *quant,
    size_t quant_picture_pitch, void *stream)
{
  return dequantise_launch ("sb2_dequantise", coeffs, nullptr, is_s32, p, quant, quant_picture_pitch, stream);
}

extern "C" int
sb2_dequantise_widen (const sb2_slab *quantised_s16, const sb2_slab *coeffs_s32, const sb2_dequant_params *p,
    const int32_t *quant, size_t quant_picture_pitch, void *stream)
{
  if (!coeffs_s32) return set_error (SB2_ERR_ARG, "sb2_dequantise_widen: null output slab");
  return dequantise_launch ("sb2_dequantise_widen", quantised_s16, coeffs_s32, 0, p, quant, quant_picture_pitch, stream);
}
