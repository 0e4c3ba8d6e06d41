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

namespace sb2 {

struct FrameArgs {
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
  const int side = 2 * ext * (h + 2 * ext);       // left+right strips, all rows
  const int caps = 2 * ext * w;                   // top+bottom caps, interior columns
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < side + caps; i += gridDim.x * blockDim.x) {
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
constexpr int UT_W = 64;     // output tile (extended coordinates)
constexpr int UT_H = 16;
constexpr int UP_W = UT_W + 8;            // phase-0 tile incl. 3 left / 4 right taps (+1 pad)
constexpr int UP_H = UT_H + 7;

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

__global__ void __launch_bounds__ (256)
upsample_kernel (const FrameArgs a)
{
  __shared__ uint8_t s0[UP_H][UP_W];      // phase 0 at clamped coordinates
  __shared__ uint8_t sv[UT_H][UP_W];      // vertical half-pel of the same columns

  const int comp = blockIdx.z % a.ncomp, pic = blockIdx.z / a.ncomp;
  const int w = a.w[comp], h = a.h[comp], ext = a.ext;
  const int x0 = blockIdx.x * UT_W - ext, y0 = blockIdx.y * UT_H - ext;
  if (x0 >= w + ext || y0 >= h + ext) return;
  uint8_t *p0 = reinterpret_cast<uint8_t *> (plane_ptr (a.planes, pic, comp));
  const int stride = a.planes.stride[comp];
  const int q = stride >> 2;

  // tile of phase 0: rows y0-3 .. y0+UT_H+3, columns x0-3 .. x0+UT_W+4, coordinates clamped
  for (int i = threadIdx.x; i < UP_H * UP_W; i += blockDim.x) {
    const int ty = i / UP_W, tx = i % UP_W;
    const int yy = clampi (y0 + ty - 3, 0, h - 1), xx = clampi (x0 + tx - 3, 0, w - 1);
    s0[ty][tx] = p0[(ptrdiff_t) yy * stride + xx];
  }
  __syncthreads ();
  // vertical phase of every tile column; rows >= h-1 are copies of the source row h-1,
  // rows < 0 are never used as filter input (they take other sources below)
  for (int i = threadIdx.x; i < UT_H * UP_W; i += blockDim.x) {
    const int ty = i / UP_W, tx = i % UP_W;
    const int y = y0 + ty;
    int v;
    if (y >= h - 1 || y < 0) v = s0[ty + 3][tx];     // s0 row ty+3 is clamp(y): row h-1 (or 0)
    else v = taps8 (&s0[ty][tx], UP_W);
    sv[ty][tx] = (uint8_t) v;
  }
  __syncthreads ();

  for (int i = threadIdx.x; i < UT_H * UT_W; i += blockDim.x) {
    const int ty = i / UT_W, tx = i % UT_W;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w + ext || y >= h + ext) continue;
    // phase 1 (schroframe.c:2022-2024): horizontal filter of phase 0 row clamp(y); left
    // border = phase 0 column 0, columns >= w-1 = phase 0 column w-1
    int v1;
    if (x < 0 || x >= w - 1) v1 = s0[ty + 3][tx + 3];
    else v1 = taps8 (&s0[ty + 3][tx], 1);
    // phase 2 (:2018-2020): vertical filter; rows above = phase 0 row 0, last row and
    // below = phase 0 row h-1, side borders replicate phase 2 itself
    const int v2 = sv[ty][tx + 3];
    // phase 3 (:2026-2028): horizontal filter of phase 2; rows above / last row and below
    // are copies of phase 1, side borders come from phase 2
    int v3;
    if (y < 0 || y >= h - 1) v3 = v1;
    else if (x < 0 || x >= w - 1) v3 = v2;
    else v3 = taps8 (&sv[ty][tx], 1);
    uint8_t *o = p0 + (ptrdiff_t) y * stride + x;
    if (a.fuse_edge && (x < 0 || x >= w || y < 0 || y >= h)) o[0] = s0[ty + 3][tx + 3];
    o[q] = (uint8_t) v1;
    o[2 * q] = (uint8_t) v2;
    o[3 * q] = (uint8_t) v3;
  }
}

// ---- downsample ---------------------------------------------------------------------
struct DownArgs {
  PlaneSet src, dst;
  int sw[SB2_MAX_COMPONENTS], sh[SB2_MAX_COMPONENTS];
  int dw[SB2_MAX_COMPONENTS], dh[SB2_MAX_COMPONENTS];
  int ncomp;
  int dst_ext;                  // > 0: also replicate the result into dst's border
};

constexpr int DT_W = 32, DT_H = 8;                 // output tile
constexpr int DS_W = 2 * DT_W + 2;                 // source columns 2x-1 .. 2x+2

__global__ void __launch_bounds__ (256)
downsample_kernel (const DownArgs a)
{
  __shared__ uint8_t sm[DT_H][DS_W];               // vertically filtered rows (8-bit intermediate)

  const int comp = blockIdx.z % a.ncomp, pic = blockIdx.z / a.ncomp;
  const int sw = a.sw[comp], sh = a.sh[comp], dw = a.dw[comp], dh = a.dh[comp];
  const int x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
  if (x0 >= dw || y0 >= dh) return;
  const uint8_t *s = reinterpret_cast<const uint8_t *> (plane_ptr (a.src, pic, comp));
  uint8_t *d = reinterpret_cast<uint8_t *> (plane_ptr (a.dst, pic, comp));
  const int ss = a.src.stride[comp], dstr = a.dst.stride[comp];

  // orc_downsample_vert_u8 (schroorc.orc:1345-1368): (6(a+d) + 26(b+c) + 32) >> 6, u8
  for (int i = threadIdx.x; i < DT_H * DS_W; i += blockDim.x) {
    const int ty = i / DS_W, tx = i % DS_W;
    const int y = y0 + ty;
    const int xx = clampi (2 * x0 - 1 + tx, 0, sw - 1);
    const int r0 = s[(ptrdiff_t) clampi (2 * y - 1, 0, sh - 1) * ss + xx];
    const int r1 = s[(ptrdiff_t) clampi (2 * y, 0, sh - 1) * ss + xx];
    const int r2 = s[(ptrdiff_t) clampi (2 * y + 1, 0, sh - 1) * ss + xx];
    const int r3 = s[(ptrdiff_t) clampi (2 * y + 2, 0, sh - 1) * ss + xx];
    sm[ty][tx] = (uint8_t) ((6 * (r0 + r3) + 26 * (r1 + r2) + 32) >> 6);
  }
  __syncthreads ();
  // horizontal pass (schroframe.c:1449-1485)
  for (int i = threadIdx.x; i < DT_H * DT_W; i += blockDim.x) {
    const int ty = i / DT_W, tx = i % DT_W;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= dw || y >= dh) continue;
    const uint8_t *t = &sm[ty][2 * tx];
    const uint8_t v = (uint8_t) clamp255 ((6 * ((int) t[0] + t[3]) + 26 * ((int) t[1] + t[2]) + 32) >> 6);
    d[(ptrdiff_t) y * dstr + x] = v;
    if (a.dst_ext > 0 && (x == 0 || x == dw - 1 || y == 0 || y == dh - 1)) {
      // schro_frame_mc_edgeextend of the result: the thread owning an edge pixel writes its
      // replicas (rows / columns outside the plane, corners included)
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
  dim3 grid (min (ceil_div (maxn, 256), 1024), 1, frames->ncomp * frames->count);
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
  dim3 grid (ceil_div (maxw, UT_W), ceil_div (maxh, UT_H), frames->ncomp * frames->count);
  {
    LaunchScope scope ("upsample", bytes, as_stream (stream));
    upsample_kernel<<<grid, 256, 0, as_stream (stream)>>> (a);
  }
  return check_cuda (cudaGetLastError (), "upsample_kernel launch");
}

static int downsample_impl (const sb2_slab *src, const sb2_slab *dst, int dst_ext, void *stream);

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
  dim3 grid (ceil_div (maxw, DT_W), ceil_div (maxh, DT_H), src->ncomp * src->count);
  {
    LaunchScope scope ("downsample", bytes, as_stream (stream));
    downsample_kernel<<<grid, 256, 0, as_stream (stream)>>> (a);
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
