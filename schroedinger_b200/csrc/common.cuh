// common.cuh -- shared helpers for the sm_100a picture-core kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "schro_b200.h"

namespace sb2 {

// error plumbing (cabi.cu)
int set_error (int code, const char *fmt, ...);
int check_cuda (cudaError_t e, const char *what);
void count_launch (unsigned n = 1);
// optional per-launch event timing (cabi.cu); tag names the kernel, bytes = algorithmic bytes
bool profiling ();
int prof_begin (const char *tag, double bytes, cudaStream_t st);
void prof_end (int id, cudaStream_t st);
struct LaunchScope {
  int id; cudaStream_t st;
  LaunchScope (const char *tag, double bytes, cudaStream_t s) : id (prof_begin (tag, bytes, s)), st (s) { count_launch (); }
  ~LaunchScope () { prof_end (id, st); }
};

static inline cudaStream_t as_stream (void *s) { return reinterpret_cast<cudaStream_t> (s); }

// A set of same-shaped planes: `count` pictures x `ncomp` components.
struct PlaneSet {
  char *base;
  size_t pic_pitch;
  size_t off[SB2_MAX_COMPONENTS];
  int stride[SB2_MAX_COMPONENTS];   // bytes
};

static inline PlaneSet planeset_from_slab (const sb2_slab *s)
{
  PlaneSet p;
  p.base = static_cast<char *> (s->base);
  p.pic_pitch = s->picture_pitch;
  for (int i = 0; i < SB2_MAX_COMPONENTS; i++) {
    p.off[i] = i < s->ncomp ? s->offset[i] : 0;
    p.stride[i] = i < s->ncomp ? s->stride[i] : 0;
  }
  return p;
}

__device__ __forceinline__ char *plane_ptr (const PlaneSet &p, int pic, int comp)
{
  return p.base + (size_t) pic * p.pic_pitch + p.off[comp];
}

static inline int ceil_div (int a, int b) { return (a + b - 1) / b; }

// Compact launch grid: blockIdx.x runs over the tiles of every component (so no CTA is launched
// just to find itself outside a smaller chroma plane -- with 4:2:0 that was half of all CTAs),
// blockIdx.y = picture.
struct TileGrid {
  int ncomp;
  int tile_start[SB2_MAX_COMPONENTS + 1];
  int tiles_x[SB2_MAX_COMPONENTS];
  unsigned magic[SB2_MAX_COMPONENTS];    // ceil (2^32 / tiles_x): tile / tiles_x without a software divide
};
struct TilePos { int comp, bx, by; };

static inline dim3 make_tile_grid (TileGrid &g, int ncomp, const int *w, const int *h, int tile_w, int tile_h, int count)
{
  int total = 0;
  g.ncomp = ncomp;
  for (int c = 0; c < SB2_MAX_COMPONENTS; c++) { g.tile_start[c] = 0; g.tiles_x[c] = 1; g.magic[c] = 0; }
  for (int c = 0; c < ncomp; c++) {
    g.tile_start[c] = total;
    g.tiles_x[c] = ceil_div (w[c], tile_w);
    g.magic[c] = (unsigned) ((0x100000000ull + (unsigned) g.tiles_x[c] - 1) / (unsigned) g.tiles_x[c]);
    total += g.tiles_x[c] * ceil_div (h[c], tile_h);
  }
  g.tile_start[ncomp < SB2_MAX_COMPONENTS ? ncomp : SB2_MAX_COMPONENTS] = total;
  return dim3 (total, count, 1);
}

#ifdef __CUDACC__
__device__ __forceinline__ TilePos tile_pos (const TileGrid &g)
{
  int t = blockIdx.x, c = 0;
#pragma unroll
  for (int k = 1; k < SB2_MAX_COMPONENTS; k++)
    if (k < g.ncomp && t >= g.tile_start[k]) c = k;
  t -= g.tile_start[c];
  TilePos p;
  p.comp = c;
  // exact for t * tiles_x < 2^32 (tiles per component stay far below that); tiles_x == 1 -> magic wraps to 0
  p.by = g.tiles_x[c] == 1 ? t : (int) __umulhi ((unsigned) t, g.magic[c]);
  p.bx = t - p.by * g.tiles_x[c];
  return p;
}
#endif

}  // namespace sb2
