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

}  // namespace sb2
