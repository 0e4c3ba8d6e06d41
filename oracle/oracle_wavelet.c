/*
 * oracle_wavelet.c -- CPU restatement of the seven Dirac integer lifting
 * wavelets, s16 and s32, forward and inverse.  TEST INFRASTRUCTURE (oracle.h).
 * Follows schroedinger/schrowaveletorc.c:60-2668 and the Orc kernels in
 * schroedinger/schroorc.orc it calls; see oracle_wavelet_body.inc.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* lifting-step kinds used by the step tables in oracle_wavelet_body.inc */
enum { K_A22, K_A11, K_M4, K_M2, K_F8A, K_F8B, K_COPY, K_HALF };

/* ---- s16: storage wraps at 16 bits, mas2/mas4 accumulate in int32 ---- */
#define T int16_t
#define WIDE int32_t
#define NAME(x) x##_s16
#define WR(x) ((int32_t)(int16_t)(x))
#define ACC(x) ((int32_t)(x))
#define ACC32(x) ((int32_t)(uint32_t)(x))
#include "oracle_wavelet_body.inc"
#undef T
#undef WIDE
#undef NAME
#undef WR
#undef ACC
#undef ACC32

/* ---- s32: everything wraps at 32 bits, avgsl is computed in 64 bits ---- */
#define T int32_t
#define WIDE int64_t
#define NAME(x) x##_s32
#define WR(x) ((int64_t)(int32_t)(uint32_t)(uint64_t)(x))
#define ACC(x) WR(x)
#define ACC32(x) WR(x)
#include "oracle_wavelet_body.inc"

void
oracle_wavelet_fwd (void *data, int stride, int width, int height, int is_s32,
    int filter)
{
  if (is_s32) fwd_s32 (data, stride, width, height, filter);
  else fwd_s16 (data, stride, width, height, filter);
}

void
oracle_wavelet_inv (void *data, int stride, int width, int height, int is_s32,
    int filter)
{
  if (is_s32) inv_s32 (data, stride, width, height, filter);
  else inv_s16 (data, stride, width, height, filter);
}

/* level l works on (width>>l) x (height>>l) with stride<<l
 * (schroedinger/schroframe.c:1214-1223) */
void
oracle_iwt_fwd (void *data, int stride, int width, int height, int is_s32,
    int filter, int depth)
{
  int level;
  for (level = 0; level < depth; level++)
    oracle_wavelet_fwd (data, stride << level, width >> level, height >> level,
        is_s32, filter);
}

/* schroedinger/schrodecoder.c:1835-1848: level = depth-1 .. 0 */
void
oracle_iwt_inv (void *data, int stride, int width, int height, int is_s32,
    int filter, int depth)
{
  int level;
  for (level = depth - 1; level >= 0; level--)
    oracle_wavelet_inv (data, stride << level, width >> level, height >> level,
        is_s32, filter);
}
