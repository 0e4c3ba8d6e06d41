/*
 * oracle_dequant.c -- CPU restatement of the decoder's dequantisation (TEST INFRASTRUCTURE).
 * Follows schroedinger/schrodecoder.c:3395-3448 (what is applied to every codeblock),
 * schroedinger/schroorc.orc:1154-1168 / 2148-2162 (the Orc programs), schroedinger/schroparams.c:319-352
 * (where a subband lives in the frame) and schroedinger/schrodecoder.c:3559-3576 (codeblock bounds).
 */
#include "oracle.h"
#include <stddef.h>

static int16_t
dq16 (int16_t v, int factor, int offset)
{
  /* copyw, signw, absw, mullw, addw, shrsw 2, mullw: every step wraps at 16 bits */
  const int sign = v > 0 ? 1 : v < 0 ? -1 : 0;
  int16_t t = (int16_t) (v < 0 ? -v : v);
  t = (int16_t) (t * (int16_t) factor);
  t = (int16_t) (t + (int16_t) offset);
  t = (int16_t) (t >> 2);
  return (int16_t) (t * sign);
}

static int32_t
dq32 (int32_t v, int factor, int offset)
{
  const int32_t sign = v > 0 ? 1 : v < 0 ? -1 : 0;
  uint32_t t = v < 0 ? 0u - (uint32_t) v : (uint32_t) v;
  int32_t u;
  t = t * (uint32_t) factor;
  t = t + (uint32_t) offset;
  u = (int32_t) t >> 2;
  return (int32_t) ((uint32_t) u * (uint32_t) sign);
}

void
oracle_dequantise_plane (void *data, int stride, int width, int height, int is_s32,
    int transform_depth, const int *hcb, const int *vcb, const int32_t *quant)
{
  int index;
  for (index = 0; index <= 3 * transform_depth; index++) {
    /* band `index`: level i (0 = coarsest), orientation bits: 1 = right half, 2 = odd row of the pair */
    const int level = index == 0 ? 0 : (index - 1) / 3;
    const int orient = index == 0 ? 0 : (index - 1) % 3 + 1;
    const int shift = transform_depth - level;
    const int bw = width >> shift, bh = height >> shift;
    const ptrdiff_t bstride = (ptrdiff_t) stride << shift;
    unsigned char *base = (unsigned char *) data;
    const int nh = index == 0 ? hcb[0] : hcb[level + 1], nv = index == 0 ? vcb[0] : vcb[level + 1];
    int cx, cy, x, y;
    if (orient & 2) base += bstride >> 1;
    if (orient & 1) base += (size_t) bw * (is_s32 ? 4 : 2);
    for (cy = 0; cy < nv; cy++) {
      const int ymin = (bh * cy) / nv, ymax = (bh * (cy + 1)) / nv;
      for (cx = 0; cx < nh; cx++) {
        const int xmin = (bw * cx) / nh, xmax = (bw * (cx + 1)) / nh;
        const int factor = quant[0], offset = quant[1];
        quant += 2;
        for (y = ymin; y < ymax; y++) {
          unsigned char *row = base + (ptrdiff_t) y * bstride;
          for (x = xmin; x < xmax; x++) {
            if (is_s32) ((int32_t *) row)[x] = dq32 (((int32_t *) row)[x], factor, offset);
            else ((int16_t *) row)[x] = dq16 (((int16_t *) row)[x], factor, offset);
          }
        }
      }
    }
  }
}
