/*
 * oracle_lowdelay.c -- CPU restatement of the VC-2 / Dirac low-delay slice decoder.  TEST
 * INFRASTRUCTURE (oracle.h).
 *
 * Follows schro_decoder_decode_lowdelay_transform_data (schroedinger/schrolowdelay.c:745-761) and what
 * it dispatches to: per slice, a 7-bit quantiser base index, the length of the luma part, then the
 * luma coefficients of the slice's codeblock of every subband and the interleaved chroma coefficients
 * (:101-178, :180-257), each an interleaved exp-Golomb signed integer (schroedinger/schrounpack.c:209-246;
 * past its end a stream reads as 1 bits, :96-103) dequantised with the subband's quantiser
 * (schroedinger/schroutils.c:179-189, or the 16-bit Orc program of the "fast" path,
 * schroedinger/schroorc.orc:1204-1217); finally DC prediction of the LL band
 * (schroedinger/schrodecoder.c:3219-3277).
 *
 * The coefficient plane is the in-place subband layout (oracle_dequantise_plane's): level l's bands at
 * stride << l.  Bits are read one at a time here; the reference's table-driven reader is equivalent.
 */
#include <limits.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct { const uint8_t *data; long pos, end; } Bits;    /* bit positions in the picture's buffer; past `end`: 1 bits */

static int
bit (Bits *b)
{
  int v = 1;
  if (b->pos < b->end) v = (b->data[b->pos >> 3] >> (7 - (b->pos & 7))) & 1;
  b->pos++;
  return v;
}

static unsigned
bits (Bits *b, int n)
{
  unsigned v = 0;
  while (n-- > 0) v = (v << 1) | (unsigned) bit (b);
  return v;
}

static int
sint (Bits *b)
{
  unsigned count = 0, value = 0;
  int v;
  while (!bit (b)) { count++; value = (value << 1) | (unsigned) bit (b); }
  v = (int) ((1u << (count & 31)) - 1u + value);
  if (v && bit (b)) v = -v;
  return v;
}

static int
ilog2up (unsigned x)
{
  int i;
  for (i = 0; i < 32; i++) { if (x == 0) return i; x >>= 1; }
  return 0;
}

static int
dequant (int q, int factor, int offset, int orc16)
{
  if (orc16) {
    /* signw / absw / mullw / addw / shrsw / mullw, every step wrapping at 16 bits, the factor and
     * offset + 2 held in int16 arrays (schrolowdelay.c:494-496) */
    /* the fast path's reader stores int16 (schrounpack.c:274-330): the value is truncated before the program runs */
    const int16_t q16 = (int16_t) q;
    int16_t s = (int16_t) ((q16 > 0) - (q16 < 0)), t = (int16_t) (q16 < 0 ? -q16 : q16);
    t = (int16_t) (t * (int16_t) factor);
    t = (int16_t) (t + (int16_t) (offset + 2));
    t = (int16_t) (t >> 2);
    return (int16_t) (t * s);
  }
  if (q == 0) return 0;
  if (q < 0) return -((-q * factor + offset + 2) >> 2);
  return (q * factor + offset + 2) >> 2;
}

/* band `index` (schro_subband_get_position order: 0 = LL, then HL, LH, HH of the coarsest level, ...)
 * of a width x height plane: first sample, stride and size in samples */
static void
band_geometry (int index, int depth, int width, int height, int stride, ptrdiff_t *first, int *bstride, int *bw, int *bh)
{
  const int level = index == 0 ? 0 : (index - 1) / 3;           /* 0 = coarsest */
  const int orient = index == 0 ? 0 : (index - 1) % 3 + 1;      /* 1 HL, 2 LH, 3 HH */
  const int shift = depth - level;                              /* bands of this level are 2^-shift of the plane */
  *bw = width >> shift;
  *bh = height >> shift;
  *bstride = stride << shift;
  *first = 0;
  if (orient & 2) *first += (ptrdiff_t) (*bstride >> 1);        /* odd rows of the level */
  if (orient & 1) *first += *bw;                                /* right half */
}

static void
put (void *plane, int is_s32, ptrdiff_t at, int v)
{
  if (is_s32) ((int32_t *) plane)[at] = v;
  else ((int16_t *) plane)[at] = (int16_t) v;
}

static void
dc_predict (void *plane, int is_s32, int stride, int w, int h)
{
  int i, j;
#define AT(x, y) (is_s32 ? ((int32_t *) plane)[(ptrdiff_t) (y) * stride + (x)] : (int) ((int16_t *) plane)[(ptrdiff_t) (y) * stride + (x)])
  for (j = 0; j < h; j++)
    for (i = 0; i < w; i++) {
      int pred;
      if (j == 0) { if (i == 0) continue; pred = AT (i - 1, 0); }
      else if (i == 0) pred = AT (0, j - 1);
      else {
        const int a = AT (i - 1, j) + AT (i, j - 1) + AT (i - 1, j - 1) + 1;
        /* schro_divide3 for s16 (schroutils.h:64), schro_divide (a, 3) for s32 (:63) */
        pred = is_s32 ? (a < 0 ? (a - 3 + 1) / 3 : a / 3) : ((a * 21845 + 10922) >> 16);
      }
      put (plane, is_s32, (ptrdiff_t) j * stride + i, AT (i, j) + pred);
    }
#undef AT
}

void
oracle_lowdelay_decode (const uint8_t *data, int data_bytes, int slice_bytes_num, int slice_bytes_denom, int n_horiz_slices,
    int n_vert_slices, int transform_depth, const int *quant_matrix, const uint32_t *table_quant,
    const uint32_t *table_offset, void **planes, const int *strides, const int *widths, const int *heights, int is_s32,
    int orc16)
{
  const int n_bytes = slice_bytes_num / slice_bytes_denom, remainder = slice_bytes_num % slice_bytes_denom;
  const int nbands = 1 + 3 * transform_depth;
  int sx, sy, offset = 0, accumulator = 0, c, i;
  for (sy = 0; sy < n_vert_slices; sy++)
    for (sx = 0; sx < n_horiz_slices; sx++) {
      int extra = 0, slice_bytes, base_index, y_length;
      Bits yb, uvb;
      accumulator += remainder;
      if (accumulator >= slice_bytes_denom) { extra = 1; accumulator -= slice_bytes_denom; }
      slice_bytes = n_bytes + extra;
      yb.data = data; yb.pos = 8L * offset; yb.end = 8L * (offset + slice_bytes);
      offset += slice_bytes;
      base_index = (int) bits (&yb, 7);
      /* the fast path takes the length field's width from n_bytes, the slow paths from this slice's size
       * (schrolowdelay.c:576 vs :122) -- the same number unless 8 * n_bytes is one below a power of two */
      y_length = (int) bits (&yb, ilog2up (8u * (unsigned) (orc16 ? n_bytes : slice_bytes)));
      uvb = yb;
      uvb.pos = yb.pos + y_length;
      /* schro_unpack_limit_bits_remaining (schrounpack.c:49-60) takes the declared length at its word: a
       * luma length beyond the slice makes the reference read on into the following slices' bytes.  Only
       * the end of the picture's buffer stops it here (the reference would read past its allocation). */
      yb.end = yb.pos + y_length;
      if (yb.end > 8L * data_bytes) yb.end = 8L * data_bytes;
      for (c = 0; c < 3; c += 2)          /* c == 0: luma, c == 2: both chroma planes interleaved */
        for (i = 0; i < nbands; i++) {
          int qi = base_index - quant_matrix[i], factor, qoffset, bw, bh, bstride, x, y, x0, x1, y0, y1;
          ptrdiff_t first;
          const int k = c ? 1 : 0;
          qi = qi < 0 ? 0 : (qi > 60 ? 60 : qi);
          factor = (int) table_quant[qi];
          qoffset = (int) table_offset[qi];
          band_geometry (i, transform_depth, widths[k], heights[k], strides[k], &first, &bstride, &bw, &bh);
          x0 = bw * sx / n_horiz_slices; x1 = bw * (sx + 1) / n_horiz_slices;        /* schro_frame_data_get_codeblock */
          y0 = bh * sy / n_vert_slices; y1 = bh * (sy + 1) / n_vert_slices;
          for (y = y0; y < y1; y++)
            for (x = x0; x < x1; x++) {
              const ptrdiff_t at = first + (ptrdiff_t) y * bstride + x;
              if (!c) put (planes[0], is_s32, at, dequant (sint (&yb), factor, qoffset, orc16));
              else {
                put (planes[1], is_s32, at, dequant (sint (&uvb), factor, qoffset, orc16));
                put (planes[2], is_s32, at, dequant (sint (&uvb), factor, qoffset, orc16));
              }
            }
        }
    }
  for (c = 0; c < 3; c++)
    dc_predict (planes[c], is_s32, strides[c] << transform_depth, widths[c] >> transform_depth, heights[c] >> transform_depth);
}
