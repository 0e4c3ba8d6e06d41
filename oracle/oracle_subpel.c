/*
 * oracle_subpel.c -- CPU restatement of the sub-pel refinement of a motion field.  TEST
 * INFRASTRUCTURE (oracle.h).
 *
 * Follows schro_encoder_motion_predict_subpel_deep (schroedinger/schromotionest.c:246-355): for
 * mvprec = 1 .. mv_precision, every block's vector is doubled and the eight sub-pel neighbours of
 * it are probed -- block fetch at that precision (schro_upsampled_frame_get_block_fast_precN,
 * schroedinger/schroframe.c:2287-2482), luma SAD against the source block -- the score being
 * entropy (schro_pack_estimate_sint of the difference to the median prediction from the ALREADY
 * REFINED left / up / up-left vectors, schroedinger/schromotion.c:259-312, schropack.c:204-226)
 * + lambda * SAD in double precision.
 *
 * Structured as the CUDA path is: the probe SADs of a block depend on its own vector only and are
 * computed first; the decisions then run in raster order.
 */
#include <limits.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static int
bits_uint (int value)
{
  unsigned int x = (unsigned int) (value + 1);
  int n = 0;
  while (x) { n++; x >>= 1; }
  return n + n - 1;
}

int oracle_bits_sint (int value);
int oracle_subpel_sample (const uint8_t *ref, int rstride, int prec, int x, int y, int a, int b);
#define bits_sint oracle_bits_sint
#define subpel_sample oracle_subpel_sample

/* schro_pack_estimate_sint (schroedinger/schropack.c:204-226) */
int
oracle_bits_sint (int value)
{
  int n;
  if (value < 0) value = -value;
  n = bits_uint (value);
  return value ? n + 1 : n;
}

static int
med3 (int a, int b, int c)
{
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return c < lo ? lo : (c > hi ? hi : c);
}

/* half-pel sample (u, v): phase ((v & 1) << 1 | (u & 1)) at (u >> 1, v >> 1) (schroframe.c:2186-2200) */
static int
half_sample (const uint8_t *ref, int rstride, int u, int v, int a, int b)
{
  const int ph = ((v & 1) << 1) | (u & 1);
  return ref[(ptrdiff_t) ph * (rstride >> 2) + (ptrdiff_t) ((v >> 1) + b) * rstride + (u >> 1) + a];
}

/* pixel (a, b) of the block fetched at (x, y) in units of 2^-prec pixels, prec 0..3
 * (schro_upsampled_frame_get_block_fast_precN, schroedinger/schroframe.c:2458-2482) */
int
oracle_subpel_sample (const uint8_t *ref, int rstride, int prec, int x, int y, int a, int b)
{
  int hx, hy, rx, ry;
  if (prec == 0) return ref[(ptrdiff_t) (y + b) * rstride + x + a];
  if (prec == 1) return half_sample (ref, rstride, x, y, a, b);
  if (prec == 2) { x <<= 1; y <<= 1; }
  hx = x >> 2; hy = y >> 2; rx = x & 3; ry = y & 3;
  if (!rx && !ry) return half_sample (ref, rstride, hx, hy, a, b);
  if (rx == 2 && !ry) return (half_sample (ref, rstride, hx, hy, a, b) + half_sample (ref, rstride, hx + 1, hy, a, b) + 1) >> 1;
  if (!rx && ry == 2) return (half_sample (ref, rstride, hx, hy, a, b) + half_sample (ref, rstride, hx, hy + 1, a, b) + 1) >> 1;
  return ((4 - ry) * (4 - rx) * half_sample (ref, rstride, hx, hy, a, b)
      + (4 - ry) * rx * half_sample (ref, rstride, hx + 1, hy, a, b)
      + ry * (4 - rx) * half_sample (ref, rstride, hx, hy + 1, a, b)
      + ry * rx * half_sample (ref, rstride, hx + 1, hy + 1, a, b) + 8) >> 4;
}

static const int probe_dx[8] = { -1, 0, 1, -1, 1, -1, 0, 1 };
static const int probe_dy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 };

void
oracle_subpel_refine (const uint8_t *orig, int orig_stride, int width, int height, int orig_ext,
    const uint8_t *upref, int rstride, int xblen, int yblen, int x_num_blocks, int y_num_blocks,
    int mv_precision, int ref_index, double lambda, OracleMotionVector *mf)
{
  const int n = x_num_blocks * y_num_blocks;
  int *err = malloc (sizeof (int) * 8 * (size_t) n);
  unsigned char *mask = malloc ((size_t) n);
  int mvprec, i, j, k;
  for (mvprec = 1; mvprec <= mv_precision; mvprec++) {
    const int x_min = -orig_ext, y_min = -orig_ext;
    const int x_max = (width << mvprec) + orig_ext, y_max = (height << mvprec) + orig_ext;
    /* 1: probe SADs around the doubled vector of every block that overlaps the picture */
    for (j = 0; j < y_num_blocks; j++)
      for (i = 0; i < x_num_blocks; i++) {
        const OracleMotionVector *mv = &mf[j * x_num_blocks + i];
        const int w = xblen < width - i * xblen ? xblen : width - i * xblen;
        const int h = yblen < height - j * yblen ? yblen : height - j * yblen;
        int x, y;
        mask[j * x_num_blocks + i] = 0;
        if (w <= 0 || h <= 0) continue;
        x = i * (xblen << mvprec) + (int16_t) (mv->v[ref_index] << 1);
        y = j * (yblen << mvprec) + (int16_t) (mv->v[2 + ref_index] << 1);
        for (k = 0; k < 8; k++) {
          const int px = x + probe_dx[k], py = y + probe_dy[k];
          int a, b, e = 0;
          if (!(x_min < px) || !(x_max > px + xblen - 1) || !(y_min < py) || !(y_max > py + yblen - 1)) continue;
          for (b = 0; b < h; b++)
            for (a = 0; a < w; a++)
              e += abs ((int) orig[(ptrdiff_t) (j * yblen + b) * orig_stride + i * xblen + a]
                  - subpel_sample (upref, rstride, mvprec, px, py, a, b));
          err[(j * x_num_blocks + i) * 8 + k] = e;
          mask[j * x_num_blocks + i] |= (unsigned char) (1 << k);
        }
      }
    /* 2: decisions in raster order; the prediction sees the refined neighbours */
    for (j = 0; j < y_num_blocks; j++)
      for (i = 0; i < x_num_blocks; i++) {
        OracleMotionVector *mv = &mf[j * x_num_blocks + i];
        int pred_x = 0, pred_y = 0, m = -1, min_error = INT_MAX;
        double min_score;
        if (!(width > i * xblen) || !(height > j * yblen)) continue;
        mv->v[ref_index] = (int16_t) (mv->v[ref_index] << 1);
        mv->v[2 + ref_index] = (int16_t) (mv->v[2 + ref_index] << 1);
        if (i > 0 && j > 0) {
          const OracleMotionVector *l = mv - 1, *u = mv - x_num_blocks, *ul = mv - x_num_blocks - 1;
          pred_x = med3 (l->v[ref_index], u->v[ref_index], ul->v[ref_index]);
          pred_y = med3 (l->v[2 + ref_index], u->v[2 + ref_index], ul->v[2 + ref_index]);
        } else if (i > 0 || j > 0) {
          const OracleMotionVector *o = i > 0 ? mv - 1 : mv - x_num_blocks;
          pred_x = o->v[ref_index];
          pred_y = o->v[2 + ref_index];
        }
        min_score = (double) (bits_sint (mv->v[ref_index] - pred_x) + bits_sint (mv->v[2 + ref_index] - pred_y))
            + lambda * (double) mv->metric;
        for (k = 0; k < 8; k++) {
          double score;
          if (!((mask[j * x_num_blocks + i] >> k) & 1)) continue;
          score = (double) (bits_sint (mv->v[ref_index] + probe_dx[k] - pred_x)
              + bits_sint (mv->v[2 + ref_index] + probe_dy[k] - pred_y))
              + lambda * (double) err[(j * x_num_blocks + i) * 8 + k];
          if (min_score > score) { min_score = score; min_error = err[(j * x_num_blocks + i) * 8 + k]; m = k; }
        }
        if (m >= 0) {
          mv->v[ref_index] = (int16_t) (mv->v[ref_index] + probe_dx[m]);
          mv->v[2 + ref_index] = (int16_t) (mv->v[2 + ref_index] + probe_dy[m]);
          mv->metric = (uint32_t) min_error;
        }
      }
  }
  free (err);
  free (mask);
}
