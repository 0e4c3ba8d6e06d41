/*
 * oracle_hbm.c -- CPU restatement of the SAD primitives and of one level of hierarchical
 * block matching.  TEST INFRASTRUCTURE (oracle.h).
 *
 * Follows schro_metric_absdiff_u8 (schroedinger/schrometric.c:10-29),
 * schro_metric_block_sad_slow (:332-375), schro_metric_scan_setup / _do_scan / _get_min
 * (:31-214) and schro_hierarchical_bm_scan_hint (schroedinger/schrohierbm.c:174-383);
 * SURVEY.md Appendix C is the step list.
 */
#include <limits.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline int clampi (int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int mini (int a, int b) { return a < b ? a : b; }
static inline int maxi (int a, int b) { return a > b ? a : b; }

uint32_t
oracle_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int width, int height)
{
  uint32_t s = 0;
  int x, y;
  for (y = 0; y < height; y++)
    for (x = 0; x < width; x++)
      s += (uint32_t) abs ((int) a[(ptrdiff_t) a_stride * y + x] - (int) b[(ptrdiff_t) b_stride * y + x]);
  return s;
}

static int
block_is_valid (const OraclePyrLevel *f, int x, int y, int sx, int sy)
{
  return !(x < -f->ext || y < -f->ext || x + sx > f->width + f->ext || y + sy > f->height + f->ext);
}

static int comp_w (const OraclePyrLevel *f, int k) { return k ? (f->width + (1 << f->h_shift) - 1) >> f->h_shift : f->width; }
static int comp_h (const OraclePyrLevel *f, int k) { return k ? (f->height + (1 << f->v_shift) - 1) >> f->v_shift : f->height; }

/* Y+U+V SAD used to rank candidates (schrometric.c:332-375): block sizes are clipped by
 * what is left of the SOURCE component only; INT_MAX when a block leaves frame+extension */
static int
block_sad3 (const OraclePyrLevel *src, const OraclePyrLevel *ref, int bw, int bh, int x, int y,
    int dx, int dy)
{
  int k, i, j, metric = 0;
  if (!block_is_valid (src, x, y, bw, bh)) return INT_MAX;
  if (!block_is_valid (ref, x + dx, y + dy, bw, bh)) return INT_MAX;
  for (k = 0; k < 3; k++) {
    int hs = k ? src->h_shift : 0, vs = k ? src->v_shift : 0;
    int sx = x >> hs, sy = y >> vs;
    int rx = (x + dx) >> hs, ry = (y + dy) >> vs;
    int w = mini (maxi (0, comp_w (src, k) - sx), bw >> hs);
    int h = mini (maxi (0, comp_h (src, k) - sy), bh >> vs);
    for (j = 0; j < h; j++)
      for (i = 0; i < w; i++)
        metric += abs ((int) src->data[k][(ptrdiff_t) src->stride[k] * (sy + j) + sx + i]
            - (int) ref->data[k][(ptrdiff_t) ref->stride[k] * (ry + j) + rx + i]);
  }
  return metric;
}

void
oracle_hbm_scan_hint (const OraclePyrLevel *src, const OraclePyrLevel *ref, int xbsep, int ybsep,
    int x_num_blocks, int y_num_blocks, int ref_index, int shift, int h_range, int use_chroma,
    const OracleMotionVector *parent, OracleMotionVector *mf)
{
  const int skip = 1 << shift;
  const int split = shift > 1 ? 0 : (shift == 1 ? 1 : 2);
  const uint32_t flags0 = (uint32_t) (ref_index + 1) | ((uint32_t) split << 3);
  const int hint_mask = ~((1 << (shift + 1)) - 1);
  OracleMotionVector zero_mv;
  int i, j, n;

  /* schro_motion_field_new + _set (schromotionest.c:395-432) */
  n = x_num_blocks * y_num_blocks;
  memset (mf, 0, sizeof (*mf) * (size_t) n);
  for (i = 0; i < n; i++) mf[i].flags = flags0;
  memset (&zero_mv, 0, sizeof (zero_mv));
  zero_mv.flags = flags0;

  for (j = 0; j < y_num_blocks; j += skip) {
    for (i = 0; i < x_num_blocks; i += skip) {
      const OracleMotionVector *cand[9], *uniq[9];
      const int x0 = (i * xbsep) >> shift, y0 = (j * ybsep) >> shift;
      int nc = 0, nu = 0, m, bw0, bh0, min_m = -1, min_metric = INT_MAX;
      int dx, dy, xmin, xmax, ymin, ymax, scan_w, scan_h, a, b;
      uint32_t best, best_c = 0, best_t = 0;
      OracleMotionVector *out;

      if (!(src->width > x0) || !(src->height > y0)) continue;
      bw0 = mini (src->width - x0, xbsep);
      bh0 = mini (src->height - y0, ybsep);

      /* candidates: zero, 5 parents (star), 3 already-scanned neighbours (schrohierbm.c:259-294) */
      cand[nc++] = &zero_mv;
      if (parent) {
        static const int off[5][2] = { {0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1} };
        int l = i & hint_mask, k = j & hint_mask;
        for (m = 0; m < 5; m++) {
          int ll = l + off[m][0] * skip * 2, kk = k + off[m][1] * skip * 2;
          if (ll >= 0 && ll < x_num_blocks && kk >= 0 && kk < y_num_blocks)
            cand[nc++] = &parent[kk * x_num_blocks + ll];
        }
      }
      if (i > 0) cand[nc++] = &mf[j * x_num_blocks + i - skip];
      if (j > 0) cand[nc++] = &mf[(j - skip) * x_num_blocks + i];
      if (i > 0 && j > 0) cand[nc++] = &mf[(j - skip) * x_num_blocks + i - skip];

      /* de-duplicate, keeping the LAST occurrence of each vector (:298-321) */
      for (m = 0; m < nc; m++) {
        int s, dup = 0;
        for (s = m + 1; s < nc && !dup; s++)
          dup = cand[m]->v[ref_index] == cand[s]->v[ref_index]
              && cand[m]->v[2 + ref_index] == cand[s]->v[2 + ref_index];
        if (!dup) uniq[nu++] = cand[m];
      }

      /* rank by 3-component SAD, first strict minimum wins (:323-346) */
      for (m = 0; m < nu; m++) {
        int metric;
        dx = uniq[m]->v[ref_index] >> shift;
        dx = clampi (dx + x0, -bw0, ref->width) - x0;
        dy = uniq[m]->v[2 + ref_index] >> shift;
        dy = clampi (dy + y0, -bh0, ref->height) - y0;
        metric = block_sad3 (src, ref, xbsep, ybsep, x0, y0, dx, dy);
        if (metric < min_metric) { min_metric = metric; min_m = m; }
      }
      if (min_m < 0) min_m = 0;   /* the reference asserts here; keep going for robustness */

      /* seed and scan window (:349-364, schrometric.c:174-214) */
      dx = uniq[min_m]->v[ref_index] >> shift;
      dy = uniq[min_m]->v[2 + ref_index] >> shift;
      dx = maxi (-bw0 - x0, mini (ref->width - x0, dx));
      dy = maxi (-bh0 - y0, mini (ref->height - y0, dy));
      xmin = maxi (maxi (-bw0, x0 + dx - h_range), -src->ext);
      ymin = maxi (maxi (-bh0, y0 + dy - h_range), -src->ext);
      xmax = mini (mini (src->width, x0 + dx + h_range), src->width - bw0 + src->ext);
      ymax = mini (mini (src->height, y0 + dy + h_range), src->height - bh0 + src->ext);
      scan_w = xmax - xmin + 1;
      scan_h = ymax - ymin + 1;

      /* full search, seed wins ties, then first strict minimum in x-outer / y-inner order
       * (schrometric.c:31-171).  Chroma term (4:2:0 duplication, :73-115) only if use_chroma. */
      {
        uint32_t seed_l, seed_c = 0;
        int bdx = dx, bdy = dy;
#define LUMA(A, B) oracle_sad_u8 (src->data[0] + (ptrdiff_t) src->stride[0] * y0 + x0, src->stride[0], \
            ref->data[0] + (ptrdiff_t) ref->stride[0] * (ymin + (B)) + xmin + (A), ref->stride[0], bw0, bh0)
        seed_l = LUMA (dx + x0 - xmin, dy + y0 - ymin);
        best = seed_l;
        if (use_chroma) {
          /* chroma_metrics[i*scan_h+j] as filled by do_scan */
          int skip_h = 1 << src->h_shift, skip_v = 1 << src->v_shift;
          int cx = x0 / skip_h, cy = y0 / skip_v, crx = xmin / skip_h, cry = ymin / skip_v;
          int cbw = bw0 / skip_h, cbh = bh0 / skip_v;
          uint32_t *cm = calloc ((size_t) scan_w * scan_h + 4 * (scan_w + scan_h) + 16, sizeof (uint32_t));
          uint32_t *tmp = calloc ((size_t) (scan_w + 2) * (scan_h + 2) * 2 + 16, sizeof (uint32_t));
          int sw = scan_w / skip_h + scan_w % skip_h, sh = scan_h / skip_v + scan_h % skip_v;
          int k;
          for (k = 1; k < 3; k++) {
            int ii, jj;
            for (ii = 0; ii < sw; ii++) {
              for (jj = 0; jj < sh; jj++) {
                uint32_t v = oracle_sad_u8 (src->data[k] + (ptrdiff_t) src->stride[k] * cy + cx, src->stride[k],
                    ref->data[k] + (ptrdiff_t) ref->stride[k] * (cry + jj) + crx + ii, ref->stride[k], cbw, cbh);
                tmp[ii * 2 * scan_h + jj * 2] = v;
                if (skip_v > 1) tmp[ii * 2 * scan_h + 1 + jj * 2] = v;
              }
              if (skip_h > 1)
                for (jj = 0; jj < scan_h; jj++)
                  tmp[(ii * 2 + 1) * scan_h + jj] = tmp[ii * 2 * scan_h + jj];
            }
            for (jj = 0; jj < scan_h; jj++)
              for (ii = 0; ii < scan_w; ii++)
                cm[ii * scan_h + jj] += tmp[ii * scan_h + jj];
          }
          seed_c = cm[(dy + y0 - ymin) + (dx + x0 - xmin) * scan_h];
          best_c = seed_c;
          best_t = seed_l + seed_c;
          for (a = 0; a < scan_w; a++)
            for (b = 0; b < scan_h; b++) {
              uint32_t l = LUMA (a, b), c = cm[a * scan_h + b];
              if (l + c < best_t) { best_t = l + c; best = l; best_c = c; bdx = xmin + a - x0; bdy = ymin + b - y0; }
            }
          free (cm);
          free (tmp);
        } else {
          for (a = 0; a < scan_w; a++)
            for (b = 0; b < scan_h; b++) {
              uint32_t l = LUMA (a, b);
              if (l < best) { best = l; bdx = xmin + a - x0; bdy = ymin + b - y0; }
            }
        }
#undef LUMA
        out = &mf[j * x_num_blocks + i];
        out->metric = best;
        out->chroma_metric = best_c;
        out->v[ref_index] = (int16_t) (bdx << shift);
        out->v[2 + ref_index] = (int16_t) (bdy << shift);
        out->flags = flags0;
      }
    }
  }
}

/* ---- the metric-scan entry points on their own (schrometric.c:31-214, 380-414) ------------- */

/* schro_metric_scan_setup (:174-214): the scan window of a block at (x, y) around (dx, dy) */
void
oracle_metric_scan_setup (const OraclePyrLevel *f, int x, int y, int bw, int bh, int dx, int dy, int dist,
    int *ref_x, int *ref_y, int *scan_w, int *scan_h)
{
  int xmin = maxi (maxi (-bw, x + dx - dist), -f->ext);
  int ymin = maxi (maxi (-bh, y + dy - dist), -f->ext);
  int xmax = mini (mini (f->width, x + dx + dist), f->width - bw + f->ext);
  int ymax = mini (mini (f->height, y + dy + dist), f->height - bh + f->ext);
  *ref_x = xmin;
  *ref_y = ymin;
  *scan_w = xmax - xmin + 1;
  *scan_h = ymax - ymin + 1;
}

/* schro_metric_scan_do_scan (:31-116): metrics[i * scan_h + j] = luma SAD at (ref_x + i, ref_y + j);
 * chroma_metrics = 0, or with use_chroma the sum of both chroma SADs computed on the sub-sampled grid
 * and duplicated onto the luma grid exactly as the reference does (4:2:0 duplication, :73-115) */
void
oracle_metric_scan_do_scan (const OraclePyrLevel *src, const OraclePyrLevel *ref, int x, int y, int bw, int bh,
    int ref_x, int ref_y, int scan_w, int scan_h, int use_chroma, uint32_t *metrics, uint32_t *chroma_metrics)
{
  int i, j, k;
  for (i = 0; i < scan_w; i++)
    for (j = 0; j < scan_h; j++)
      metrics[i * scan_h + j] = oracle_sad_u8 (src->data[0] + (ptrdiff_t) src->stride[0] * y + x, src->stride[0],
          ref->data[0] + (ptrdiff_t) ref->stride[0] * (ref_y + j) + ref_x + i, ref->stride[0], bw, bh);
  memset (chroma_metrics, 0, sizeof (uint32_t) * 42 * 42);
  if (!use_chroma) return;
  {
    const int skip_h = 1 << src->h_shift, skip_v = 1 << src->v_shift;
    const int cx = x / skip_h, cy = y / skip_v, crx = ref_x / skip_h, cry = ref_y / skip_v;
    const int cbw = bw / skip_h, cbh = bh / skip_v;
    const int sw = scan_w / skip_h + scan_w % skip_h, sh = scan_h / skip_v + scan_h % skip_v;
    uint32_t *tmp = calloc (42 * 42 * 4 + 64, sizeof (uint32_t));
    for (k = 1; k < 3; k++) {
      for (i = 0; i < sw; i++) {
        for (j = 0; j < sh; j++) {
          const uint32_t v = oracle_sad_u8 (src->data[k] + (ptrdiff_t) src->stride[k] * cy + cx, src->stride[k],
              ref->data[k] + (ptrdiff_t) ref->stride[k] * (cry + j) + crx + i, ref->stride[k], cbw, cbh);
          tmp[i * 2 * scan_h + j * 2] = v;
          if (skip_v > 1) tmp[i * 2 * scan_h + 1 + j * 2] = v;
        }
        if (skip_h > 1)
          for (j = 0; j < scan_h; j++) tmp[(i * 2 + 1) * scan_h + j] = tmp[i * 2 * scan_h + j];
      }
      for (j = 0; j < scan_h; j++)
        for (i = 0; i < scan_w; i++) chroma_metrics[i * scan_h + j] += tmp[i * scan_h + j];
    }
    free (tmp);
  }
}

/* schro_metric_scan_get_min (:121-171): the seed (gravity) wins ties, else the first strict minimum
 * in x-outer / y-inner order; returns the luma metric */
uint32_t
oracle_metric_scan_get_min (const uint32_t *metrics, const uint32_t *chroma_metrics, int x, int y, int ref_x, int ref_y,
    int scan_w, int scan_h, int gravity_x, int gravity_y, int use_chroma, int *dx, int *dy, uint32_t *chroma_error)
{
  int i = gravity_x + x - ref_x, j = gravity_y + y - ref_y;
  uint32_t min_metric = metrics[j + i * scan_h], min_chroma = 0, min_total = 0;
  if (use_chroma) {
    min_chroma = chroma_metrics[j + i * scan_h];
    min_total = min_metric + min_chroma;
  }
  for (i = 0; i < scan_w; i++)
    for (j = 0; j < scan_h; j++) {
      const uint32_t m = metrics[i * scan_h + j], c = chroma_metrics[i * scan_h + j];
      if (use_chroma ? (m + c < min_total) : (m < min_metric)) {
        min_total = m + c;
        min_metric = m;
        min_chroma = c;
        *dx = ref_x + i - x;
        *dy = ref_y + j - y;
      }
    }
  *chroma_error = min_chroma;
  return min_metric;
}

/* schro_metric_fast_block (:410-414) = schro_metric_block_sad_slow (:332-375) */
int
oracle_metric_fast_block (const OraclePyrLevel *src, const OraclePyrLevel *ref, int bw, int bh, int x, int y,
    int dx, int dy)
{
  return block_sad3 (src, ref, bw, bh, x, y, dx, dy);
}
