/*
 * oracle_rough.c -- CPU restatement of the rough (bigblock) motion search.  TEST INFRASTRUCTURE
 * (oracle.h).
 *
 * Follows schro_rough_me_heirarchical_scan_nohint (schroedinger/schroroughmotion.c:62-143: a
 * dependency-free full search of every block of one pyramid level) and
 * schro_rough_me_heirarchical_scan_hint (:145-300: candidates = zero, the four nearest parents,
 * left / up / up-left of the same level, ranked by luma SAD, then a small scan around the winner),
 * both on top of the metric-scan functions restated in oracle_hbm.c.
 *
 * One point where the reference's result is not defined: a block that lies entirely outside its
 * level's frame (x >= width or y >= height; happens when x_num_blocks * xbsep over-covers the
 * picture).  At a hint level all its candidates are skipped, the zero vector becomes the seed, and
 * schro_metric_scan_get_min (schroedinger/schrometric.c:137-139) starts from a metrics[] entry
 * outside the scan window -- stale stack memory of the previous block's scan.  Every SAD of an
 * empty block is 0, so the outcome only depends on whether that stale word is 0: if it is, the
 * block keeps the zero vector, else it takes the window's first position; the metric is 0 either way.
 * Here (and in the CUDA path) such a block keeps the zero vector -- what the compiled reference
 * produced for three out of four such blocks in the test pictures (they occurred at level 1 only,
 * whose outside blocks feed nothing inside the rough search).  tests/test_oracle_rough.py compares
 * the blocks that overlap their frame on such sizes, and every block on sizes the grid covers exactly.
 */
#include <limits.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline int mini (int a, int b) { return a < b ? a : b; }
static inline int maxi (int a, int b) { return a > b ? a : b; }

static void
field_reset (OracleMotionVector *mf, int n)
{
  int i;
  /* schro_motion_field_set (mf, 0, 1) (schromotionest.c:416-432): split 0, pred_mode 1 */
  memset (mf, 0, sizeof (*mf) * (size_t) n);
  for (i = 0; i < n; i++) mf[i].flags = 1;
}

/* the scan both functions end with: window around (x + dx, y + dy), SADs, first strict minimum
 * starting from the seed position (seed_inside == 0: from +infinity, see the header) */
static void
scan_block (const OraclePyrLevel *src, const OraclePyrLevel *ref, int x, int y, int bw, int bh, int dx, int dy,
    int distance, int seed_inside, int shift, int ref_index, int nohint, OracleMotionVector *mv)
{
  int ref_x, ref_y, scan_w, scan_h, a, b, best_a, best_b;
  uint32_t best;
  oracle_metric_scan_setup (src, x, y, bw, bh, dx, dy, distance, &ref_x, &ref_y, &scan_w, &scan_h);
  if (scan_w <= 0 || scan_h <= 0) {
    /* the nohint function clears index 0, the hint function index `ref` (:105-109, :285-289) */
    mv->v[nohint ? 0 : ref_index] = 0;
    mv->v[2 + (nohint ? 0 : ref_index)] = 0;
    mv->metric = (uint32_t) INT_MAX;
    return;
  }
  if (!nohint && !seed_inside) {
    /* empty block at a hint level: see the header */
    mv->metric = 0;
    mv->v[ref_index] = (int16_t) (dx << shift);
    mv->v[2 + ref_index] = (int16_t) (dy << shift);
    return;
  }
  if (seed_inside) {
    best_a = x + dx - ref_x;
    best_b = y + dy - ref_y;
    best = oracle_sad_u8 (src->data[0] + (ptrdiff_t) src->stride[0] * y + x, src->stride[0],
        ref->data[0] + (ptrdiff_t) ref->stride[0] * (ref_y + best_b) + ref_x + best_a, ref->stride[0], bw, bh);
  } else {
    best_a = best_b = 0;
    best = 0xffffffffu;
  }
  for (a = 0; a < scan_w; a++)
    for (b = 0; b < scan_h; b++) {
      const uint32_t m = oracle_sad_u8 (src->data[0] + (ptrdiff_t) src->stride[0] * y + x, src->stride[0],
          ref->data[0] + (ptrdiff_t) ref->stride[0] * (ref_y + b) + ref_x + a, ref->stride[0], bw, bh);
      if (m < best) { best = m; best_a = a; best_b = b; }
    }
  mv->metric = best;
  mv->v[ref_index] = (int16_t) ((ref_x + best_a - x) << shift);
  mv->v[2 + ref_index] = (int16_t) ((ref_y + best_b - y) << shift);
}

void
oracle_rough_scan_nohint (const OraclePyrLevel *src, const OraclePyrLevel *ref, int xbsep, int ybsep,
    int x_num_blocks, int y_num_blocks, int ref_index, int shift, int distance, OracleMotionVector *mf)
{
  const int skip = 1 << shift;
  int i, j;
  field_reset (mf, x_num_blocks * y_num_blocks);
  for (j = 0; j < y_num_blocks; j += skip)
    for (i = 0; i < x_num_blocks; i += skip) {
      const int x = (i >> shift) * xbsep, y = (j >> shift) * ybsep;
      const int bw = mini (src->width - x, xbsep), bh = mini (src->height - y, ybsep);
      /* gravity = the window's first position (:98-102): the search is a plain first strict minimum */
      scan_block (src, ref, x, y, bw, bh, 0, 0, distance, 0, shift, ref_index, 1, &mf[j * x_num_blocks + i]);
    }
}

void
oracle_rough_scan_hint (const OraclePyrLevel *src, const OraclePyrLevel *ref, int xbsep, int ybsep,
    int x_num_blocks, int y_num_blocks, int ref_index, int shift, int distance,
    const OracleMotionVector *parent, OracleMotionVector *mf)
{
  const int skip = 1 << shift;
  const int hint_mask = ~((1 << (shift + 1)) - 1);
  int i, j;
  field_reset (mf, x_num_blocks * y_num_blocks);
  for (j = 0; j < y_num_blocks; j += skip)
    for (i = 0; i < x_num_blocks; i += skip) {
      int cdx[8], cdy[8], n = 0, m, best_m = 0, best_metric = INT_MAX;
      const int x = (i * xbsep) >> shift, y = (j * ybsep) >> shift;
      const int ow = maxi (0, src->width - x), oh = maxi (0, src->height - y);     /* schro_frame_get_subdata */
      const int w = mini (xbsep, ow), h = mini (ybsep, oh);
      cdx[n] = 0; cdy[n] = 0; n++;
      for (m = 0; m < 4; m++) {
        const int l = (i + skip * (-1 + 2 * (m & 1))) & hint_mask;
        const int k = (j + skip * (-1 + (m & 2))) & hint_mask;
        if (l >= 0 && l < x_num_blocks && k >= 0 && k < y_num_blocks) {
          cdx[n] = parent[k * x_num_blocks + l].v[ref_index];
          cdy[n] = parent[k * x_num_blocks + l].v[2 + ref_index];
          n++;
        }
      }
      if (i > 0) { cdx[n] = mf[j * x_num_blocks + i - skip].v[ref_index]; cdy[n] = mf[j * x_num_blocks + i - skip].v[2 + ref_index]; n++; }
      if (j > 0) { cdx[n] = mf[(j - skip) * x_num_blocks + i].v[ref_index]; cdy[n] = mf[(j - skip) * x_num_blocks + i].v[2 + ref_index]; n++; }
      if (i > 0 && j > 0) {
        cdx[n] = mf[(j - skip) * x_num_blocks + i - skip].v[ref_index];
        cdy[n] = mf[(j - skip) * x_num_blocks + i - skip].v[2 + ref_index];
        n++;
      }
      for (m = 0; m < n; m++) {
        const int rx = (i * xbsep + cdx[m]) >> shift, ry = (j * ybsep + cdy[m]) >> shift;
        int metric;
        if (rx < 0 || ry < 0) continue;
        if (w == 0 || h == 0) continue;
        if (maxi (0, ref->width - rx) < w || maxi (0, ref->height - ry) < h) continue;
        metric = (int) oracle_sad_u8 (src->data[0] + (ptrdiff_t) src->stride[0] * y + x, src->stride[0],
            ref->data[0] + (ptrdiff_t) ref->stride[0] * ry + rx, ref->stride[0], w, h);
        if (metric < best_metric) { best_metric = metric; best_m = m; }
      }
      {
        const int dx = cdx[best_m] >> shift, dy = cdy[best_m] >> shift;
        const int bw = mini (src->width - x, xbsep), bh = mini (src->height - y, ybsep);
        scan_block (src, ref, x, y, bw, bh, dx, dy, distance, bw > 0 && bh > 0, shift, ref_index, 0,
            &mf[j * x_num_blocks + i]);
      }
    }
}
