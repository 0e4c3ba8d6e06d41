/*
 * oracle_split2.c -- CPU restatement of the split-2 pass of the encoder's mode decision.  TEST
 * INFRASTRUCTURE (oracle.h).
 *
 * Follows schro_do_split2 (schroedinger/schromotionest.c:1601-1802) + schro_motion_copy_to
 * (:1511-1523) applied to every superblock, the first step schro_mode_decision (:2587-2685) takes
 * for each of them.  Per block inside the picture the candidates are
 *   - each reference's sub-pel vector: luma SAD from the field + chroma SADs at the halved vector
 *     (schro_get_split2_metric, :1527-1594), cost = entropy of the vector against its prediction
 *     (schro_motion_block_estimate_entropy, :1243-1282; schro_motion_vector_prediction,
 *     schroedinger/schromotion.c:315-368; schro_pack_estimate_sint) + lambda * error;
 *   - both vectors together (schro_metric_get_biref, schroedinger/schrometric.c:273-304), if the
 *     luma blocks lie inside the extended frame;
 *   - a DC block (schro_block_average, :481-516) when the best error so far exceeds four per sample.
 * The prediction of a block reads the DECIDED left / up / up-left blocks, so decisions run in
 * raster order (equivalent to the reference's superblock order: those three neighbours are always
 * decided first in both).
 *
 * Three properties of the reference that this file reproduces because the results depend on them:
 *  1. The winner of a single-reference candidate records best_error = the LUMA metric only
 *     (:1683), although the score used luma + chroma.
 *  2. With mv_precision >= 2 the three component fetches of the bi-reference candidate land in one
 *     scratch block per reference (fd[ref].data, :2599-2610; the SchroFrameData copies made at
 *     schroframe.c:2470-2478 share its pointer), so by the time the metrics are taken (:1741-1750)
 *     the top-left chroma-block-sized corner of the "luma" prediction holds the V prediction, and
 *     the U metric is taken against the V prediction.
 *  3. With one reference width[] / height[] stay zero (:1644-1645), so the DC candidate is tried
 *     whenever the best error is positive (:1766).
 * Blocks outside the picture end up as {split 2, pred_mode 1, zero vector} with 2 bits of entropy
 * (:1650-1665 + the copy at :1511-1523); they never feed a block inside.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline int mini (int a, int b) { return a < b ? a : b; }

static int
med3 (int a, int b, int c)
{
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return c < lo ? lo : (c > hi ? hi : c);
}

#define MV_PRED_MODE(m) ((m)->flags & 3)
#define MV_GLOBAL(m) (((m)->flags >> 2) & 1)

static void
vector_prediction (const OracleMotionVector *motion, int nbx, int x, int y, int mode, int *px, int *py)
{
  int vx[3], vy[3], n = 0, k;
  const int dxs[3] = { -1, 0, -1 }, dys[3] = { 0, -1, -1 };
  for (k = 0; k < 3; k++) {
    const OracleMotionVector *mv;
    if ((dxs[k] && x == 0) || (dys[k] && y == 0)) continue;
    mv = &motion[(y + dys[k]) * nbx + x + dxs[k]];
    if (!MV_GLOBAL (mv) && (MV_PRED_MODE (mv) & mode)) {
      vx[n] = mv->v[mode - 1];
      vy[n] = mv->v[2 + mode - 1];
      n++;
    }
  }
  if (n == 0) { *px = 0; *py = 0; }
  else if (n == 1) { *px = vx[0]; *py = vy[0]; }
  else if (n == 2) { *px = (vx[0] + vx[1] + 1) >> 1; *py = (vy[0] + vy[1] + 1) >> 1; }
  else { *px = med3 (vx[0], vx[1], vx[2]); *py = med3 (vy[0], vy[1], vy[2]); }
}

/* entropy of a split-2, non-global vector block (:1266-1281) */
static int
vector_entropy (const OracleMotionVector *motion, int nbx, int x, int y)
{
  const OracleMotionVector *mv = &motion[y * nbx + x];
  int e = 0, px, py;
  if (MV_PRED_MODE (mv) & 1) {
    vector_prediction (motion, nbx, x, y, 1, &px, &py);
    e += oracle_bits_sint (mv->v[0] - px) + oracle_bits_sint (mv->v[2] - py);
  }
  if (MV_PRED_MODE (mv) & 2) {
    vector_prediction (motion, nbx, x, y, 2, &px, &py);
    e += oracle_bits_sint (mv->v[1] - px) + oracle_bits_sint (mv->v[3] - py);
  }
  return e;
}

void
oracle_split2_decide (const OracleSplit2Params *p, const uint8_t *const src[3], const int src_stride[3],
    const uint8_t *const ref0[3], const uint8_t *const ref1[3], const int rstride[3],
    const OracleMotionVector *field0, const OracleMotionVector *field1, OracleMotionVector *motion,
    int *sb_error, int *sb_entropy)
{
  const int nbx = p->x_num_blocks, nby = p->y_num_blocks, prec = p->mv_precision;
  const int comp_w[3] = { p->xblen, p->xblen >> p->h_shift, p->xblen >> p->h_shift };
  const int comp_h[3] = { p->yblen, p->yblen >> p->v_shift, p->yblen >> p->v_shift };
  const int cw = (p->width + (1 << p->h_shift) - 1) >> p->h_shift, ch = (p->height + (1 << p->v_shift) - 1) >> p->v_shift;
  const int plane_w[3] = { p->width, cw, cw }, plane_h[3] = { p->height, ch, ch };
  const int xmin = -p->orig_ext, ymin = -p->orig_ext;
  const int xmax = (p->width << prec) + p->orig_ext, ymax = (p->height << prec) + p->orig_ext;
  const uint8_t *const *refs[2] = { ref0, ref1 };
  const OracleMotionVector *fields[2] = { field0, field1 };
  int bx, by, k, r, a, b;

  memset (sb_error, 0, sizeof (int) * (size_t) ((nbx / 4) * (nby / 4)));
  memset (sb_entropy, 0, sizeof (int) * (size_t) ((nbx / 4) * (nby / 4)));
  memset (motion, 0, sizeof (*motion) * (size_t) (nbx * nby));
  for (by = 0; by < nby; by++)
    for (bx = 0; bx < nbx; bx++) {
      OracleMotionVector *mv = &motion[by * nbx + bx], best;
      const int sb = (by / 4) * (nbx / 4) + bx / 4;
      double min_score = HUGE_VAL, score;
      int entropy[2] = { 0, 0 }, width[3] = { 0, 0, 0 }, height[3] = { 0, 0, 0 }, w[3], h[3];
      int best_entropy = INT_MAX, best_error = INT_MAX, error;
      memset (&best, 0, sizeof (best));
      best.flags = (2u << 3) | 1u;
      if (!(p->width > bx * p->xblen) || !(p->height > by * p->yblen)) {
        *mv = best;
        sb_entropy[sb] += 2;
        continue;
      }
      for (k = 0; k < 3; k++) {
        w[k] = mini (comp_w[k], plane_w[k] - bx * comp_w[k]);
        h[k] = mini (comp_h[k], plane_h[k] - by * comp_h[k]);
      }
      /* one reference at a time */
      for (r = 0; r < p->num_refs; r++) {
        *mv = fields[r][by * nbx + bx];
        mv->flags = (mv->flags & ~0x1fu) | (2u << 3) | (uint32_t) (r + 1);
        entropy[r] = vector_entropy (motion, nbx, bx, by);
        if (mv->metric == (uint32_t) INT_MAX) {
          error = INT_MAX;
        } else {
          uint32_t chroma = 0;
          for (k = 1; k < 3; k++) {
            const int x = (mv->v[r] >> p->h_shift) + ((bx * comp_w[k]) << prec);
            const int y = (mv->v[2 + r] >> p->v_shift) + ((by * comp_h[k]) << prec);
            for (b = 0; b < h[k]; b++)
              for (a = 0; a < w[k]; a++)
                chroma += (uint32_t) abs ((int) src[k][(ptrdiff_t) (by * comp_h[k] + b) * src_stride[k] + bx * comp_w[k] + a]
                    - oracle_subpel_sample (refs[r][k], rstride[k], prec, x, y, a, b));
          }
          mv->chroma_metric = chroma;
          error = (int) (chroma + mv->metric);
        }
        score = entropy[r] + error * p->lambda;
        if (min_score > score) {
          min_score = score;
          best = *mv;
          best_entropy = entropy[r];
          best_error = (int) mv->metric;                /* property 1 */
        }
      }
      /* both references */
      if (p->num_refs > 1) {
        int biref = 1, pos_x[3][2], pos_y[3][2];
        mv->v[0] = field0[by * nbx + bx].v[0];
        mv->v[2] = field0[by * nbx + bx].v[2];
        mv->v[1] = field1[by * nbx + bx].v[1];
        mv->v[3] = field1[by * nbx + bx].v[3];
        mv->flags = (mv->flags & ~0x7u) | 3u;
        for (k = 0; k < 3; k++) {
          width[k] = w[k];
          height[k] = h[k];
          for (r = 0; r < 2; r++) {
            pos_x[k][r] = (k ? mv->v[r] >> p->h_shift : mv->v[r]) + bx * (comp_w[k] << prec);
            pos_y[k][r] = (k ? mv->v[2 + r] >> p->v_shift : mv->v[2 + r]) + by * (comp_h[k] << prec);
            if (k == 0 && biref && (xmin > pos_x[k][r] || ymin > pos_y[k][r]
                    || !(xmax > pos_x[k][r] + width[k] - 1) || !(ymax > pos_y[k][r] + height[k] - 1))) {
              biref = 0;
              break;
            }
          }
        }
        if (biref) {
          uint32_t m[3] = { 0, 0, 0 };
          for (k = 0; k < 3; k++)
            for (b = 0; b < height[k]; b++)
              for (a = 0; a < width[k]; a++) {
                /* property 2: which component's prediction the scratch block holds at (a, b) */
                const int kk = prec >= 2 ? ((a < width[2] && b < height[2]) ? 2 : (k == 1 ? 2 : k)) : k;
                const int p0 = oracle_subpel_sample (ref0[kk], rstride[kk], prec, pos_x[kk][0], pos_y[kk][0], a, b);
                const int p1 = oracle_subpel_sample (ref1[kk], rstride[kk], prec, pos_x[kk][1], pos_y[kk][1], a, b);
                m[k] += (uint32_t) abs ((int) src[k][(ptrdiff_t) (by * comp_h[k] + b) * src_stride[k] + bx * comp_w[k] + a]
                    - ((p0 + p1 + 1) >> 1));
              }
          mv->metric = m[0];
          mv->chroma_metric = m[1] + m[2];
          score = entropy[0] + entropy[1] + (mv->metric + mv->chroma_metric) * p->lambda;
          if (min_score > score) {
            best_error = (int) (mv->metric + mv->chroma_metric);
            best_entropy = entropy[0] + entropy[1];
            best = *mv;
            min_score = score;
          }
        }
      }
      /* DC block (property 3: width / height are zero with one reference) */
      if (4 * (width[0] * height[0] + 2 * width[1] * height[1]) < best_error) {
        int dc_entropy;
        mv->flags = (mv->flags & ~0x1fu) | (2u << 3);
        error = 0;
        for (k = 0; k < 3; k++) {
          int sum = 0, n = w[k] * h[k], ave;
          for (b = 0; b < h[k]; b++)
            for (a = 0; a < w[k]; a++)
              sum += src[k][(ptrdiff_t) (by * comp_h[k] + b) * src_stride[k] + bx * comp_w[k] + a];
          ave = (sum + n / 2) / n;
          for (b = 0; b < h[k]; b++)
            for (a = 0; a < w[k]; a++)
              error += abs (ave - (int) src[k][(ptrdiff_t) (by * comp_h[k] + b) * src_stride[k] + bx * comp_w[k] + a]);
          mv->v[k] = (int16_t) (ave - 128);
        }
        mv->metric = (uint32_t) error;
        dc_entropy = oracle_bits_sint (mv->v[0]) + oracle_bits_sint (mv->v[1]) + oracle_bits_sint (mv->v[2]);
        if (error < best_error) {
          best = *mv;
          best_error = error;
          best_entropy = dc_entropy;
        }
      }
      *mv = best;
      sb_error[sb] = (int) ((unsigned) sb_error[sb] + (unsigned) best_error);
      sb_entropy[sb] = (int) ((unsigned) sb_entropy[sb] + (unsigned) best_entropy);
    }
}
