/*
 * oracle_motion.c -- CPU restatement of the OBMC motion-compensation renderer.
 * TEST INFRASTRUCTURE (oracle.h).
 *
 * The reference scatters: for every block it fetches a (sub-pel) reference block and
 * multiply-accumulates it with the OBMC window into an s16 strip
 * (schro_motion_render_u8, schroedinger/schromotion8.c:700-929).  This restatement
 * gathers: every output pixel sums the contributions of the blocks covering it.  The
 * two are the same modulo 2^16 (Orc addw/mullw wrap; schromotion8.c:15-167), which is
 * what the s16 accumulator holds.
 *
 * Follows: weights schroedinger/schromotion.c:40-93; block fetch
 * schromotion8.c:303-335 + schroedinger/schroframe.c:2166-2482; per-mode prediction
 * schromotion8.c:336-657; border blocks :659-698; strip finish schroorc.orc:636-673.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline int clampi (int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int w16 (int x) { return (int) (int16_t) x; }
static inline int mini (int a, int b) { return a < b ? a : b; }

/* schromotion.c:40-49 */
static int
get_ramp (int x, int offset)
{
  if (offset == 1) return x == 0 ? 3 : 5;
  return 1 + (6 * x + offset - 1) / (2 * offset - 1);
}

/* schromotion.c:52-79 */
static void
obmc_weights (int *w, int len, int off)
{
  int i;
  for (i = 0; i < len; i++) {
    if (off == 0) w[i] = 8;
    else if (i < 2 * off) w[i] = get_ramp (i, off);
    else if (len - 1 - i < 2 * off) w[i] = get_ramp (len - 1 - i, off);
    else w[i] = 8;
  }
}

typedef struct {
  const OracleObmcParams *p;
  int width, height;
  int xoff, yoff;
  int max_fast_x, max_fast_y;
  int rstride;
} Ctx;

/* half-pel sample (u,v): phase ((v&1)<<1)|(u&1) at (u>>1, v>>1)   schroframe.c:2186-2200 */
static inline int
halfpel (const uint8_t *ref, int rstride, int u, int v, int a, int b)
{
  int ph = ((v & 1) << 1) | (u & 1);
  return ref[(ptrdiff_t) ph * (rstride >> 2) + (ptrdiff_t) ((v >> 1) + b) * rstride + (u >> 1) + a];
}

/* reference sample of block (i,j) at block pixel (a,b)
 * schromotion8.c:303-335, schroframe.c:2288-2482 */
static int
fetch (const Ctx *c, const uint8_t *ref, int i, int j, int a, int b, int dx, int dy)
{
  const OracleObmcParams *p = c->p;
  int prec = p->mv_precision;
  int px, py, exp;
  dx >>= p->h_shift;
  dy >>= p->v_shift;
  px = ((p->xbsep * i - c->xoff) << prec) + dx;
  py = ((p->ybsep * j - c->yoff) << prec) + dy;
  exp = 32 << prec;
  px = clampi (px, -exp, c->max_fast_x + exp - 1);
  py = clampi (py, -exp, c->max_fast_y + exp - 1);
  if (prec == 0)
    return ref[(ptrdiff_t) (py + b) * c->rstride + px + a];
  if (prec == 1)
    return halfpel (ref, c->rstride, px, py, a, b);
  if (prec == 2) { px <<= 1; py <<= 1; }
  {
    int hx = px >> 2, hy = py >> 2, rx = px & 3, ry = py & 3;
    if (rx == 0 && ry == 0)
      return halfpel (ref, c->rstride, hx, hy, a, b);
    if (ry == 0 && rx == 2)
      return (halfpel (ref, c->rstride, hx, hy, a, b) + halfpel (ref, c->rstride, hx + 1, hy, a, b) + 1) >> 1;
    if (ry == 2 && rx == 0)
      return (halfpel (ref, c->rstride, hx, hy, a, b) + halfpel (ref, c->rstride, hx, hy + 1, a, b) + 1) >> 1;
    {
      /* orc_combine4_nxm_u8, schroorc.orc:1635-1662: 16-bit, +8 >>4, saturate to u8 */
      int w00 = (4 - ry) * (4 - rx), w01 = (4 - ry) * rx, w10 = ry * (4 - rx), w11 = ry * rx;
      int t = w16 (w00 * halfpel (ref, c->rstride, hx, hy, a, b));
      t = w16 (t + w16 (w01 * halfpel (ref, c->rstride, hx + 1, hy, a, b)));
      t = w16 (t + w16 (w10 * halfpel (ref, c->rstride, hx, hy + 1, a, b)));
      t = w16 (t + w16 (w11 * halfpel (ref, c->rstride, hx + 1, hy + 1, a, b)));
      t = w16 (t + 8) >> 4;
      return clampi (t, 0, 255);
    }
  }
}

void
oracle_obmc_render (const OracleObmcParams *p, const OracleMotionVector *mvs,
    const uint8_t *ref0, const uint8_t *ref1, int rstride, int width, int height,
    int16_t *acc, int acc_stride, void *residual, int res_stride, int res_is_s32,
    int add, uint8_t *out, int out_stride)
{
  Ctx c;
  int wx[64], wy[64];
  int x, y;
  const int simple = (p->weight1 == 1 && p->weight2 == 1 && p->weight_bits == 1);
  const int noscale = (p->weight1 + p->weight2 == (1 << p->weight_bits));
  int max_x_blocks, max_y_blocks;

  c.p = p;
  c.width = width;
  c.height = height;
  c.xoff = (p->xblen - p->xbsep) / 2;
  c.yoff = (p->yblen - p->ybsep) / 2;
  c.max_fast_x = (width - p->xblen) << p->mv_precision;
  c.max_fast_y = (height - p->yblen) << p->mv_precision;
  c.rstride = rstride;
  obmc_weights (wx, p->xblen, c.xoff);
  obmc_weights (wy, p->yblen, c.yoff);
  /* blocks [1,max) x [1,max) take the fast (whole-block) path, schromotion8.c:795-853 */
  max_x_blocks = mini (p->x_num_blocks - 1, (width - c.xoff) / p->xbsep);
  max_y_blocks = mini (p->y_num_blocks - 1, (height - c.yoff) / p->ybsep);

  for (y = 0; y < height; y++) {
    for (x = 0; x < width; x++) {
      int sum = 0;
      int j, i;
      /* only blocks with  bsep*k - off <= pos < bsep*k - off + blen  cover this pixel */
      int j0 = (y + c.yoff - p->yblen + 1 > 0) ? (y + c.yoff - p->yblen + p->ybsep) / p->ybsep : 0;
      int j1 = mini (p->y_num_blocks - 1, (y + c.yoff) / p->ybsep);
      int i0 = (x + c.xoff - p->xblen + 1 > 0) ? (x + c.xoff - p->xblen + p->xbsep) / p->xbsep : 0;
      int i1 = mini (p->x_num_blocks - 1, (x + c.xoff) / p->xbsep);
      for (j = j0; j <= j1; j++) {
        int b = y - (p->ybsep * j - c.yoff);
        if (b < 0 || b >= p->yblen) continue;
        for (i = i0; i <= i1; i++) {
          int a = x - (p->xbsep * i - c.xoff);
          const OracleMotionVector *mv = &mvs[j * p->x_num_blocks + i];
          int mode = mv->flags & 3;
          int fast = (i >= 1 && i < max_x_blocks && j >= 1 && j < max_y_blocks);
          int v;
          if (a < 0 || a >= p->xblen) continue;
          if (mode == 0) {
            v = mv->v[p->comp] + 128;
            v = fast ? w16 (v) : (uint8_t) v;       /* :343-355 vs :570-580 */
          } else if (mode == 3) {
            int s0 = fetch (&c, ref0, i, j, a, b, mv->v[0], mv->v[2]);
            int s1 = fetch (&c, ref1, i, j, a, b, mv->v[1], mv->v[3]);
            if (simple) {
              v = (s0 + s1 + 1) >> 1;                /* avgub */
            } else if (fast) {                       /* block_acc_biref, :127-167 */
              int t = w16 (s0 * w16 (p->weight1 << (6 - p->weight_bits)));
              int u = w16 (s1 * w16 (p->weight2 << (6 - p->weight_bits)));
              t = w16 (t + u);
              t = w16 (t + 32);
              v = t >> 6;
            } else {                                 /* orc_combine2_nxm_u8, schroorc.orc:1737-1756 */
              int t = w16 (s0 * w16 (p->weight1));
              int u = w16 (s1 * w16 (p->weight2));
              t = w16 (t + u);
              t = w16 (t + ((1 << p->weight_bits) >> 1));
              v = clampi (t >> p->weight_bits, 0, 255);
            }
          } else {
            int s = (mode == 1) ? fetch (&c, ref0, i, j, a, b, mv->v[0], mv->v[2])
                                : fetch (&c, ref1, i, j, a, b, mv->v[1], mv->v[3]);
            if (fast) {
              if (simple) {
                v = s;
              } else {                               /* block_acc_scaled, :41-71 */
                int t = w16 (s * w16 ((p->weight1 + p->weight2) << (6 - p->weight_bits)));
                t = w16 (t + 32);
                v = t >> 6;
              }
            } else {
              if (noscale) v = s;                    /* :384-398 */
              else v = (uint8_t) ((s * (p->weight1 + p->weight2) + (1 << (p->weight_bits - 1))) >> p->weight_bits);
            }
          }
          if (fast) {
            sum += v * (wx[a] * wy[b]);
          } else {
            /* schro_motion_block_accumulate_slow, :659-698: border blocks absorb the
             * weight of the neighbour that does not exist */
            int w_x = wx[a], w_y = wy[b];
            if (x < c.xoff) w_x += wx[2 * c.xoff - a - 1];
            if (x >= p->x_num_blocks * p->xbsep - c.xoff) w_x += wx[2 * (p->xblen - c.xoff) - a - 1];
            if (y < c.yoff) w_y += wy[2 * c.yoff - b - 1];
            if (y >= p->y_num_blocks * p->ybsep - c.yoff) w_y += wy[2 * (p->yblen - c.yoff) - b - 1];
            sum += v * w_x * w_y;
          }
        }
      }
      {
        int a16 = w16 (sum);
        if (add) {
          /* orc_rrshift6_add_s16_2d / _s32_2d, schroorc.orc:636-660 */
          int r = res_is_s32
              ? w16 (((const int32_t *) ((const char *) residual + (ptrdiff_t) res_stride * y))[x])
              : ((const int16_t *) ((const char *) residual + (ptrdiff_t) res_stride * y))[x];
          int t = w16 (a16 + 32) >> 6;
          t = w16 (r + t);
          out[(ptrdiff_t) out_stride * y + x] = (uint8_t) clampi (t, 0, 255);
          if (acc) ((int16_t *) ((char *) acc + (ptrdiff_t) acc_stride * y))[x] = (int16_t) a16;
        } else {
          /* orc_rrshift6_sub_s16_2d, schroorc.orc:663-673 */
          int16_t *r = (int16_t *) ((char *) residual + (ptrdiff_t) res_stride * y) + x;
          int t = w16 (a16 - 8160) >> 6;
          *r = (int16_t) w16 (*r - t);
          if (acc) ((int16_t *) ((char *) acc + (ptrdiff_t) acc_stride * y))[x] = (int16_t) t;
        }
      }
    }
  }
}

/* ---- the reference's per-pixel renderer, which it uses whenever global motion is on ----------------
 * schro_motion_render_ref (schroedinger/schromotionref.c:245-330): a pixel is the rounded sum of the (at
 * most four) blocks covering it, each block's pixel weighted by the OBMC ramp (:175-236: picture-edge
 * pixels take weight 8), fetched pixel by pixel with clamped coordinates
 * (schro_upsampled_frame_get_pixel_precN, schroedinger/schroframe.c:2033-2046, 2124-2143, 2209-2265); a
 * block flagged using_global takes its vector from the picture's global-motion model at that pixel
 * (:22-41).  The prediction is clamped to 0..255 BEFORE the residual is added, unlike the block renderer. */
static int
pixel_prec1 (const uint8_t *ref, int rstride, int w, int h, int x, int y)
{
  x = clampi (x, 0, w * 2 - 2);
  y = clampi (y, 0, h * 2 - 2);
  return halfpel (ref, rstride, x, y, 0, 0);
}

static int
pixel_precn (const uint8_t *ref, int rstride, int w, int h, int x, int y, int prec)
{
  int hx, hy, rx, ry, v;
  if (prec == 0) return ref[(ptrdiff_t) clampi (y, 0, h - 1) * rstride + clampi (x, 0, w - 1)];
  if (prec == 1) return pixel_prec1 (ref, rstride, w, h, x, y);
  if (prec == 2) { x <<= 1; y <<= 1; }
  hx = x >> 2; hy = y >> 2; rx = x & 3; ry = y & 3;
  v = (4 - ry) * (4 - rx) * pixel_prec1 (ref, rstride, w, h, hx, hy)
      + (4 - ry) * rx * pixel_prec1 (ref, rstride, w, h, hx + 1, hy)
      + ry * (4 - rx) * pixel_prec1 (ref, rstride, w, h, hx, hy + 1)
      + ry * rx * pixel_prec1 (ref, rstride, w, h, hx + 1, hy + 1);
  return (v + 8) >> 4;
}

static void
global_vector (const int *gm, int x, int y, int *dx, int *dy)
{
  /* gm: b0 b1 a_exp a00 a01 a10 a11 c_exp c0 c1 (SchroGlobalMotion, schroedinger/schroparams.h:18-29) */
  const int alpha = gm[2], beta = gm[7];
  const int scale = (1 << beta) - (gm[8] * x + gm[9] * y);
  *dx = (scale * (gm[3] * x + gm[4] * y + (1 << alpha) * gm[0])) >> (alpha + beta);
  *dy = (scale * (gm[5] * x + gm[6] * y + (1 << alpha) * gm[1])) >> (alpha + beta);
}

void
oracle_obmc_render_ref (const OracleObmcParams *p, const int *global_motion, const OracleMotionVector *mvs,
    const uint8_t *ref0, const uint8_t *ref1, int rstride, int width, int height, int16_t *acc, int acc_stride,
    int16_t *residual, int res_stride, int add, uint8_t *out, int out_stride)
{
  const int xoff = (p->xblen - p->xbsep) / 2, yoff = (p->yblen - p->ybsep) / 2;
  const int W = p->xbsep * p->x_num_blocks, H = p->ybsep * p->y_num_blocks;
  int x, y, di, dj;
  for (y = 0; y < height; y++)
    for (x = 0; x < width; x++) {
      const int i0 = (x + xoff) / p->xbsep - 1, j0 = (y + yoff) / p->ybsep - 1;
      int value = 0, line;
      for (dj = 0; dj < 2; dj++)
        for (di = 0; di < 2; di++) {
          const int i = i0 + di, j = j0 + dj;
          int xmin, ymin, xmax, ymax, wx, wy, v = 0, mode;
          const OracleMotionVector *mv;
          if (i < 0 || j < 0 || i >= p->x_num_blocks || j >= p->y_num_blocks) continue;
          xmin = i * p->xbsep - xoff; ymin = j * p->ybsep - yoff;
          xmax = (i + 1) * p->xbsep + xoff; ymax = (j + 1) * p->ybsep + yoff;
          if (x < xmin || y < ymin || x >= xmax || y >= ymax) continue;
          if (xoff == 0 || x < xoff || x >= W - xoff) wx = 8;
          else if (x - xmin < 2 * xoff) wx = get_ramp (x - xmin, xoff);
          else if (xmax - 1 - x < 2 * xoff) wx = get_ramp (xmax - 1 - x, xoff);
          else wx = 8;
          if (yoff == 0 || y < yoff || y >= H - yoff) wy = 8;
          else if (y - ymin < 2 * yoff) wy = get_ramp (y - ymin, yoff);
          else if (ymax - 1 - y < 2 * yoff) wy = get_ramp (ymax - 1 - y, yoff);
          else wy = 8;
          mv = &mvs[j * p->x_num_blocks + i];
          mode = mv->flags & 3;
          if (mode == 0) v = mv->v[p->comp] + 128;
          else {
            int d[2][2], r, s[2] = { 0, 0 };
            for (r = 0; r < 2; r++) {
              if (!((mode >> r) & 1)) continue;
              if ((mv->flags >> 2) & 1) global_vector (global_motion + 10 * r, x, y, &d[r][0], &d[r][1]);
              else { d[r][0] = mv->v[r]; d[r][1] = mv->v[2 + r]; }
              d[r][0] >>= p->h_shift; d[r][1] >>= p->v_shift;
              s[r] = pixel_precn (r ? ref1 : ref0, rstride, width, height, (x << p->mv_precision) + d[r][0],
                  (y << p->mv_precision) + d[r][1], p->mv_precision);
            }
            if (mode == 3) v = p->weight1 * s[0] + p->weight2 * s[1];
            else v = (p->weight1 + p->weight2) * s[mode == 1 ? 0 : 1];
            v = (v + (1 << (p->weight_bits - 1))) >> p->weight_bits;
          }
          value += v * wx * wy;
        }
      line = clampi ((value + 32) >> 6, 0, 255) - 128;
      if (acc) acc[(ptrdiff_t) y * acc_stride + x] = (int16_t) line;
      if (add) out[(ptrdiff_t) y * out_stride + x] = (uint8_t) clampi (residual[(ptrdiff_t) y * res_stride + x] + line + 128, 0, 255);
      else residual[(ptrdiff_t) y * res_stride + x] = (int16_t) (residual[(ptrdiff_t) y * res_stride + x] - line);
    }
}
