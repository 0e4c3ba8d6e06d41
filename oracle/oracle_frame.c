/*
 * oracle_frame.c -- CPU restatement of the reference-frame operations of the picture
 * core: motion-compensation edge extension, the 8-tap half-pel upsampler with its
 * border rules, and the 4-tap pyramid downsampler.  TEST INFRASTRUCTURE (oracle.h).
 *
 * Written as pure per-pixel definitions (every output pixel as a closed-form function
 * of the source plane), unlike the reference's pass-by-pass in-place code:
 *   schro_frame_mc_edgeextend         schroedinger/schroframe.c:1940-1997
 *   schro_frame_upsample_horiz/_vert  schroedinger/schroframe.c:1515-1645
 *   schro_upsampled_frame_upsample    schroedinger/schroframe.c:2000-2030
 *   schro_frame_downsample            schroedinger/schroframe.c:1400-1513,
 *                                     schroedinger/schroorc.orc:1345-1397
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline int clampi (int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

static const int up_taps[8] = { -1, 3, -7, 21, 21, -7, 3, -1 };

/* schroframe.c:1950-1990: every row gets its first pixel replicated to the left and its
 * last pixel to the right, then rows 0 / h-1 (with their borders) are replicated up/down */
void
oracle_mc_edgeextend (uint8_t *data, int stride, int width, int height, int ext)
{
  int x, y;
  for (y = -ext; y < height + ext; y++) {
    uint8_t *row = data + (ptrdiff_t) stride * y;
    const uint8_t *src = data + (ptrdiff_t) stride * clampi (y, 0, height - 1);
    for (x = -ext; x < width + ext; x++) {
      if (y >= 0 && y < height && x >= 0 && x < width) continue;
      row[x] = src[clampi (x, 0, width - 1)];
    }
  }
}

typedef struct { const uint8_t *p0; int stride, w, h; } Src;

static inline int p0_at (const Src *s, int x, int y)
{
  return s->p0[(ptrdiff_t) s->stride * clampi (y, 0, s->h - 1) + clampi (x, 0, s->w - 1)];
}

/* vertical half-pel sample between rows y and y+1 (schroframe.c:1576-1645):
 * rows 0..h-2 are filtered with clamped row indices, row h-1 is a copy of the source */
static inline int vphase (const Src *s, int x, int y)
{
  int j, acc = 16;
  if (y >= s->h - 1) return p0_at (s, x, s->h - 1);
  for (j = 0; j < 8; j++) acc += up_taps[j] * p0_at (s, x, y + j - 3);
  return clampi (acc >> 5, 0, 255);
}

/* horizontal half-pel sample of a row given by a getter (schroframe.c:1515-1574):
 * columns 0..w-2 filtered with clamped column indices; column w-1 ends up a copy of the
 * source (the n>8 copy at :1552-1553, and the border fill at :1957-1962 for any n) */
static inline int hphase_p0 (const Src *s, int x, int y)
{
  int j, acc = 16;
  if (x >= s->w - 1) return p0_at (s, s->w - 1, y);
  for (j = 0; j < 8; j++) acc += up_taps[j] * p0_at (s, clampi (x + j - 3, 0, s->w - 1), y);
  return clampi (acc >> 5, 0, 255);
}

static inline int hphase_v (const Src *s, int x, int y)
{
  int j, acc = 16;
  if (x >= s->w - 1) return vphase (s, s->w - 1, y);
  for (j = 0; j < 8; j++) acc += up_taps[j] * vphase (s, clampi (x + j - 3, 0, s->w - 1), y);
  return clampi (acc >> 5, 0, 255);
}

/* Final content of phases 1..3 at any coordinate of the extended plane, following the
 * fill order of schro_upsampled_frame_upsample (schroframe.c:2018-2028). */
static int phase1 (const Src *s, int x, int y)
{
  int yy = clampi (y, 0, s->h - 1);       /* rows above/below replicate phase 1's own rows */
  if (x < 0) return p0_at (s, 0, yy);     /* left border comes from phase 0 */
  return hphase_p0 (s, x, yy);            /* x >= w-1 gives phase 0's last pixel */
}

static int phase2 (const Src *s, int x, int y)
{
  if (y < 0) return p0_at (s, x, 0);              /* rows above: phase 0 row 0 with its border */
  if (y >= s->h - 1) return p0_at (s, x, s->h - 1); /* last row and below: phase 0 row h-1 */
  return vphase (s, clampi (x, 0, s->w - 1), y);  /* side borders replicate phase 2 itself */
}

static int phase3 (const Src *s, int x, int y)
{
  if (y < 0) return phase1 (s, x, 0);             /* rows above: phase 1 row 0 with its border */
  if (y >= s->h - 1) return phase1 (s, x, s->h - 1);
  if (x < 0) return vphase (s, 0, y);             /* side borders come from phase 2 */
  return hphase_v (s, x, y);
}

void
oracle_upsample (uint8_t *data, int stride, int width, int height, int ext)
{
  Src s = { data, stride, width, height };
  int q = stride >> 2;
  int x, y;
  /* phase 0 must be read before anything is written: snapshot it (the phases do not
   * overlap in memory, so this is only for clarity of the restatement) */
  for (y = -ext; y < height + ext; y++) {
    uint8_t *row = data + (ptrdiff_t) stride * y;
    for (x = -ext; x < width + ext; x++) {
      row[q + x] = (uint8_t) phase1 (&s, x, y);
      row[2 * q + x] = (uint8_t) phase2 (&s, x, y);
      row[3 * q + x] = (uint8_t) phase3 (&s, x, y);
    }
  }
}

/* schroframe.c:1400-1513: vertical (6,26,26,6)+32>>6 with an 8-bit intermediate, then the
 * same horizontally; indices clamped */
void
oracle_downsample (uint8_t *dest, int dstride, int dwidth, int dheight,
    const uint8_t *src, int sstride, int swidth, int sheight)
{
  uint8_t *tmp = malloc ((size_t) swidth);
  int x, y;
  for (y = 0; y < dheight; y++) {
    const uint8_t *r0 = src + (ptrdiff_t) sstride * clampi (2 * y - 1, 0, sheight - 1);
    const uint8_t *r1 = src + (ptrdiff_t) sstride * clampi (2 * y + 0, 0, sheight - 1);
    const uint8_t *r2 = src + (ptrdiff_t) sstride * clampi (2 * y + 1, 0, sheight - 1);
    const uint8_t *r3 = src + (ptrdiff_t) sstride * clampi (2 * y + 2, 0, sheight - 1);
    for (x = 0; x < swidth; x++)
      tmp[x] = (uint8_t) ((6 * (r0[x] + r3[x]) + 26 * (r1[x] + r2[x]) + 32) >> 6);
    for (x = 0; x < dwidth; x++) {
      int a = tmp[clampi (2 * x - 1, 0, swidth - 1)], b = tmp[clampi (2 * x, 0, swidth - 1)];
      int c = tmp[clampi (2 * x + 1, 0, swidth - 1)], d = tmp[clampi (2 * x + 2, 0, swidth - 1)];
      dest[(ptrdiff_t) dstride * y + x] = (uint8_t) clampi ((6 * (a + d) + 26 * (b + c) + 32) >> 6, 0, 255);
    }
  }
  free (tmp);
}

/* ---- combine / convert glue ------------------------------------------------------------ */
static int
load_sample (const void *row, int depth, int x)
{
  if (depth == 0) return ((const uint8_t *) row)[x];
  if (depth == 1) return ((const int16_t *) row)[x];
  return ((const int32_t *) row)[x];
}

/* one sample through the reference's converter chain (schroframe.c:905-925 picks the pair) */
static int
convert_sample (int v, int sdepth, int ddepth)
{
  if (sdepth == ddepth) return v;
  if (ddepth == 0) {
    int t;
    if (sdepth == 1) {
      t = (int16_t) (v + 128);                                   /* addw wraps (schroorc.orc:504-511) */
    } else {
      /* the shipped program (schroorc-dist.c:4274-4306, what the library runs) is addl, convsuslw,
       * convsuswb: the sum wraps at 32 bits, saturates to 0..65535, and that word is then read as
       * SIGNED by the byte narrowing -- so 32768..65535 come out as 0, not 255.  (schroorc.orc:513-521
       * says convssslw; the generated file is the authority.) */
      int32_t w = (int32_t) ((uint32_t) v + 128u);
      t = (int16_t) (w < 0 ? 0 : w > 65535 ? 65535 : w);
    }
    return t < 0 ? 0 : t > 255 ? 255 : t;                        /* convsuswb */
  }
  if (ddepth == 1) {
    if (sdepth == 0) return v - 128;                             /* convubw, subw (:524-531) */
    return (int16_t) v;                                          /* convlw truncates (:483-487) */
  }
  if (sdepth == 0) return v - 128;                               /* convubw, subw, convswl (:533-540) */
  return v;                                                      /* convswl (:497-501) */
}

void
oracle_convert_plane (void *dst, int dstride, int ddepth, int dwidth, int dheight,
    const void *src, int sstride, int sdepth, int swidth, int sheight)
{
  int x, y;
  for (y = 0; y < dheight; y++) {
    const uint8_t *srow = (const uint8_t *) src + (ptrdiff_t) (y < sheight ? y : sheight - 1) * sstride;
    uint8_t *drow = (uint8_t *) dst + (ptrdiff_t) y * dstride;
    for (x = 0; x < dwidth; x++) {
      const int v = convert_sample (load_sample (srow, sdepth, x < swidth ? x : swidth - 1), sdepth, ddepth);
      if (ddepth == 0) drow[x] = (uint8_t) v;
      else if (ddepth == 1) ((int16_t *) drow)[x] = (int16_t) v;
      else ((int32_t *) drow)[x] = v;
    }
  }
}

void
oracle_add_plane (int16_t *dst, int dstride, int dwidth, int dheight,
    const void *src, int sstride, int sdepth, int swidth, int sheight, int subtract)
{
  const int w = dwidth < swidth ? dwidth : swidth, h = dheight < sheight ? dheight : sheight;
  int x, y;
  for (y = 0; y < h; y++) {
    const uint8_t *srow = (const uint8_t *) src + (ptrdiff_t) y * sstride;
    int16_t *drow = (int16_t *) ((uint8_t *) dst + (ptrdiff_t) y * dstride);
    for (x = 0; x < w; x++) {
      const int v = load_sample (srow, sdepth, x);
      drow[x] = (int16_t) (subtract ? drow[x] - v : drow[x] + v);   /* addw / subw wrap */
    }
  }
}
