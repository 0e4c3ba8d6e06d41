/*
 * schro_host_frame.c -- the reference's frame-preparation, OBMC and SAD entry points on top
 * of the CUDA layer.
 *
 *   schro_frame_mc_edgeextend        schroedinger/schroframe.c:1986-1997
 *   schro_upsampled_frame_upsample   schroedinger/schroframe.c:2000-2030
 *   schro_frame_upsample_horiz/vert  schroedinger/schroframe.c:1557-1645
 *   schro_frame_downsample           schroedinger/schroframe.c:1505-1513
 *   schro_motion_new / _free         schroedinger/schromotion.c:14-38
 *   schro_motion_render[_u8]         schroedinger/schromotion.c:95-155, schromotion8.c:700
 *   schro_motion_init_obmc_weight    schroedinger/schromotion.c:52-93
 *   schro_metric_absdiff_u8 / _get / _get_dc / _get_biref   schroedinger/schrometric.c:10-304
 *
 * Frames in host memory are staged H2D / D2H on the calling thread's stream (whole frame
 * regions in one DMA each); frames from the CUDA domain are used where they lie.
 */
#include "schro_host.h"
#include <stdlib.h>
#include <string.h>

/* A frame as the CUDA layer sees it: a one-picture slab plus, for host frames, the
 * device staging copy of the whole region. */
typedef struct {
  sb2_slab slab;
  SchroFrame *frame;
  void *dev_region;     /* NULL for device frames */
  size_t region_bytes;
  int bpp;
} Staged;

static size_t
frame_region_bytes (const SchroFrame *f)
{
  return (size_t) f->components[0].length + (size_t) f->components[1].length +
      (size_t) f->components[2].length;
}

static void
stage_describe (Staged *s, SchroFrame *f, char *base)
{
  int k;
  memset (&s->slab, 0, sizeof (s->slab));
  s->slab.base = base;
  s->slab.picture_pitch = s->region_bytes;
  s->slab.count = 1;
  s->slab.ncomp = 3;
  for (k = 0; k < 3; k++) {
    s->slab.offset[k] = (size_t) ((char *) f->components[k].data - (char *) f->regions[0]);
    s->slab.stride[k] = f->components[k].stride;
    s->slab.width[k] = f->components[k].width;
    s->slab.height[k] = f->components[k].height;
  }
}

/* upload != 0: copy the region to the device (host frames only) */
static void
stage_in (Sb2hContext *cx, Staged *s, SchroFrame *f, int which_buf, int upload)
{
  s->frame = f;
  s->bpp = sb2h_bpp (f->format);
  s->region_bytes = frame_region_bytes (f);
  SB2H_ASSERT (f->regions[0] != NULL);
  if (sb2h_mem_kind (f->regions[0]) == SB2H_MEM_DEVICE) {
    s->dev_region = NULL;
    stage_describe (s, f, f->regions[0]);
    sb2h_frame_use (cx, f->regions[0]);
    return;
  }
  s->dev_region = sb2h_dev_buffer (cx, which_buf, s->region_bytes + 256);
  stage_describe (s, f, s->dev_region);
  if (upload)
    SB2H_CUDA (cudaMemcpyAsync (s->dev_region, f->regions[0], s->region_bytes, cudaMemcpyDefault,
            cx->stream));
}

static void
stage_out (Sb2hContext *cx, Staged *s)
{
  if (s->dev_region)
    SB2H_CUDA (cudaMemcpyAsync (s->frame->regions[0], s->dev_region, s->region_bytes,
            cudaMemcpyDefault, cx->stream));
}

/* End of a call that used the staged frames st[0..n): if any of them lives in host memory its
 * result must be there when the call returns, so wait; otherwise leave the work in flight and
 * note which device frames (bit i of `written`) it writes. */
static void
stage_finish (Sb2hContext *cx, Staged **st, int n, unsigned written)
{
  int i, host = 0;
  for (i = 0; i < n; i++)
    if (st[i] && st[i]->dev_region) host = 1;
  if (host) {
    sb2h_sync (cx);
    return;
  }
  for (i = 0; i < n; i++)
    if (st[i] && ((written >> i) & 1)) sb2h_frame_wrote (cx, st[i]->frame->regions[0]);
  cx->dirty = 1;
}

static void
require_u8 (const SchroFrame *f, const char *who)
{
  if (SCHRO_FRAME_FORMAT_DEPTH (f->format) != SCHRO_FRAME_FORMAT_DEPTH_U8)
    sb2h_fatal (who, "needs a u8 frame (format 0x%x)", (unsigned) f->format);
}

void
schro_frame_mc_edgeextend (SchroFrame *frame)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s;
  require_u8 (frame, __func__);
  stage_in (cx, &s, frame, SB2H_BUF_IN, 1);
  SB2H_CHECK (sb2_mc_edgeextend (&s.slab, frame->extension, 0, cx->stream), "sb2_mc_edgeextend");
  stage_out (cx, &s);
  {
    Staged *st[1] = { &s };
    stage_finish (cx, st, 1, 1u);
  }
}

void
schro_upsampled_frame_upsample (SchroFrame *df)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s;
  if (df->upsample_done) return;
  SB2H_ASSERT (df->is_upsampled);
  require_u8 (df, __func__);
  df->upsample_done = 1;
  stage_in (cx, &s, df, SB2H_BUF_IN, 1);
  SB2H_CHECK (sb2_upsample (&s.slab, df->extension, cx->stream), "sb2_upsample");
  stage_out (cx, &s);
  {
    Staged *st[1] = { &s };
    stage_finish (cx, st, 1, 1u);
  }
}

static void
upsample_1d (SchroFrameData *dest, SchroFrameData *src, int vertical)
{
  Sb2hContext *cx = sb2h_context ();
  const int w = src->width, h = vertical ? dest->height : dest->height;
  if (SCHRO_FRAME_FORMAT_DEPTH (dest->format) != SCHRO_FRAME_FORMAT_DEPTH_U8 ||
      src->format != dest->format)
    sb2h_fatal (__func__, "unimplemented format (as the reference, schroframe.c:1563-1568)");
  if (sb2h_mem_kind (src->data) == SB2H_MEM_DEVICE) sb2h_ptr_use (cx, src->data);
  if (sb2h_mem_kind (dest->data) == SB2H_MEM_DEVICE) sb2h_ptr_use (cx, dest->data);
  if (sb2h_mem_kind (src->data) == SB2H_MEM_DEVICE && sb2h_mem_kind (dest->data) == SB2H_MEM_DEVICE) {
    SB2H_CHECK (sb2_upsample_plane_1d (dest->data, dest->stride, src->data, src->stride, w, h,
            vertical, cx->stream), "sb2_upsample_plane_1d");
    sb2h_ptr_wrote (cx, dest->data);
  } else {
    const size_t pitch = ((size_t) w + 15) & ~(size_t) 15;
    uint8_t *din = sb2h_dev_buffer (cx, SB2H_BUF_IN, pitch * h);
    uint8_t *dout = sb2h_dev_buffer (cx, SB2H_BUF_OUT, pitch * h);
    sb2h_copy_rect (cx, din, pitch, src->data, src->stride, w, h);
    SB2H_CHECK (sb2_upsample_plane_1d (dout, (int) pitch, din, (int) pitch, w, h, vertical,
            cx->stream), "sb2_upsample_plane_1d");
    sb2h_copy_rect (cx, dest->data, dest->stride, dout, pitch, w, h);
  }
  sb2h_sync (cx);
}

void
schro_frame_upsample_horiz (SchroFrameData *dest, SchroFrameData *src)
{
  upsample_1d (dest, src, 0);
}

void
schro_frame_upsample_vert (SchroFrameData *dest, SchroFrameData *src)
{
  upsample_1d (dest, src, 1);
}

void
schro_frame_downsample (SchroFrame *dest, SchroFrame *src)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s, d;
  require_u8 (src, __func__);
  require_u8 (dest, __func__);
  stage_in (cx, &s, src, SB2H_BUF_IN, 1);
  stage_in (cx, &d, dest, SB2H_BUF_OUT, 1);     /* keep dest's borders as they are */
  SB2H_CHECK (sb2_downsample (&s.slab, &d.slab, cx->stream), "sb2_downsample");
  stage_out (cx, &d);
  {
    Staged *st[2] = { &s, &d };
    stage_finish (cx, st, 2, 2u);
  }
}

/* ---- combine / convert glue (schroframe.c:870-1182) --------------------------- */
static int
depth_code (SchroFrameFormat format, const char *who)
{
  if (format & 0x100) sb2h_fatal (who, "packed formats are outside the picture core (format 0x%x)", (unsigned) format);
  switch (SCHRO_FRAME_FORMAT_DEPTH (format)) {
    case SCHRO_FRAME_FORMAT_DEPTH_U8: return 0;
    case SCHRO_FRAME_FORMAT_DEPTH_S16: return 1;
    default: return 2;
  }
}

void
schro_frame_convert (SchroFrame *dest, SchroFrame *src)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s, d;
  const int sd = depth_code (src->format, __func__), dd = depth_code (dest->format, __func__);
  SB2H_ASSERT (dest != NULL && src != NULL);
  if ((dest->format & 3) != (src->format & 3))
    sb2h_fatal (__func__, "chroma resampling is outside the picture core (formats 0x%x -> 0x%x)",
        (unsigned) src->format, (unsigned) dest->format);
  stage_in (cx, &s, src, SB2H_BUF_IN, 1);
  stage_in (cx, &d, dest, SB2H_BUF_OUT, 1);      /* keep dest's borders as they are */
  SB2H_CHECK (sb2_frame_convert (&s.slab, sd, &d.slab, dd, cx->stream), "sb2_frame_convert");
  stage_out (cx, &d);
  {
    Staged *st[2] = { &s, &d };
    stage_finish (cx, st, 2, 2u);
  }
}

static void
frame_add_sub (SchroFrame *dest, SchroFrame *src, int subtract, const char *who)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s, d;
  const int sd = depth_code (src->format, who);
  SB2H_ASSERT (dest != NULL && src != NULL);
  /* the reference's tables (schroframe.c:984-1047): s16 += s16 | u8, same chroma format */
  if (depth_code (dest->format, who) != 1 || sd > 1 || (dest->format & 3) != (src->format & 3))
    sb2h_fatal (who, "%s function unimplemented (formats 0x%x, 0x%x)", subtract ? "subtract" : "add",
        (unsigned) dest->format, (unsigned) src->format);
  stage_in (cx, &s, src, SB2H_BUF_IN, 1);
  stage_in (cx, &d, dest, SB2H_BUF_OUT, 1);
  SB2H_CHECK (sb2_frame_add (&d.slab, &s.slab, sd, subtract, cx->stream), "sb2_frame_add");
  stage_out (cx, &d);
  {
    Staged *st[2] = { &s, &d };
    stage_finish (cx, st, 2, 2u);
  }
}

void
schro_frame_add (SchroFrame *dest, SchroFrame *src)
{
  frame_add_sub (dest, src, 0, __func__);
}

void
schro_frame_subtract (SchroFrame *dest, SchroFrame *src)
{
  frame_add_sub (dest, src, 1, __func__);
}

/* schro_frame_shift_left / _right (schroedinger/schroframe.c:1238-1291) */
static void
frame_shift (SchroFrame *frame, int shift, int right, const char *who)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s;
  const int d = depth_code (frame->format, who);
  if (d == 0 || (!right && d != 1)) sb2h_fatal (who, "unimplemented format 0x%x", (unsigned) frame->format);
  stage_in (cx, &s, frame, SB2H_BUF_IN, 1);
  SB2H_CHECK (sb2_frame_shift (&s.slab, d, shift, right, cx->stream), "sb2_frame_shift");
  stage_out (cx, &s);
  {
    Staged *st[1] = { &s };
    stage_finish (cx, st, 1, 1u);
  }
}

void
schro_frame_shift_left (SchroFrame *frame, int shift)
{
  frame_shift (frame, shift, 0, __func__);
}

void
schro_frame_shift_right (SchroFrame *frame, int shift)
{
  frame_shift (frame, shift, 1, __func__);
}

/* schro_frame_md5 (schroedinger/schroframe.c:1817-1861): the MD5 block function (RFC 1321) chained
 * over every row of every component in 64-byte blocks, a row's tail zero-padded to a block; no length
 * padding, the four state words are the result.  The hash is a serial chain: frames on the device are
 * brought to the host and hashed there. */
static void
md5_block (uint32_t *st, const uint8_t *p)
{
  static const uint8_t rot[64] = { 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9,
    14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15,
    21, 6, 10, 15, 21 };
  static uint32_t K[64];
  static int have_k = 0;
  uint32_t m[16], a = st[0], b = st[1], c = st[2], d = st[3];
  int i;
  if (!have_k) {
    /* floor (2^32 * |sin (i + 1)|), RFC 1321 3.4 -- integers, computed once with a table-free recurrence
     * would need libm; the constants are public and fixed, so they are listed */
    static const uint32_t k[64] = {
      0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
      0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
      0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
      0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
      0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
      0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
      0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
      0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
    memcpy (K, k, sizeof (K));
    have_k = 1;
  }
  for (i = 0; i < 16; i++)
    m[i] = (uint32_t) p[4 * i] | ((uint32_t) p[4 * i + 1] << 8) | ((uint32_t) p[4 * i + 2] << 16) | ((uint32_t) p[4 * i + 3] << 24);
  for (i = 0; i < 64; i++) {
    uint32_t f, t;
    int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    t = a + f + K[i] + m[g];
    a = d; d = c; c = b;
    b = b + ((t << rot[i]) | (t >> (32 - rot[i])));
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

void
schro_frame_md5 (SchroFrame *frame, uint32_t *state)
{
  Sb2hContext *cx = sb2h_context ();
  const size_t bytes = frame_region_bytes (frame);
  uint8_t *host = frame->regions[0], *tmp = NULL;
  int k, y, x;
  if (sb2h_mem_kind (frame->regions[0]) == SB2H_MEM_DEVICE) {
    sb2h_frame_use (cx, frame->regions[0]);
    tmp = sb2h_pinned_pool_alloc (bytes);
    SB2H_CUDA (cudaMemcpyAsync (tmp, frame->regions[0], bytes, cudaMemcpyDefault, cx->stream));
    sb2h_sync (cx);
    host = tmp;
  }
  state[0] = 0x67452301;
  state[1] = 0xefcdab89;
  state[2] = 0x98badcfe;
  state[3] = 0x10325476;
  for (k = 0; k < 3; k++) {
    const SchroFrameData *c = &frame->components[k];
    /* the reference walks `width` BYTES of every line whatever the sample size (schroframe.c:1833-1846) */
    for (y = 0; y < c->height; y++) {
      const uint8_t *line = host + ((const uint8_t *) c->data - (const uint8_t *) frame->regions[0]) + (size_t) c->stride * y;
      for (x = 0; x + 63 < c->width; x += 64) md5_block (state, line + x);
      if (x < c->width) {
        uint8_t pad[64];
        memcpy (pad, line + x, (size_t) (c->width - x));
        memset (pad + (c->width - x), 0, (size_t) (64 - (c->width - x)));
        md5_block (state, pad);
      }
    }
  }
  if (tmp) sb2h_pinned_pool_free (tmp);
}

/* ---- low-delay slices -------------------------------------------------------
 * schro_decoder_decode_lowdelay_transform_data (schroedinger/schrolowdelay.c:745-761) takes a
 * SchroPicture; it reads three things from it: params, lowdelay_buffer->data and transform_frame
 * (compat/schro_lowdelay.c is the reference-side half).  The quantiser tables are the Dirac
 * specification's (quant factor 4 * 2^(q/4) in its integer form, offsets (f + 1) / 2), the values of
 * schro_table_quant / schro_table_offset_1_2 (schroedinger/schrotables.c); tests/test_lowdelay_gpu.py checks
 * them against the reference's tables. */
static void
lowdelay_tables (uint32_t *quant, uint32_t *offset)
{
  int q;
  for (q = 0; q < 61; q++) {
    const uint64_t base = (uint64_t) 1 << (q / 4);
    uint64_t f;
    switch (q & 3) {
      case 0: f = 4 * base; break;
      case 1: f = (503829 * base + 52958) / 105917; break;
      case 2: f = (665857 * base + 58854) / 117708; break;
      default: f = (440253 * base + 32722) / 65444; break;
    }
    quant[q] = (uint32_t) f;
    offset[q] = q == 0 ? 1 : (q == 1 ? 2 : (uint32_t) ((f + 1) / 2));
  }
}

void
schro_b200_decode_lowdelay_transform_data (SchroParams *params, const uint8_t *data, int length, SchroFrame *transform_frame)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s;
  sb2_lowdelay_params p;
  const int d = depth_code (transform_frame->format, __func__);
  const int covered = !((params->iwt_luma_width | params->iwt_luma_height | params->iwt_chroma_width | params->iwt_chroma_height)
      & ((1 << params->transform_depth) - 1));
  void *dev_data;
  int i;
  SB2H_ASSERT (params && data && length > 0 && transform_frame);
  if (d == 0) sb2h_fatal (__func__, "the transform frame must be s16 or s32 (format 0x%x)", (unsigned) transform_frame->format);
  /* the reference's own preconditions (schrolowdelay.c:579-582) */
  SB2H_ASSERT ((params->iwt_luma_width % params->n_horiz_slices) == 0 && (params->iwt_luma_height % params->n_vert_slices) == 0);
  SB2H_ASSERT ((params->iwt_chroma_width % params->n_horiz_slices) == 0 && (params->iwt_chroma_height % params->n_vert_slices) == 0);
  memset (&p, 0, sizeof (p));
  p.transform_depth = params->transform_depth;
  p.n_horiz_slices = params->n_horiz_slices;
  p.n_vert_slices = params->n_vert_slices;
  p.slice_bytes_num = params->slice_bytes_num;
  p.slice_bytes_denom = params->slice_bytes_denom;
  for (i = 0; i < 1 + 3 * params->transform_depth; i++) p.quant_matrix[i] = params->quant_matrix[i];
  lowdelay_tables (p.table_quant, p.table_offset);
  /* the thread's own staging buffer, not the shared block pool: a pool block freed with this call's work still
   * in flight would make the next taker's stream wait for it, chaining the workers' decodes one behind another
   * (measured: 3.6 ms per call with 8 workers against 0.5 ms of GPU work) */
  dev_data = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, (size_t) length + 16);
  SB2H_CUDA (cudaMemcpyAsync (dev_data, data, (size_t) length, cudaMemcpyDefault, cx->stream));
  /* every sample of the frame is written when the sizes are multiples of 1 << depth: a host frame then needs no upload */
  stage_in (cx, &s, transform_frame, SB2H_BUF_IN, !covered);
  /* the frame's components may be larger than the iwt area (never smaller): describe the iwt area */
  s.slab.width[0] = params->iwt_luma_width;
  s.slab.height[0] = params->iwt_luma_height;
  s.slab.width[1] = s.slab.width[2] = params->iwt_chroma_width;
  s.slab.height[1] = s.slab.height[2] = params->iwt_chroma_height;
  SB2H_CHECK (sb2_lowdelay_decode (&p, dev_data, (size_t) length, (size_t) length, &s.slab, d == 2, cx->stream), "sb2_lowdelay_decode");
  stage_out (cx, &s);
  {
    Staged *st[1] = { &s };
    stage_finish (cx, st, 1, 1u);
  }
  /* the slice bytes may be pageable or page-locked host memory the caller reuses: wait for the upload */
  if (sb2h_mem_kind (data) != SB2H_MEM_DEVICE) sb2h_sync (cx);
}

/* A batch of low-delay intra pictures in one go: what a caller that holds several independent pictures (an
 * intra-only stream: every picture is its own GOP) should use instead of one call chain per picture --
 * the slices of all pictures are uploaded, then ONE slice-decode launch, ONE inverse transform with the
 * combine fused into its last level, and the 8-bit pictures come back; one wait for the whole batch.
 * Equivalent to, for every i: schro_b200_decode_lowdelay_transform_data (params, data[i], length, T),
 * schro_b200_frame_inverse_iwt_combine (outputs[i], T, params, shift).  All pictures share `params`;
 * data[i] point to `length` bytes each (pinned memory for full-rate DMA); outputs[i] are u8 host frames of
 * one layout.  Falls back to the per-picture calls when the fused kernel does not cover the shape. */
void
schro_b200_decode_lowdelay_pictures (SchroParams *params, int n, const uint8_t *const *data, int length,
    SchroFrame *const *outputs, int is_s32, int shift)
{
  Sb2hContext *cx = sb2h_context ();
  sb2_lowdelay_params p;
  sb2_slab coef, out;
  const int bpp = is_s32 ? 4 : 2;
  const size_t pitch = ((size_t) length + 255) & ~(size_t) 255;
  size_t coef_pic = 0, out_pic = 0, ws_bytes;
  char *dev_data, *dev_coef, *dev_out;
  void *ws;
  int i, k, rc;
  SB2H_ASSERT (params && n > 0 && data && outputs && length > 0);
  memset (&p, 0, sizeof (p));
  p.transform_depth = params->transform_depth;
  p.n_horiz_slices = params->n_horiz_slices;
  p.n_vert_slices = params->n_vert_slices;
  p.slice_bytes_num = params->slice_bytes_num;
  p.slice_bytes_denom = params->slice_bytes_denom;
  for (i = 0; i < 1 + 3 * params->transform_depth; i++) p.quant_matrix[i] = params->quant_matrix[i];
  lowdelay_tables (p.table_quant, p.table_offset);
  /* device slabs: dense coefficient planes (stride = width rounded up to 16 bytes), the outputs' own layout */
  memset (&coef, 0, sizeof (coef));
  memset (&out, 0, sizeof (out));
  coef.ncomp = out.ncomp = 3;
  coef.count = out.count = n;
  for (k = 0; k < 3; k++) {
    const SchroFrameData *oc = &outputs[0]->components[k];
    coef.width[k] = k ? params->iwt_chroma_width : params->iwt_luma_width;
    coef.height[k] = k ? params->iwt_chroma_height : params->iwt_luma_height;
    coef.stride[k] = (coef.width[k] * bpp + 15) & ~15;
    coef.offset[k] = coef_pic;
    coef_pic += ((size_t) coef.stride[k] * coef.height[k] + 255) & ~(size_t) 255;
    out.width[k] = oc->width < coef.width[k] ? oc->width : coef.width[k];
    out.height[k] = oc->height < coef.height[k] ? oc->height : coef.height[k];
    out.stride[k] = oc->stride;
    out.offset[k] = (size_t) ((char *) oc->data - (char *) outputs[0]->regions[0]);
    out_pic += (size_t) oc->length;
  }
  out_pic = (out_pic + 255) & ~(size_t) 255;
  coef.picture_pitch = coef_pic;
  out.picture_pitch = out_pic;
  dev_data = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, pitch * n + 256);
  dev_coef = sb2h_dev_buffer (cx, SB2H_BUF_IN, coef_pic * n + 256);
  dev_out = sb2h_dev_buffer (cx, SB2H_BUF_OUT, out_pic * n + 256);
  coef.base = dev_coef;
  out.base = dev_out;
  for (i = 0; i < n; i++)
    SB2H_CUDA (cudaMemcpyAsync (dev_data + pitch * i, data[i], (size_t) length, cudaMemcpyDefault, cx->stream));
  if ((params->iwt_luma_width | params->iwt_luma_height | params->iwt_chroma_width | params->iwt_chroma_height)
      & ((1 << params->transform_depth) - 1))
    SB2H_CUDA (cudaMemsetAsync (dev_coef, 0, coef_pic * n, cx->stream));       /* samples no subband covers */
  SB2H_CHECK (sb2_lowdelay_decode (&p, (const uint8_t *) dev_data, (size_t) length, pitch, &coef, is_s32, cx->stream),
      "sb2_lowdelay_decode");
  ws_bytes = sb2_iwt_workspace_bytes (&coef, is_s32, params->transform_depth, 0);
  ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
  rc = sb2_iwt_inverse_convert (&coef, &out, is_s32, params->wavelet_filter_index, params->transform_depth, shift, ws,
      ws_bytes, cx->stream);
  if (rc == SB2_ERR_UNSUPPORTED) {
    /* shapes the fused kernel does not cover: in place, then shift and convert */
    ws_bytes = sb2_iwt_workspace_bytes (&coef, is_s32, params->transform_depth, 1);
    ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
    SB2H_CHECK (sb2_iwt_inverse (&coef, &coef, is_s32, params->wavelet_filter_index, params->transform_depth, ws, ws_bytes,
            cx->stream), "sb2_iwt_inverse");
    if (shift) SB2H_CHECK (sb2_frame_shift (&coef, is_s32 ? 2 : 1, shift, 1, cx->stream), "sb2_frame_shift");
    SB2H_CHECK (sb2_frame_convert (&coef, is_s32 ? 2 : 1, &out, 0, cx->stream), "sb2_frame_convert");
  } else {
    SB2H_CHECK (rc, "sb2_iwt_inverse_convert");
  }
  for (i = 0; i < n; i++) {
    if (outputs[i]->extension == 0) {
      /* no border: the region is the three planes back to back -- one DMA per picture (the few bytes of row
       * padding beyond a plane's width carry no meaning) */
      SB2H_CUDA (cudaMemcpyAsync (outputs[i]->regions[0], dev_out + out_pic * i, frame_region_bytes (outputs[i]),
              cudaMemcpyDefault, cx->stream));
      continue;
    }
    /* frames with a border: plane by plane, so that the border stays as it is */
    for (k = 0; k < 3; k++)
      sb2h_copy_rect (cx, (char *) outputs[i]->regions[0] + out.offset[k], (size_t) out.stride[k],
          dev_out + out_pic * i + out.offset[k], (size_t) out.stride[k], (size_t) out.width[k], out.height[k]);
  }
  cx->dirty = 1;
  sb2h_sync (cx);
}

/* ---- OBMC ------------------------------------------------------------------ */
SchroMotion *
schro_motion_new (SchroParams *params, SchroFrame *ref1, SchroFrame *ref2)
{
  SchroMotion *motion = calloc (1, sizeof (SchroMotion));
  motion->params = params;
  motion->src1 = ref1;
  motion->src2 = ref2;
  motion->motion_vectors = calloc ((size_t) params->x_num_blocks * params->y_num_blocks,
      sizeof (SchroMotionVector));
  return motion;
}

void
schro_motion_free (SchroMotion *motion)
{
  free (motion->motion_vectors);
  free (motion);
}

static int
ramp (int x, int offset)
{
  if (offset == 1) return x == 0 ? 3 : 5;
  return 1 + (6 * x + offset - 1) / (2 * offset - 1);
}

/* Host-side table only (the kernels build their own copy); kept because it is public API. */
void
schro_motion_init_obmc_weight (SchroMotion *motion)
{
  int i, j;
  for (i = 0; i < motion->xblen; i++) {
    int w;
    if (motion->xoffset == 0) w = 8;
    else if (i < 2 * motion->xoffset) w = ramp (i, motion->xoffset);
    else if (motion->xblen - 1 - i < 2 * motion->xoffset) w = ramp (motion->xblen - 1 - i, motion->xoffset);
    else w = 8;
    motion->weight_x[i] = w;
  }
  for (j = 0; j < motion->yblen; j++) {
    int w;
    if (motion->yoffset == 0) w = 8;
    else if (j < 2 * motion->yoffset) w = ramp (j, motion->yoffset);
    else if (motion->yblen - 1 - j < 2 * motion->yoffset) w = ramp (motion->yblen - 1 - j, motion->yoffset);
    else w = 8;
    motion->weight_y[j] = w;
  }
  if (motion->obmc_weight.data) {
    for (j = 0; j < motion->yblen; j++) {
      int16_t *row = (int16_t *) ((char *) motion->obmc_weight.data + (size_t) motion->obmc_weight.stride * j);
      for (i = 0; i < motion->xblen; i++) row[i] = (int16_t) (motion->weight_x[i] * motion->weight_y[j]);
    }
  }
}

static void
motion_render (SchroMotion *motion, SchroFrame *dest, SchroFrame *addframe, int add,
    SchroFrame *output_frame, int use_ref_renderer)
{
  Sb2hContext *cx = sb2h_context ();
  SchroParams *params = motion->params;
  Staged r0, r1, acc, res, out;
  sb2_obmc_params p;
  const size_t nmv = (size_t) params->x_num_blocks * params->y_num_blocks;
  void *dmv;
  int res_is_s32;

  if (params->num_refs == 1) SB2H_ASSERT (params->picture_weight_2 == 1);   /* schromotion8.c:711 */
  if (add && !output_frame) sb2h_fatal (__func__, "add needs an output frame");
  require_u8 (motion->src1, __func__);
  if (motion->src1->extension < 32 || !motion->src1->is_upsampled)
    sb2h_fatal (__func__, "reference frames must be upsampled with extension >= 32");
  res_is_s32 = SCHRO_FRAME_FORMAT_DEPTH (addframe->format) == SCHRO_FRAME_FORMAT_DEPTH_S32;

  memset (&p, 0, sizeof (p));
  p.xbsep = params->xbsep_luma;
  p.ybsep = params->ybsep_luma;
  p.xblen = params->xblen_luma;
  p.yblen = params->yblen_luma;
  p.x_num_blocks = params->x_num_blocks;
  p.y_num_blocks = params->y_num_blocks;
  p.mv_precision = params->mv_precision;
  p.picture_weight_1 = params->picture_weight_1;
  p.picture_weight_2 = params->picture_weight_2;
  p.picture_weight_bits = params->picture_weight_bits;
  p.chroma_h_shift = SCHRO_CHROMA_FORMAT_H_SHIFT (params->video_format->chroma_format);
  p.chroma_v_shift = SCHRO_CHROMA_FORMAT_V_SHIFT (params->video_format->chroma_format);

  stage_in (cx, &r0, motion->src1, SB2H_BUF_IN, 1);
  if (motion->src2) stage_in (cx, &r1, motion->src2, SB2H_BUF_OUT, 1);
  stage_in (cx, &acc, dest, SB2H_BUF_AUX0, 1);     /* keeps padding / borders of dest intact */
  stage_in (cx, &res, addframe, SB2H_BUF_AUX1, 1);
  if (output_frame) stage_in (cx, &out, output_frame, SB2H_BUF_AUX2, 1);
  dmv = sb2h_dev_buffer (cx, SB2H_BUF_AUX3, nmv * sizeof (SchroMotionVector));
  /* the vectors come from plain malloc'd memory (schro_motion_new): staged through the thread's page-locked block */
  sb2h_upload_staged (cx, dmv, motion->motion_vectors, nmv * sizeof (SchroMotionVector));
  /* the rendered area is dest's (schromotion8.c:722-751); addframe may be iwt-padded */
  if (use_ref_renderer) {
    /* global motion: the reference's per-pixel renderer (schromotion.c:113-121, schromotionref.c:245-330) */
    int gm[20], r;
    if (res_is_s32) sb2h_fatal (__func__, "the per-pixel renderer adds s16 residuals only (schromotionref.c:252)");
    for (r = 0; r < 2; r++) {
      const SchroGlobalMotion *g = &params->global_motion[r];
      const int v[10] = { g->b0, g->b1, g->a_exp, g->a00, g->a01, g->a10, g->a11, g->c_exp, g->c0, g->c1 };
      memcpy (gm + 10 * r, v, sizeof (v));
    }
    SB2H_CHECK (sb2_obmc_render_ref (&p, gm, dmv, nmv, &r0.slab, motion->src2 ? &r1.slab : NULL, &acc.slab,
            &res.slab, add, output_frame ? &out.slab : NULL, cx->stream), "sb2_obmc_render_ref");
  } else {
    SB2H_CHECK (sb2_obmc_render (&p, dmv, nmv, &r0.slab, motion->src2 ? &r1.slab : NULL, &acc.slab,
            &res.slab, res_is_s32, add, output_frame ? &out.slab : NULL, cx->stream), "sb2_obmc_render");
  }
  stage_out (cx, &acc);
  if (add) stage_out (cx, &out);
  else stage_out (cx, &res);
  if (sb2h_mem_kind (motion->motion_vectors) != SB2H_MEM_PAGEABLE) {
    /* vectors in page-locked memory are read by the DMA engine after the copy call returns */
    sb2h_sync (cx);
  } else {
    Staged *st[5] = { &r0, motion->src2 ? &r1 : NULL, &acc, &res, output_frame ? &out : NULL };
    stage_finish (cx, st, 5, 4u | (add ? 16u : 8u));
  }
}

void
schro_motion_render_u8 (SchroMotion *motion, SchroFrame *dest, SchroFrame *addframe, int add,
    SchroFrame *output_frame)
{
  motion_render (motion, dest, addframe, add, output_frame, 0);
}

/* schroedinger/schromotionref.c:245 */
void
schro_motion_render_ref (SchroMotion *motion, SchroFrame *dest, SchroFrame *addframe, int add,
    SchroFrame *output_frame)
{
  motion_render (motion, dest, addframe, add, output_frame, 1);
}

/* the dispatcher (schroedinger/schromotion.c:95-155): global motion takes the per-pixel renderer */
void
schro_motion_render (SchroMotion *motion, SchroFrame *dest, SchroFrame *addframe, int add,
    SchroFrame *output_frame)
{
  motion_render (motion, dest, addframe, add, output_frame, motion->params->have_global_motion ? 1 : 0);
}

/* ---- SAD primitives ---------------------------------------------------------- */
static const uint8_t *
block_to_device (Sb2hContext *cx, int which, const uint8_t *p, int stride, int w, int h, int *dstride)
{
  if (sb2h_mem_kind (p) == SB2H_MEM_DEVICE) {
    sb2h_ptr_use (cx, p);          /* may still be written by another thread's stream-ordered call */
    *dstride = stride;
    return p;
  }
  {
    const size_t pitch = ((size_t) w + 15) & ~(size_t) 15;
    uint8_t *d = sb2h_dev_buffer (cx, which, pitch * h + 256);
    sb2h_copy_rect (cx, d, pitch, p, (size_t) stride, (size_t) w, h);
    *dstride = (int) pitch;
    return d;
  }
}

int
schro_metric_absdiff_u8 (uint8_t *a, int a_stride, uint8_t *b, int b_stride, int width, int height)
{
  Sb2hContext *cx = sb2h_context ();
  int as, bs;
  const uint8_t *da = block_to_device (cx, SB2H_BUF_IN, a, a_stride, width, height, &as);
  const uint8_t *db = block_to_device (cx, SB2H_BUF_OUT, b, b_stride, width, height, &bs);
  int64_t *off = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, 64);
  uint32_t result = 0;
  SB2H_CUDA (cudaMemsetAsync (off, 0, 64, cx->stream));
  SB2H_CHECK (sb2_sad_u8 (da, as, db, bs, off, off + 1, 1, width, height, (uint32_t *) (off + 2),
          cx->stream), "sb2_sad_u8");
  SB2H_CUDA (cudaMemcpyAsync (&result, off + 2, sizeof (result), cudaMemcpyDefault, cx->stream));
  sb2h_sync (cx);
  return (int) result;
}

int
schro_metric_get (SchroFrameData *src1, SchroFrameData *src2, int width, int height)
{
  return schro_metric_absdiff_u8 (src1->data, src1->stride, src2->data, src2->stride, width, height);
}

int
schro_metric_get_dc (SchroFrameData *src, int value, int width, int height)
{
  Sb2hContext *cx = sb2h_context ();
  int as, result = 0;
  const uint8_t *da;
  int *dres;
  SB2H_ASSERT (src->width >= width && src->height >= height);
  da = block_to_device (cx, SB2H_BUF_IN, src->data, src->stride, width, height, &as);
  dres = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, 64);
  SB2H_CHECK (sb2_sad_dc_u8 (da, as, value, width, height, dres, cx->stream), "sb2_sad_dc_u8");
  SB2H_CUDA (cudaMemcpyAsync (&result, dres, sizeof (result), cudaMemcpyDefault, cx->stream));
  sb2h_sync (cx);
  return result;
}

int
schro_metric_get_biref (SchroFrameData *fd, SchroFrameData *src1, int weight1, SchroFrameData *src2,
    int weight2, int shift, int width, int height)
{
  Sb2hContext *cx = sb2h_context ();
  int as, s1s, s2s, result = 0;
  const uint8_t *da = block_to_device (cx, SB2H_BUF_IN, fd->data, fd->stride, width, height, &as);
  const uint8_t *d1 = block_to_device (cx, SB2H_BUF_OUT, src1->data, src1->stride, width, height, &s1s);
  const uint8_t *d2 = block_to_device (cx, SB2H_BUF_AUX1, src2->data, src2->stride, width, height, &s2s);
  int *dres = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, 64);
  SB2H_CHECK (sb2_sad_biref_u8 (da, as, d1, s1s, weight1, d2, s2s, weight2, shift, width, height, dres,
          cx->stream), "sb2_sad_biref_u8");
  SB2H_CUDA (cudaMemcpyAsync (&result, dres, sizeof (result), cudaMemcpyDefault, cx->stream));
  sb2h_sync (cx);
  return result;
}

/* ---- the scan of one block and the 3-component block SAD (schrometric.c:31-214, 332-414) ---- */
void
schro_metric_scan_setup (SchroMetricScan *scan, int dx, int dy, int dist, int use_chroma)
{
  int xmin, xmax, ymin, ymax;
  SB2H_ASSERT (scan && scan->frame && scan->ref_frame && dist > 0);
  /* the window is clipped to the block-sized margin around the picture and to the frame's border */
  xmin = scan->x + dx - dist;
  xmax = scan->x + dx + dist;
  ymin = scan->y + dy - dist;
  ymax = scan->y + dy + dist;
  if (xmin < -scan->block_width) xmin = -scan->block_width;
  if (ymin < -scan->block_height) ymin = -scan->block_height;
  if (xmax > scan->frame->width) xmax = scan->frame->width;
  if (ymax > scan->frame->height) ymax = scan->frame->height;
  if (xmin < -scan->frame->extension) xmin = -scan->frame->extension;
  if (ymin < -scan->frame->extension) ymin = -scan->frame->extension;
  if (xmax > scan->frame->width - scan->block_width + scan->frame->extension)
    xmax = scan->frame->width - scan->block_width + scan->frame->extension;
  if (ymax > scan->frame->height - scan->block_height + scan->frame->extension)
    ymax = scan->frame->height - scan->block_height + scan->frame->extension;
  scan->ref_x = xmin;
  scan->ref_y = ymin;
  scan->scan_width = xmax - xmin + 1;
  scan->scan_height = ymax - ymin + 1;
  scan->use_chroma = use_chroma;
  SB2H_ASSERT (scan->scan_width <= SCHRO_LIMIT_METRIC_SCAN && scan->scan_height <= SCHRO_LIMIT_METRIC_SCAN);
}

void
schro_metric_scan_do_scan (SchroMetricScan *scan)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s, r;
  sb2_metric_scan_desc d;
  char *dbuf;
  const size_t grid = sizeof (uint32_t) * SCHRO_LIMIT_METRIC_SCAN * SCHRO_LIMIT_METRIC_SCAN;
  const size_t used = sizeof (uint32_t) * (size_t) scan->scan_width * scan->scan_height;
  const int hs = SCHRO_FRAME_FORMAT_H_SHIFT (scan->frame->format), vs = SCHRO_FRAME_FORMAT_V_SHIFT (scan->frame->format);
  /* the reference's own preconditions (schrometric.c:38-45) */
  SB2H_ASSERT (scan->scan_width > 0 && scan->scan_height > 0);
  SB2H_ASSERT (scan->ref_x >= -scan->frame->extension && scan->ref_y >= -scan->frame->extension);
  SB2H_ASSERT (scan->ref_x + scan->block_width + scan->scan_width - 1 <= scan->frame->width + scan->frame->extension);
  SB2H_ASSERT (scan->ref_y + scan->block_height + scan->scan_height - 1 <= scan->frame->height + scan->frame->extension);
  if (scan->use_chroma && !(hs == 1 && vs == 1))
    sb2h_fatal (__func__, "chroma scans are defined for 4:2:0 only (the reference's duplication scheme, schrometric.c:73-115)");
  require_u8 (scan->frame, __func__);
  require_u8 (scan->ref_frame, __func__);
  stage_in (cx, &s, scan->frame, SB2H_BUF_IN, 1);
  stage_in (cx, &r, scan->ref_frame, SB2H_BUF_OUT, 1);
  d.picture = 0;
  d.x = scan->x;
  d.y = scan->y;
  d.block_width = scan->block_width;
  d.block_height = scan->block_height;
  d.ref_x = scan->ref_x;
  d.ref_y = scan->ref_y;
  d.scan_width = scan->scan_width;
  d.scan_height = scan->scan_height;
  dbuf = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, 256 + 2 * grid);
  SB2H_CUDA (cudaMemcpyAsync (dbuf, &d, sizeof (d), cudaMemcpyDefault, cx->stream));
  SB2H_CHECK (sb2_metric_scan (&s.slab, &r.slab, hs, vs, scan->use_chroma, (const sb2_metric_scan_desc *) dbuf, 1,
          (uint32_t *) (dbuf + 256), (uint32_t *) (dbuf + 256 + grid), cx->stream), "sb2_metric_scan");
  memset (scan->chroma_metrics, 0, sizeof (scan->chroma_metrics));
  SB2H_CUDA (cudaMemcpyAsync (scan->metrics, dbuf + 256, used, cudaMemcpyDefault, cx->stream));
  if (scan->use_chroma)
    SB2H_CUDA (cudaMemcpyAsync (scan->chroma_metrics, dbuf + 256 + grid, used, cudaMemcpyDefault, cx->stream));
  sb2h_sync (cx);
}

int
schro_metric_scan_get_min (SchroMetricScan *scan, int *dx, int *dy, uint32_t *chroma_error)
{
  /* the seed (gravity) position starts as the best, a later position replaces it only when strictly
   * better, positions are visited column by column */
  const int h = scan->scan_height;
  int i = scan->gravity_x + scan->x - scan->ref_x, j = scan->gravity_y + scan->y - scan->ref_y;
  uint32_t best = scan->metrics[j + i * h], best_chroma = 0, best_total = 0;
  SB2H_ASSERT (scan->scan_width > 0 && scan->scan_height > 0);
  if (scan->use_chroma) {
    best_chroma = scan->chroma_metrics[j + i * h];
    best_total = best + best_chroma;
  }
  for (i = 0; i < scan->scan_width; i++)
    for (j = 0; j < h; j++) {
      const uint32_t m = scan->metrics[i * h + j], c = scan->chroma_metrics[i * h + j];
      const int better = scan->use_chroma ? (m + c < best_total) : (m < best);
      if (better) {
        best = m;
        best_chroma = c;
        best_total = m + c;
        *dx = scan->ref_x + i - scan->x;
        *dy = scan->ref_y + j - scan->y;
      }
    }
  *chroma_error = best_chroma;
  return (int) best;
}

static int
metric_block_sad3 (SchroMetricInfo *info, int x, int y, int dx, int dy)
{
  Sb2hContext *cx = sb2h_context ();
  Staged s, r;
  sb2_metric_block_desc d;
  char *dbuf;
  int result = 0;
  require_u8 (info->frame, __func__);
  require_u8 (info->ref_frame, __func__);
  stage_in (cx, &s, info->frame, SB2H_BUF_IN, 1);
  stage_in (cx, &r, info->ref_frame, SB2H_BUF_OUT, 1);
  d.picture = 0;
  d.x = x;
  d.y = y;
  d.dx = dx;
  d.dy = dy;
  dbuf = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, 512);
  SB2H_CUDA (cudaMemcpyAsync (dbuf, &d, sizeof (d), cudaMemcpyDefault, cx->stream));
  SB2H_CHECK (sb2_metric_block_sad3 (&s.slab, &r.slab, info->frame->extension, info->block_width[0], info->block_height[0],
          info->h_shift[1], info->v_shift[1], (const sb2_metric_block_desc *) dbuf, 1, (int *) (dbuf + 256), cx->stream),
      "sb2_metric_block_sad3");
  SB2H_CUDA (cudaMemcpyAsync (&result, dbuf + 256, sizeof (result), cudaMemcpyDefault, cx->stream));
  sb2h_sync (cx);
  return result;
}

void
schro_metric_info_init (SchroMetricInfo *info, SchroFrame *frame, SchroFrame *ref_frame, int block_width, int block_height)
{
  int k;
  memset (info, 0, sizeof (*info));
  info->frame = frame;
  info->ref_frame = ref_frame;
  for (k = 0; k < 3; k++) {
    info->h_shift[k] = k ? SCHRO_FRAME_FORMAT_H_SHIFT (frame->format) : 0;
    info->v_shift[k] = k ? SCHRO_FRAME_FORMAT_V_SHIFT (frame->format) : 0;
    info->block_width[k] = block_width >> info->h_shift[k];
    info->block_height[k] = block_height >> info->v_shift[k];
  }
  info->metric = metric_block_sad3;
  info->metric_right = metric_block_sad3;
  info->metric_bottom = metric_block_sad3;
  info->metric_corner = metric_block_sad3;
}

int
schro_metric_fast_block (SchroMetricInfo *info, int x, int y, int dx, int dy)
{
  return info->metric (info, x, y, dx, dy);
}

/* ---- schro_frame_dup* (schroframe.c:699-724) -------------------------------------------- */
SchroFrame *
schro_frame_dup_full (SchroFrame *frame, int extension, int is_upsampled)
{
  SchroFrame *dup = schro_frame_new_and_alloc_full (frame->domain, frame->format, frame->width, frame->height,
      extension, is_upsampled);
  schro_frame_convert (dup, frame);
  return dup;
}

SchroFrame *
schro_frame_dup_extended (SchroFrame *frame, int extension)
{
  return schro_frame_dup_full (frame, extension, 0);
}

SchroFrame *
schro_frame_dup (SchroFrame *frame)
{
  return schro_frame_dup_extended (frame, 0);
}
