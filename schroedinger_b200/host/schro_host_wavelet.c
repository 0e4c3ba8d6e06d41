/*
 * schro_host_wavelet.c -- the reference's wavelet entry points on top of the CUDA layer.
 *
 *   schro_wavelet_transform_2d          schroedinger/schrowaveletorc.c:60-117
 *   schro_wavelet_inverse_transform_2d  schroedinger/schrowaveletorc.c:121-188
 *   schro_frame_iwt_transform           schroedinger/schroframe.c:1192-1228
 *   schro_frame_inverse_iwt_transform   schroedinger/schrodecoder.c:1809-1853
 *
 * Host buffers are staged H2D / D2H on the calling thread's stream; device
 * (CUDA-domain) buffers are transformed where they lie.  `tmp` is ignored (the
 * kernels keep their line buffers in shared memory).
 */
#include "schro_host.h"
#include <string.h>

typedef struct {
  void *data[3];
  int stride[3], width[3], height[3];
  int ncomp;
} PlaneList;

static size_t
dense_layout (const PlaneList *pl, int bpp, sb2_slab *slab)
{
  size_t pos = 0;
  int c;
  memset (slab, 0, sizeof (*slab));
  slab->count = 1;
  slab->ncomp = pl->ncomp;
  for (c = 0; c < pl->ncomp; c++) {
    slab->stride[c] = (pl->width[c] * bpp + 15) & ~15;
    slab->width[c] = pl->width[c];
    slab->height[c] = pl->height[c];
    slab->offset[c] = pos;
    pos += ((size_t) slab->stride[c] * pl->height[c] + 255) & ~(size_t) 255;
  }
  slab->picture_pitch = pos;
  return pos;
}

/* multi-level transform of up to three planes that share format and depth */
static void
run_planes (const PlaneList *src, const PlaneList *dst, int is_s32, int filter, int depth,
    int inverse)
{
  Sb2hContext *cx = sb2h_context ();
  const int bpp = is_s32 ? 4 : 2;
  /* zero-copy needs BOTH sides on the device; any other combination is staged (the rectangle
   * copies below take host or device memory on either side) */
  const int on_device = sb2h_mem_kind (src->data[0]) == SB2H_MEM_DEVICE && sb2h_mem_kind (dst->data[0]) == SB2H_MEM_DEVICE;
  sb2_slab sin, sout;
  size_t ws_bytes;
  void *ws;
  int c, rc;

  /* device planes may still be the target of another thread's stream-ordered call */
  for (c = 0; c < src->ncomp; c++) {
    if (sb2h_mem_kind (src->data[c]) == SB2H_MEM_DEVICE) sb2h_ptr_use (cx, src->data[c]);
    if (dst->data[c] != src->data[c] && sb2h_mem_kind (dst->data[c]) == SB2H_MEM_DEVICE) sb2h_ptr_use (cx, dst->data[c]);
  }
  if (on_device) {
    /* zero-copy: describe the planes relative to the lowest address */
    char *base = src->data[0], *dbase = dst->data[0];
    int in_place = 1;
    for (c = 0; c < src->ncomp; c++) {
      if ((char *) src->data[c] < base) base = src->data[c];
      if ((char *) dst->data[c] < dbase) dbase = dst->data[c];
      if (src->data[c] != dst->data[c]) in_place = 0;
    }
    memset (&sin, 0, sizeof (sin));
    sin.count = 1;
    sin.ncomp = src->ncomp;
    sout = sin;
    sin.base = base;
    sout.base = dbase;
    for (c = 0; c < src->ncomp; c++) {
      sin.offset[c] = (size_t) ((char *) src->data[c] - base);
      sin.stride[c] = src->stride[c];
      sin.width[c] = sout.width[c] = src->width[c];
      sin.height[c] = sout.height[c] = src->height[c];
      sout.offset[c] = (size_t) ((char *) dst->data[c] - dbase);
      sout.stride[c] = dst->stride[c];
    }
    ws_bytes = sb2_iwt_workspace_bytes (&sin, is_s32, depth, 1);
    ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
    (void) in_place;
    rc = inverse ? sb2_iwt_inverse (&sin, &sout, is_s32, filter, depth, ws, ws_bytes, cx->stream)
                 : sb2_iwt_forward (&sin, &sout, is_s32, filter, depth, ws, ws_bytes, cx->stream);
    SB2H_CHECK (rc, inverse ? "sb2_iwt_inverse" : "sb2_iwt_forward");
    for (c = 0; c < dst->ncomp; c++) sb2h_ptr_wrote (cx, dst->data[c]);
    sb2h_sync (cx);
    return;
  }

  /* host memory: H2D, transform out of place, D2H */
  {
    size_t bytes = dense_layout (src, bpp, &sin);
    sout = sin;
    sin.base = sb2h_dev_buffer (cx, SB2H_BUF_IN, bytes);
    sout.base = sb2h_dev_buffer (cx, SB2H_BUF_OUT, bytes);
    ws_bytes = sb2_iwt_workspace_bytes (&sin, is_s32, depth, 0);
    ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
    for (c = 0; c < src->ncomp; c++)
      sb2h_copy_rect (cx, (char *) sin.base + sin.offset[c], sin.stride[c], src->data[c],
          src->stride[c], (size_t) src->width[c] * bpp, src->height[c]);
    rc = inverse ? sb2_iwt_inverse (&sin, &sout, is_s32, filter, depth, ws, ws_bytes, cx->stream)
                 : sb2_iwt_forward (&sin, &sout, is_s32, filter, depth, ws, ws_bytes, cx->stream);
    SB2H_CHECK (rc, inverse ? "sb2_iwt_inverse" : "sb2_iwt_forward");
    for (c = 0; c < src->ncomp; c++)
      sb2h_copy_rect (cx, dst->data[c], dst->stride[c], (char *) sout.base + sout.offset[c],
          sout.stride[c], (size_t) src->width[c] * bpp, src->height[c]);
    sb2h_sync (cx);
  }
}

static int
depth_is_s32 (SchroFrameFormat format)
{
  int d = SCHRO_FRAME_FORMAT_DEPTH (format);
  if (d == SCHRO_FRAME_FORMAT_DEPTH_S16) return 0;
  if (d == SCHRO_FRAME_FORMAT_DEPTH_S32) return 1;
  sb2h_fatal (__func__, "wavelets need an s16 or s32 frame (format 0x%x)", (unsigned) format);
  return 0;
}

void
schro_wavelet_transform_2d (SchroFrameData *fd, int filter, int16_t *tmp)
{
  PlaneList pl;
  (void) tmp;
  if (filter < 0 || filter > 6) sb2h_fatal (__func__, "bad filter %d", filter);
  pl.ncomp = 1;
  pl.data[0] = fd->data;
  pl.stride[0] = fd->stride;
  pl.width[0] = fd->width;
  pl.height[0] = fd->height;
  run_planes (&pl, &pl, depth_is_s32 (fd->format), filter, 1, 0);
}

void
schro_wavelet_inverse_transform_2d (SchroFrameData *fd_dest, SchroFrameData *fd_src,
    int filter, int16_t *tmp)
{
  PlaneList ps, pd;
  (void) tmp;
  if (filter < 0 || filter > 6) sb2h_fatal (__func__, "bad filter %d", filter);
  SB2H_ASSERT (SCHRO_FRAME_FORMAT_DEPTH (fd_dest->format) == SCHRO_FRAME_FORMAT_DEPTH (fd_src->format));
  SB2H_ASSERT (fd_dest->width == fd_src->width && fd_dest->height == fd_src->height);
  ps.ncomp = pd.ncomp = 1;
  ps.data[0] = fd_src->data;  ps.stride[0] = fd_src->stride;
  ps.width[0] = fd_src->width; ps.height[0] = fd_src->height;
  pd = ps;
  pd.data[0] = fd_dest->data; pd.stride[0] = fd_dest->stride;
  /* The reference leaves the vertically un-lifted rows in src when dest != src
   * (schrowaveletorc.c:1483-1521) -- a side effect no caller relies on
   * (all callers pass dest == src, schrodecoder.c:1835-1848); here src is left intact. */
  run_planes (&ps, &pd, depth_is_s32 (fd_dest->format), filter, 1, 1);
}

/* CUDA-domain frame whose planes are exactly the transform area: transform out of place into
 * a fresh region from the same domain and swap the regions -- "in place" for the caller, no
 * staging copy and no copy back on the device. */
static int
frame_iwt_swap (SchroFrame *frame, SchroParams *params, int inverse)
{
  Sb2hContext *cx = sb2h_context ();
  const int is_s32 = SCHRO_FRAME_FORMAT_DEPTH (frame->format) == SCHRO_FRAME_FORMAT_DEPTH_S32;
  sb2_slab sin, sout;
  size_t total = 0, ws_bytes;
  char *old_region = frame->regions[0], *new_region;
  void *ws;
  int k, rc;

  if (!frame->domain || !old_region || frame->extension != 0 || frame->is_upsampled) return 0;
  if (sb2h_mem_kind (old_region) != SB2H_MEM_DEVICE) return 0;
  for (k = 0; k < 3; k++) {
    const SchroFrameData *c = &frame->components[k];
    if (c->width != (k ? params->iwt_chroma_width : params->iwt_luma_width) ||
        c->height != (k ? params->iwt_chroma_height : params->iwt_luma_height))
      return 0;
    total += (size_t) c->length;
  }
  memset (&sin, 0, sizeof (sin));
  sin.base = old_region;
  sin.picture_pitch = total;
  sin.count = 1;
  sin.ncomp = 3;
  for (k = 0; k < 3; k++) {
    sin.offset[k] = (size_t) ((char *) frame->components[k].data - old_region);
    sin.stride[k] = frame->components[k].stride;
    sin.width[k] = frame->components[k].width;
    sin.height[k] = frame->components[k].height;
  }
  sb2h_frame_use (cx, old_region);
  new_region = sb2h_spare_take (cx, frame->domain, (int) total);
  if (!new_region) new_region = schro_memory_domain_alloc (frame->domain, (int) total);
  sb2h_frame_use (cx, new_region);
  sout = sin;
  sout.base = new_region;
  ws_bytes = sb2_iwt_workspace_bytes (&sin, is_s32, params->transform_depth, 0);
  ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
  rc = inverse ? sb2_iwt_inverse (&sin, &sout, is_s32, params->wavelet_filter_index, params->transform_depth,
                     ws, ws_bytes, cx->stream)
               : sb2_iwt_forward (&sin, &sout, is_s32, params->wavelet_filter_index, params->transform_depth,
                     ws, ws_bytes, cx->stream);
  SB2H_CHECK (rc, inverse ? "sb2_iwt_inverse" : "sb2_iwt_forward");
  /* no wait: the new region carries the write, the old one is parked by the domain until the
   * transform (and anything else in flight) has finished reading it */
  sb2h_frame_wrote (cx, new_region);
  for (k = 0; k < 3; k++)
    frame->components[k].data = new_region + sin.offset[k];
  frame->regions[0] = new_region;
  /* the old region becomes this thread's spare for its next transform when only this thread ever used it;
   * otherwise the domain parks it until every thread's work on it has finished */
  if (!sb2h_spare_put (cx, frame->domain, old_region, (int) total))
    schro_memory_domain_memfree (frame->domain, old_region);
  return 1;
}

static void
frame_iwt (SchroFrame *frame, SchroParams *params, int inverse)
{
  PlaneList pl;
  int k;
  if (frame_iwt_swap (frame, params, inverse)) return;
  pl.ncomp = 3;
  for (k = 0; k < 3; k++) {
    pl.data[k] = frame->components[k].data;
    pl.stride[k] = frame->components[k].stride;
    pl.width[k] = k ? params->iwt_chroma_width : params->iwt_luma_width;
    pl.height[k] = k ? params->iwt_chroma_height : params->iwt_luma_height;
  }
  run_planes (&pl, &pl, depth_is_s32 (frame->format), params->wavelet_filter_index,
      params->transform_depth, inverse);
}

void
schro_frame_iwt_transform (SchroFrame *frame, SchroParams *params)
{
  frame_iwt (frame, params, 0);
}

void
schro_frame_inverse_iwt_transform (SchroFrame *frame, SchroParams *params)
{
  frame_iwt (frame, params, 1);
}

/* ---- inverse transform + combine in one pass -----------------------------
 * What the decoder does with a non-reference intra picture after the inverse transform
 * (schro_decoder_x_combine, schroedinger/schrodecoder.c:2054-2061): schro_frame_shift_right (frame, shift)
 * when the stream is deeper than the output, then schro_frame_convert (output, frame).  Here the three
 * steps are one call and -- for planes the register-chunk kernels cover -- one pass: the last wavelet level
 * writes the 8-bit picture directly.  `frame` (the coefficients) is left untouched; `output` is u8 of the
 * same chroma format, no larger than the transform's area.  Frames may be host or CUDA-domain memory. */
void
schro_b200_frame_inverse_iwt_combine (SchroFrame *output, SchroFrame *frame, SchroParams *params, int shift)
{
  Sb2hContext *cx = sb2h_context ();
  const int is_s32 = depth_is_s32 (frame->format);
  sb2_slab sin, sout;
  size_t tin = 0, tout = 0, ws_bytes;
  char *rin = frame->regions[0], *rout = output->regions[0], *din, *dout;
  void *ws;
  int k, rc, host_in, host_out;
  SB2H_ASSERT (output && frame && params && rin && rout);
  if (SCHRO_FRAME_FORMAT_DEPTH (output->format) != SCHRO_FRAME_FORMAT_DEPTH_U8 ||
      SCHRO_FRAME_FORMAT_DEPTH (frame->format) == SCHRO_FRAME_FORMAT_DEPTH_U8 || (output->format & 3) != (frame->format & 3))
    sb2h_fatal (__func__, "needs an s16 / s32 coefficient frame and a u8 output of the same chroma format (0x%x, 0x%x)",
        (unsigned) frame->format, (unsigned) output->format);
  for (k = 0; k < 3; k++) { tin += (size_t) frame->components[k].length; tout += (size_t) output->components[k].length; }
  host_in = sb2h_mem_kind (rin) != SB2H_MEM_DEVICE;
  host_out = sb2h_mem_kind (rout) != SB2H_MEM_DEVICE;
  if (host_in) {
    din = sb2h_dev_buffer (cx, SB2H_BUF_IN, tin + 256);
    SB2H_CUDA (cudaMemcpyAsync (din, rin, tin, cudaMemcpyDefault, cx->stream));
  } else {
    din = rin;
    sb2h_frame_use (cx, rin);
  }
  if (host_out) {
    dout = sb2h_dev_buffer (cx, SB2H_BUF_OUT, tout + 256);
    SB2H_CUDA (cudaMemcpyAsync (dout, rout, tout, cudaMemcpyDefault, cx->stream));     /* keeps borders / padding as they are */
  } else {
    dout = rout;
    sb2h_frame_use (cx, rout);
  }
  memset (&sin, 0, sizeof (sin));
  memset (&sout, 0, sizeof (sout));
  sin.base = din; sin.picture_pitch = tin; sin.count = 1; sin.ncomp = 3;
  sout.base = dout; sout.picture_pitch = tout; sout.count = 1; sout.ncomp = 3;
  for (k = 0; k < 3; k++) {
    sin.offset[k] = (size_t) ((char *) frame->components[k].data - rin);
    sin.stride[k] = frame->components[k].stride;
    sin.width[k] = k ? params->iwt_chroma_width : params->iwt_luma_width;
    sin.height[k] = k ? params->iwt_chroma_height : params->iwt_luma_height;
    sout.offset[k] = (size_t) ((char *) output->components[k].data - rout);
    sout.stride[k] = output->components[k].stride;
    sout.width[k] = output->components[k].width < sin.width[k] ? output->components[k].width : sin.width[k];
    sout.height[k] = output->components[k].height < sin.height[k] ? output->components[k].height : sin.height[k];
  }
  ws_bytes = sb2_iwt_workspace_bytes (&sin, is_s32, params->transform_depth, 0);
  ws = sb2h_dev_buffer (cx, SB2H_BUF_WS, ws_bytes);
  rc = sb2_iwt_inverse_convert (&sin, &sout, is_s32, params->wavelet_filter_index, params->transform_depth, shift, ws, ws_bytes,
      cx->stream);
  if (rc == SB2_ERR_UNSUPPORTED) {
    /* shapes the fused kernel does not cover: the three steps one after the other, into a scratch plane set */
    sb2_slab stmp = sin;
    const size_t full = sb2_iwt_workspace_bytes (&sin, is_s32, params->transform_depth, 0);
    char *tmp = sb2h_dev_buffer (cx, SB2H_BUF_AUX1, tin + 256);
    (void) full;
    stmp.base = tmp;
    SB2H_CHECK (sb2_iwt_inverse (&sin, &stmp, is_s32, params->wavelet_filter_index, params->transform_depth, ws, ws_bytes,
            cx->stream), "sb2_iwt_inverse");
    if (shift) SB2H_CHECK (sb2_frame_shift (&stmp, is_s32 ? 2 : 1, shift, 1, cx->stream), "sb2_frame_shift");
    SB2H_CHECK (sb2_frame_convert (&stmp, is_s32 ? 2 : 1, &sout, 0, cx->stream), "sb2_frame_convert");
  } else {
    SB2H_CHECK (rc, "sb2_iwt_inverse_convert");
  }
  if (host_out) {
    SB2H_CUDA (cudaMemcpyAsync (rout, dout, tout, cudaMemcpyDefault, cx->stream));
    sb2h_sync (cx);
  } else {
    sb2h_frame_wrote (cx, rout);
    cx->dirty = 1;
    if (host_in) sb2h_sync (cx);          /* the caller may reuse its host coefficients */
  }
}

/* ---- dequantisation on the device -------------------------------
 * New entry point (the reference dequantises inside its entropy decoder, codeblock by codeblock:
 * schrodecoder.c:3395-3448): the frame holds the QUANTISED coefficients in the in-place subband
 * layout, `pairs` the (quant_factor, quant_offset + 2) of every codeblock in the order
 * component, band index, codeblock row, codeblock column (include/schro_b200.h). */
void
schro_b200_frame_dequantise (SchroFrame *frame, SchroParams *params, const int32_t *pairs)
{
  Sb2hContext *cx = sb2h_context ();
  const int is_s32 = depth_is_s32 (frame->format);
  const int bpp = is_s32 ? 4 : 2;
  sb2_dequant_params p;
  sb2_slab slab;
  size_t npairs, total = 0;
  char *region = frame->regions[0], *dev_region;
  void *dpairs;
  int k, host;

  SB2H_ASSERT (frame && params && pairs && region);
  memset (&p, 0, sizeof (p));
  p.transform_depth = params->transform_depth;
  if (p.transform_depth < 1 || p.transform_depth > SB2_DEQUANT_MAX_LEVELS)
    sb2h_fatal (__func__, "transform depth %d", p.transform_depth);
  for (k = 0; k <= SB2_DEQUANT_MAX_LEVELS; k++) {
    p.horiz_codeblocks[k] = k <= p.transform_depth && params->horiz_codeblocks[k] > 0 ? params->horiz_codeblocks[k] : 1;
    p.vert_codeblocks[k] = k <= p.transform_depth && params->vert_codeblocks[k] > 0 ? params->vert_codeblocks[k] : 1;
  }
  npairs = sb2_dequant_table_pairs (&p, 3);
  for (k = 0; k < 3; k++) total += (size_t) frame->components[k].length;
  host = sb2h_mem_kind (region) != SB2H_MEM_DEVICE;
  if (host) {
    dev_region = sb2h_dev_buffer (cx, SB2H_BUF_IN, total + 256);
    SB2H_CUDA (cudaMemcpyAsync (dev_region, region, total, cudaMemcpyDefault, cx->stream));
  } else {
    dev_region = region;
    sb2h_frame_use (cx, region);
  }
  memset (&slab, 0, sizeof (slab));
  slab.base = dev_region;
  slab.picture_pitch = total;
  slab.count = 1;
  slab.ncomp = 3;
  for (k = 0; k < 3; k++) {
    slab.offset[k] = (size_t) ((char *) frame->components[k].data - region);
    slab.stride[k] = frame->components[k].stride;
    slab.width[k] = k ? params->iwt_chroma_width : params->iwt_luma_width;
    slab.height[k] = k ? params->iwt_chroma_height : params->iwt_luma_height;
  }
  (void) bpp;
  dpairs = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, npairs * 2 * sizeof (int32_t));
  SB2H_CUDA (cudaMemcpyAsync (dpairs, pairs, npairs * 2 * sizeof (int32_t), cudaMemcpyDefault, cx->stream));
  SB2H_CHECK (sb2_dequantise (&slab, is_s32, &p, dpairs, npairs, cx->stream), "sb2_dequantise");
  if (host) {
    SB2H_CUDA (cudaMemcpyAsync (region, dev_region, total, cudaMemcpyDefault, cx->stream));
    sb2h_sync (cx);
  } else if (sb2h_mem_kind (pairs) != SB2H_MEM_PAGEABLE) {
    sb2h_sync (cx);        /* a page-locked table is read by the DMA engine after the copy call returns */
  } else {
    sb2h_frame_wrote (cx, region);
  }
}

/* The widening variant: `src` holds the quantised coefficients as s16 (what an entropy decoder
 * produces for any practical quantiser), `dest` receives the dequantised s32 coefficient frame
 * of a >8-bit stream (orc_dequantise_s32_ip_2d on the sign-extended values).  The host uploads
 * half the bytes of the s32 frame.  Frames in host memory are staged; CUDA-domain frames are
 * used where they lie and the call returns without waiting. */
void
schro_b200_frame_dequantise_widen (SchroFrame *dest, SchroFrame *src, SchroParams *params, const int32_t *pairs)
{
  Sb2hContext *cx = sb2h_context ();
  sb2_dequant_params p;
  sb2_slab in, out;
  size_t npairs, tin = 0, tout = 0;
  char *dev_in, *dev_out;
  void *dpairs;
  int k, host_in, host_out;

  SB2H_ASSERT (dest && src && params && pairs && dest->regions[0] && src->regions[0]);
  if (depth_is_s32 (src->format) || !depth_is_s32 (dest->format))
    sb2h_fatal (__func__, "needs an s16 source and an s32 destination (formats 0x%x, 0x%x)",
        (unsigned) src->format, (unsigned) dest->format);
  memset (&p, 0, sizeof (p));
  p.transform_depth = params->transform_depth;
  if (p.transform_depth < 1 || p.transform_depth > SB2_DEQUANT_MAX_LEVELS)
    sb2h_fatal (__func__, "transform depth %d", p.transform_depth);
  for (k = 0; k <= SB2_DEQUANT_MAX_LEVELS; k++) {
    p.horiz_codeblocks[k] = k <= p.transform_depth && params->horiz_codeblocks[k] > 0 ? params->horiz_codeblocks[k] : 1;
    p.vert_codeblocks[k] = k <= p.transform_depth && params->vert_codeblocks[k] > 0 ? params->vert_codeblocks[k] : 1;
  }
  npairs = sb2_dequant_table_pairs (&p, 3);
  for (k = 0; k < 3; k++) {
    tin += (size_t) src->components[k].length;
    tout += (size_t) dest->components[k].length;
  }
  host_in = sb2h_mem_kind (src->regions[0]) != SB2H_MEM_DEVICE;
  host_out = sb2h_mem_kind (dest->regions[0]) != SB2H_MEM_DEVICE;
  if (host_in) {
    dev_in = sb2h_dev_buffer (cx, SB2H_BUF_IN, tin + 256);
    SB2H_CUDA (cudaMemcpyAsync (dev_in, src->regions[0], tin, cudaMemcpyDefault, cx->stream));
  } else {
    dev_in = src->regions[0];
    sb2h_frame_use (cx, dev_in);
  }
  if (host_out) dev_out = sb2h_dev_buffer (cx, SB2H_BUF_OUT, tout + 256);
  else {
    dev_out = dest->regions[0];
    sb2h_frame_use (cx, dev_out);
  }
  memset (&in, 0, sizeof (in));
  memset (&out, 0, sizeof (out));
  in.base = dev_in;
  out.base = dev_out;
  in.picture_pitch = tin;
  out.picture_pitch = tout;
  in.count = out.count = 1;
  in.ncomp = out.ncomp = 3;
  for (k = 0; k < 3; k++) {
    in.offset[k] = (size_t) ((char *) src->components[k].data - (char *) src->regions[0]);
    out.offset[k] = (size_t) ((char *) dest->components[k].data - (char *) dest->regions[0]);
    in.stride[k] = src->components[k].stride;
    out.stride[k] = dest->components[k].stride;
    in.width[k] = out.width[k] = k ? params->iwt_chroma_width : params->iwt_luma_width;
    in.height[k] = out.height[k] = k ? params->iwt_chroma_height : params->iwt_luma_height;
    if (src->components[k].width < in.width[k] || src->components[k].height < in.height[k] ||
        dest->components[k].width < in.width[k] || dest->components[k].height < in.height[k])
      sb2h_fatal (__func__, "component %d is smaller than the %dx%d coefficient plane", k, in.width[k], in.height[k]);
  }
  dpairs = sb2h_dev_buffer (cx, SB2H_BUF_AUX0, npairs * 2 * sizeof (int32_t));
  SB2H_CUDA (cudaMemcpyAsync (dpairs, pairs, npairs * 2 * sizeof (int32_t), cudaMemcpyDefault, cx->stream));
  SB2H_CHECK (sb2_dequantise_widen (&in, &out, &p, dpairs, npairs, cx->stream), "sb2_dequantise_widen");
  if (host_out) {
    SB2H_CUDA (cudaMemcpyAsync (dest->regions[0], dev_out, tout, cudaMemcpyDefault, cx->stream));
    sb2h_sync (cx);
  } else {
    sb2h_frame_wrote (cx, dev_out);
    cx->dirty = 1;
    /* page-locked source or table memory is read by the DMA engine after the copy call returns */
    if (sb2h_mem_kind (pairs) != SB2H_MEM_PAGEABLE || (host_in && sb2h_mem_kind (src->regions[0]) != SB2H_MEM_PAGEABLE))
      sb2h_sync (cx);
  }
}
