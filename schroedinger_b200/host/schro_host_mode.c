/*
 * schro_host_mode.c -- the split-2 pass of the encoder's mode decision behind a host entry point.
 *
 *   schro_do_split2 + schro_motion_copy_to for every superblock   schroedinger/schromotionest.c:1601-1802, 1511-1523
 *
 * schro_do_split2 is static in the reference and takes a SchroMe, the structure private to
 * schromotionest.c; the things it reads through the accessors (schro_me_params / _lambda / _src / _ref /
 * _split2_mf / _motion, :2789-2900) are passed explicitly here, as for the sub-pel refinement.
 */
#include "schro_host.h"
#include <stdlib.h>
#include <string.h>

void
schro_b200_mode_decision_split2 (SchroParams *params, double lambda, SchroFrame *orig_frame,
    SchroFrame **upsampled_refs, SchroMotionField **split2_mfs, SchroMotion *motion, int *sb_error, int *sb_entropy)
{
  Sb2hContext *cx = sb2h_context ();
  const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;
  const size_t nsb = (size_t) (params->x_num_blocks / 4) * (params->y_num_blocks / 4);
  const size_t ws_bytes = sb2_split2_workspace_bytes (params->x_num_blocks, params->y_num_blocks, 1);
  void *dev_orig = NULL, *dev_ws, *dev_motion, *dev_field[2] = { NULL, NULL }, *dev_ref[2] = { NULL, NULL };
  int *dev_sb, *host_sb;
  sb2_slab os, rs[2];
  sb2_split2_params p;
  int ref;
  SB2H_ASSERT (params && orig_frame && upsampled_refs && split2_mfs && motion && motion->motion_vectors);
  SB2H_ASSERT (params->num_refs >= 1 && params->num_refs <= 2);
  sb2h_level_slab (cx, orig_frame, &dev_orig, &os);
  dev_ws = sb2h_pool_alloc (ws_bytes);
  dev_motion = sb2h_pool_alloc (n * sizeof (SchroMotionVector));
  dev_sb = sb2h_pool_alloc (2 * nsb * sizeof (int));
  host_sb = sb2h_pinned_pool_alloc (2 * nsb * sizeof (int) + n * sizeof (SchroMotionVector));
  for (ref = 0; ref < params->num_refs; ref++) {
    SchroFrame *up = upsampled_refs[ref];
    SchroMotionField *mf = split2_mfs[ref];
    SB2H_ASSERT (up && mf && mf->x_num_blocks == params->x_num_blocks && mf->y_num_blocks == params->y_num_blocks);
    SB2H_ASSERT (up->is_upsampled);
    if (!up->upsample_done) schro_upsampled_frame_upsample (up);
    sb2h_level_slab (cx, up, &dev_ref[ref], &rs[ref]);
    dev_field[ref] = sb2h_pool_alloc (n * sizeof (SchroMotionVector));
    sb2h_upload_staged (cx, dev_field[ref], mf->motion_vectors, n * sizeof (SchroMotionVector));
  }
  memset (&p, 0, sizeof (p));
  p.xblen = params->xbsep_luma;
  p.yblen = params->ybsep_luma;
  p.x_num_blocks = params->x_num_blocks;
  p.y_num_blocks = params->y_num_blocks;
  p.mv_precision = params->mv_precision;
  p.num_refs = params->num_refs;
  p.chroma_h_shift = SCHRO_CHROMA_FORMAT_H_SHIFT (params->video_format->chroma_format);
  p.chroma_v_shift = SCHRO_CHROMA_FORMAT_V_SHIFT (params->video_format->chroma_format);
  p.orig_extension = orig_frame->extension;
  p.lambda = lambda;
  SB2H_CHECK (sb2_split2_decide (&p, &os, &rs[0], params->num_refs > 1 ? &rs[1] : NULL, upsampled_refs[0]->extension,
          dev_field[0], dev_field[1], n, dev_motion, n, dev_sb, dev_sb + nsb, dev_ws, ws_bytes, cx->stream),
      "sb2_split2_decide");
  /* results come back through page-locked memory (one DMA each, no pageable staging inside the driver) */
  SB2H_CUDA (cudaMemcpyAsync (host_sb, dev_sb, 2 * nsb * sizeof (int), cudaMemcpyDeviceToHost, cx->stream));
  SB2H_CUDA (cudaMemcpyAsync (host_sb + 2 * nsb, dev_motion, n * sizeof (SchroMotionVector), cudaMemcpyDeviceToHost, cx->stream));
  cx->dirty = 1;
  sb2h_sync (cx);
  memcpy (motion->motion_vectors, host_sb + 2 * nsb, n * sizeof (SchroMotionVector));
  if (sb_error) memcpy (sb_error, host_sb, nsb * sizeof (int));
  if (sb_entropy) memcpy (sb_entropy, host_sb + nsb, nsb * sizeof (int));
  sb2h_pinned_pool_free (host_sb);
  for (ref = 0; ref < 2; ref++) {
    sb2h_pool_free (dev_field[ref]);
    sb2h_pool_free (dev_ref[ref]);
  }
  sb2h_pool_free (dev_sb);
  sb2h_pool_free (dev_motion);
  sb2h_pool_free (dev_ws);
  sb2h_pool_free (dev_orig);
}
