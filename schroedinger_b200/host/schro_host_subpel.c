/*
 * schro_host_subpel.c -- sub-pel refinement of the motion fields behind the reference's entry point.
 *
 *   schro_encoder_motion_predict_subpel_deep   schroedinger/schromotionest.c:246-355
 *
 * The reference's function takes a SchroMe, a structure private to schromotionest.c that it only
 * touches through five accessors (schro_me_params / _lambda / _src / _ref / _subpel_mf, :2789-2885).
 * The library exports the same work with those five things passed explicitly; compat/schro_subpel_deep.c
 * is the reference-side half that keeps the symbol and its signature.
 */
#include "schro_host.h"
#include <stdlib.h>
#include <string.h>

void
schro_b200_motion_predict_subpel_deep (SchroParams *params, double lambda, SchroFrame *orig_frame,
    SchroFrame **upsampled_refs, SchroMotionField **subpel_mfs)
{
  Sb2hContext *cx = sb2h_context ();
  const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;
  const size_t ws_bytes = sb2_subpel_workspace_bytes (params->x_num_blocks, params->y_num_blocks, 1);
  void *dev_orig = NULL, *dev_ws, *dev_field[2] = { NULL, NULL }, *dev_ref[2] = { NULL, NULL };
  sb2_slab os, rs;
  int ref;
  SB2H_ASSERT (params && orig_frame && upsampled_refs && subpel_mfs);
  SB2H_ASSERT (params->num_refs >= 0 && params->num_refs <= 2);
  if (params->num_refs == 0 || params->mv_precision < 1) return;
  sb2h_level_slab (cx, orig_frame, &dev_orig, &os);
  dev_ws = sb2h_pool_alloc (ws_bytes);
  for (ref = 0; ref < params->num_refs; ref++) {
    SchroFrame *up = upsampled_refs[ref];
    SchroMotionField *mf = subpel_mfs[ref];
    sb2_subpel_params p;
    SB2H_ASSERT (up && mf && mf->x_num_blocks == params->x_num_blocks && mf->y_num_blocks == params->y_num_blocks);
    SB2H_ASSERT (up->is_upsampled);
    if (!up->upsample_done) schro_upsampled_frame_upsample (up);
    sb2h_level_slab (cx, up, &dev_ref[ref], &rs);
    dev_field[ref] = sb2h_pool_alloc (n * sizeof (SchroMotionVector));
    sb2h_upload_staged (cx, dev_field[ref], mf->motion_vectors, n * sizeof (SchroMotionVector));
    memset (&p, 0, sizeof (p));
    p.xblen = params->xbsep_luma;
    p.yblen = params->ybsep_luma;
    p.x_num_blocks = params->x_num_blocks;
    p.y_num_blocks = params->y_num_blocks;
    p.mv_precision = params->mv_precision;
    p.ref_index = ref;
    p.orig_extension = orig_frame->extension;
    p.lambda = lambda;
    SB2H_CHECK (sb2_subpel_refine (&p, &os, &rs, up->extension, dev_field[ref], n, dev_ws, ws_bytes, cx->stream),
        "sb2_subpel_refine");
    SB2H_CUDA (cudaMemcpyAsync (mf->motion_vectors, dev_field[ref], n * sizeof (SchroMotionVector), cudaMemcpyDefault, cx->stream));
  }
  cx->dirty = 1;
  sb2h_sync (cx);
  for (ref = 0; ref < 2; ref++) {
    sb2h_pool_free (dev_field[ref]);
    sb2h_pool_free (dev_ref[ref]);
  }
  sb2h_pool_free (dev_ws);
  sb2h_pool_free (dev_orig);
}
