/* compat/schro_subpel_deep.c -- schro_encoder_motion_predict_subpel_deep, reference side.
 *
 * The reference's function (schroedinger/schromotionest.c:246-355) takes a SchroMe, whose layout is
 * private to schromotionest.c; it reads it through five public accessors only.  This file keeps the
 * symbol and its signature, reads the same five things and hands them to libschro_b200
 * (schro_b200_motion_predict_subpel_deep, include/schro_b200_compat.h).  Compiled AGAINST THE
 * REFERENCE'S OWN HEADERS by oracle/build_ref.sh; it replaces the body in schromotionest.c, so
 * schro_encoder_predict_subpel_picture (schroedinger/schroencoder.c:2316-2320) is unchanged. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <schroedinger/schro.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schromotionest.h>

void schro_b200_motion_predict_subpel_deep (SchroParams * params, double lambda, SchroFrame * orig_frame,
    SchroFrame ** upsampled_refs, SchroMotionField ** subpel_mfs);

void
schro_encoder_motion_predict_subpel_deep (SchroMe * me)
{
  SchroParams *params = schro_me_params (me);
  SchroFrame *up[2] = { NULL, NULL };
  SchroMotionField *mf[2] = { NULL, NULL };
  int ref;

  for (ref = 0; ref < params->num_refs; ref++) {
    up[ref] = schro_me_ref (me, ref);
    mf[ref] = schro_me_subpel_mf (me, ref);
  }
  schro_b200_motion_predict_subpel_deep (params, schro_me_lambda (me), schro_me_src (me), up, mf);
}
