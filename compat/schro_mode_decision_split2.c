/* compat/schro_mode_decision_split2.c -- the split-2 pass of schro_mode_decision, reference side.
 *
 * schro_do_split2 (schroedinger/schromotionest.c:1601-1802) is static and takes the SchroMe that is
 * private to schromotionest.c; everything it reads from it is available through public accessors
 * (schro_me_params / _lambda / _src / _ref / _split2_mf / _motion, schromotionest.h:126-143).  This
 * helper reads those and hands them to libschro_b200 (schro_b200_mode_decision_split2,
 * include/schro_b200_compat.h), which runs the pass for every superblock of the picture at once.  Compiled
 * AGAINST THE REFERENCE'S OWN HEADERS by oracle/build_ref.sh.  A maintainer calls it at the top of
 * schro_mode_decision and takes block.mv / .error / .entropy from its results instead of calling
 * schro_do_split2 per superblock (INTEGRATION.md says when that is exact). */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <schroedinger/schro.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schromotionest.h>

void schro_b200_mode_decision_split2 (SchroParams * params, double lambda, SchroFrame * orig_frame,
    SchroFrame ** upsampled_refs, SchroMotionField ** split2_mfs, SchroMotion * motion, int *sb_error,
    int *sb_entropy);

/* sb_error / sb_entropy: (x_num_blocks / 4) * (y_num_blocks / 4) ints, SchroBlock.error / .entropy of every
 * superblock in raster order; schro_me_motion (me)->motion_vectors receives the decided blocks */
void
schro_mode_decision_split2_pass (SchroMe * me, int *sb_error, int *sb_entropy)
{
  SchroParams *params = schro_me_params (me);
  SchroFrame *up[2] = { NULL, NULL };
  SchroMotionField *mf[2] = { NULL, NULL };
  int ref;

  for (ref = 0; ref < params->num_refs; ref++) {
    up[ref] = schro_me_ref (me, ref);
    mf[ref] = schro_me_split2_mf (me, ref);
  }
  schro_b200_mode_decision_split2 (params, schro_me_lambda (me), schro_me_src (me), up, mf,
      schro_me_motion (me), sb_error, sb_entropy);
}
