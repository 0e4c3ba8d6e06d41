/* compat/schro_rough_me_new.c -- the rough motion search's constructor, reference side.
 *
 * schro_rough_me_new (schroedinger/schroroughmotion.c:21-33) stores two SchroEncoderFrame pointers;
 * the level functions then read the frame's params, the encoder's downsample_levels, which of
 * ref_frame[0] / [1] the reference is (:75-78), and both pictures' filtered + downsampled frames
 * (get_downsampled, :303-312).  libschro_b200 does not know SchroEncoderFrame, so it exports the
 * constructor with those things passed explicitly (schro_rough_me_new_from_frames,
 * include/schro_b200_compat.h).  This file is compiled AGAINST THE REFERENCE'S OWN HEADERS and
 * replaces the body in schroroughmotion.c; schro_encoder_motion_predict_rough
 * (schroedinger/schromotionest.c:72-75) keeps calling schro_rough_me_new (frame, ref) unchanged.
 * oracle/build_ref.sh compiles it where the reference is present. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <schroedinger/schro.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schromotionest.h>

SchroRoughME *schro_rough_me_new_from_frames (SchroEncoderFrame * frame, SchroEncoderFrame * ref_frame,
    SchroParams * params, int ref, int levels, SchroFrame ** src_frames, SchroFrame ** ref_frames);

SchroRoughME *
schro_rough_me_new (SchroEncoderFrame * frame, SchroEncoderFrame * ref)
{
  SchroFrame *src[SCHRO_MAX_HIER_LEVELS + 1];
  SchroFrame *rf[SCHRO_MAX_HIER_LEVELS + 1];
  int levels = frame->encoder->downsample_levels;
  int which = ref == frame->ref_frame[0] ? 0 : (ref == frame->ref_frame[1] ? 1 : -1);
  int i;

  SCHRO_ASSERT (which != -1);
  SCHRO_ASSERT (frame->have_downsampling && ref->have_downsampling);
  SCHRO_ASSERT (levels < SCHRO_MAX_HIER_LEVELS);
  src[0] = frame->filtered_frame;
  rf[0] = ref->filtered_frame;
  for (i = 0; i < levels; i++) {
    src[i + 1] = frame->downsampled_frames[i];
    rf[i + 1] = ref->downsampled_frames[i];
  }
  return schro_rough_me_new_from_frames (frame, ref, &frame->params, which, levels, src, rf);
}
