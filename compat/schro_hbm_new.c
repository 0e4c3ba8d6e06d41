/* compat/schro_hbm_new.c -- the one constructor of the hot path that reads an encoder structure.
 *
 * schro_hbm_new (schroedinger/schrohierbm.c:25-64) takes a SchroEncoderFrame and reads five things
 * from it: the frame's params, the encoder's downsample_levels and enable_chroma_me, and the
 * filtered + downsampled frames of the picture and of its reference.  libschro_b200 does not know
 * SchroEncoderFrame (the encoder is out of scope), so it exports the same
 * constructor with those five things passed explicitly, schro_hbm_new_from_frames
 * (include/schro_b200_compat.h).  This file is the reference-side half: it is compiled AGAINST THE
 * REFERENCE'S OWN HEADERS and dropped into the reference's tree in place of the body in
 * schrohierbm.c, so that schro_encoder_predict_pel_picture and every other caller keep calling
 * schro_hbm_new (frame, ref) unchanged.
 *
 *   gcc -DSCHRO_ENABLE_UNSTABLE_API -I<reference> -I<orc> -c compat/schro_hbm_new.c
 *
 * oracle/build_ref.sh compiles it where the reference is present (oracle/_ref/obj/compat_schro_hbm_new.o)
 * and tests/test_compat_shim.py checks what it defines and what it leaves to the library. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <schroedinger/schro.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schromotionest.h>

/* exported by libschro_b200.so; frames[0] = full resolution, frames[i] = pyramid level i */
SchroHierBm *schro_hbm_new_from_frames (SchroParams * params, int ref,
    int hierarchy_levels, int use_chroma, SchroFrame ** src_frames, SchroFrame ** ref_frames);

SchroHierBm *
schro_hbm_new (SchroEncoderFrame * frame, int ref)
{
  SchroEncoderFrame *ref_frame = frame->ref_frame[ref];
  SchroFrame *src[SCHRO_MAX_HIER_LEVELS + 1];
  SchroFrame *rf[SCHRO_MAX_HIER_LEVELS + 1];
  int levels = frame->encoder->downsample_levels;
  int i;

  SCHRO_ASSERT (ref_frame);
  SCHRO_ASSERT (levels <= SCHRO_MAX_HIER_LEVELS);
  src[0] = frame->filtered_frame;
  rf[0] = ref_frame->filtered_frame;
  for (i = 0; i < levels; i++) {
    SCHRO_ASSERT (frame->downsampled_frames[i] && ref_frame->downsampled_frames[i]);
    src[i + 1] = frame->downsampled_frames[i];
    rf[i + 1] = ref_frame->downsampled_frames[i];
  }
  /* the library takes its own references on the frames, as the reference's constructor does */
  return schro_hbm_new_from_frames (&frame->params, ref, levels,
      frame->encoder->enable_chroma_me ? TRUE : FALSE, src, rf);
}
