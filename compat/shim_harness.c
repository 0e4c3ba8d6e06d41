/* compat/shim_harness.c -- TEST INFRASTRUCTURE for compat/schro_hbm_new.c.
 *
 * Compiled against the reference's headers together with the shim into oracle/_ref/libcompat_shim.so
 * (oracle/build_ref.sh).  It builds the two SchroEncoderFrame structures schro_hbm_new reads -- as the
 * reference declares them, schroencoder.h -- around frames and params supplied by the test, and calls
 * schro_hbm_new (frame, 0) exactly as schro_encoder_predict_pel_picture does (schromotionest.c:76).
 * schro_hbm_new_from_frames resolves to libschro_b200.so, loaded before this library. */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <schroedinger/schro.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schromotionest.h>

void schro_debug_log (int level, const char *file, const char *function, int line, const char *format, ...)
{
  (void) level; (void) file; (void) function; (void) line; (void) format;
}

typedef struct {
  SchroEncoder encoder;
  SchroEncoderFrame frame, ref_frame;
} ShimFixture;

/* returns the SchroHierBm of schro_hbm_new (frame, 0); *fixture_out must outlive it (the matcher keeps
 * a pointer to frame->params, as in the reference) and is released with compat_shim_free */
SchroHierBm *
compat_shim_hbm_new (SchroParams * params, int levels, int enable_chroma_me,
    SchroFrame ** src, SchroFrame ** ref, void **fixture_out)
{
  ShimFixture *fx = calloc (1, sizeof (ShimFixture));
  int i;
  fx->encoder.downsample_levels = levels;
  fx->encoder.enable_chroma_me = enable_chroma_me;
  fx->frame.encoder = &fx->encoder;
  fx->ref_frame.encoder = &fx->encoder;
  fx->frame.params = *params;
  fx->frame.ref_frame[0] = &fx->ref_frame;
  fx->frame.filtered_frame = src[0];
  fx->ref_frame.filtered_frame = ref[0];
  for (i = 0; i < levels; i++) {
    fx->frame.downsampled_frames[i] = src[i + 1];
    fx->ref_frame.downsampled_frames[i] = ref[i + 1];
  }
  *fixture_out = fx;
  return schro_hbm_new (&fx->frame, 0);
}

/* the same for the rough search: schro_rough_me_new (frame, frame->ref_frame[ref]) as
 * schro_encoder_motion_predict_rough calls it (schromotionest.c:72-75) */
SchroRoughME *
compat_shim_rough_me_new (SchroParams * params, int levels, int ref, SchroFrame ** src, SchroFrame ** reff,
    void **fixture_out)
{
  ShimFixture *fx = calloc (1, sizeof (ShimFixture));
  int i;
  fx->encoder.downsample_levels = levels;
  fx->frame.encoder = &fx->encoder;
  fx->ref_frame.encoder = &fx->encoder;
  fx->frame.params = *params;
  fx->frame.ref_frame[ref] = &fx->ref_frame;
  fx->frame.filtered_frame = src[0];
  fx->ref_frame.filtered_frame = reff[0];
  fx->frame.have_downsampling = TRUE;
  fx->ref_frame.have_downsampling = TRUE;
  for (i = 0; i < levels; i++) {
    fx->frame.downsampled_frames[i] = src[i + 1];
    fx->ref_frame.downsampled_frames[i] = reff[i + 1];
  }
  *fixture_out = fx;
  return schro_rough_me_new (&fx->frame, fx->frame.ref_frame[ref]);
}

/* the sub-pel refinement: schro_encoder_motion_predict_subpel_deep (me) on a SchroMe stand-in.  The real
 * SchroMe is private to schromotionest.c, so the five accessors the shim calls are defined here over
 * this fixture -- in the reference's tree they are the reference's own (schromotionest.c:2789-2885). */
struct _SchroMe {
  SchroParams *params;
  double lambda;
  SchroFrame *src, *ref[2];
  SchroMotionField *mf[2];
};
SchroParams *schro_me_params (SchroMe * me) { return me->params; }
double schro_me_lambda (SchroMe * me) { return me->lambda; }
SchroFrame *schro_me_src (SchroMe * me) { return me->src; }
SchroFrame *schro_me_ref (SchroMe * me, int ref) { return me->ref[ref]; }
SchroMotionField *schro_me_subpel_mf (SchroMe * me, int ref) { return me->mf[ref]; }

void
compat_shim_subpel_deep (SchroParams * params, double lambda, SchroFrame * src, SchroFrame ** refs, SchroMotionField ** mfs)
{
  struct _SchroMe me;
  int r;
  memset (&me, 0, sizeof (me));
  me.params = params;
  me.lambda = lambda;
  me.src = src;
  for (r = 0; r < params->num_refs; r++) { me.ref[r] = refs[r]; me.mf[r] = mfs[r]; }
  schro_encoder_motion_predict_subpel_deep (&me);
}

/* the low-delay slice decoder: schro_decoder_decode_lowdelay_transform_data (picture) on a SchroPicture as the
 * reference declares it (schrodecoder.h), filled with what the function reads */
#include <schroedinger/schrodecoder.h>
void
compat_shim_lowdelay (SchroParams * params, unsigned char *data, int length, SchroFrame * transform_frame)
{
  SchroPicture *pic = calloc (1, sizeof (SchroPicture));
  SchroBuffer buf;
  memset (&buf, 0, sizeof (buf));
  buf.data = data;
  buf.length = (unsigned) length;
  buf.ref_count = 1;
  pic->params = *params;
  pic->lowdelay_buffer = &buf;
  pic->transform_frame = transform_frame;
  schro_decoder_decode_lowdelay_transform_data (pic);
  free (pic);
}

void compat_shim_free (void *fixture) { free (fixture); }
