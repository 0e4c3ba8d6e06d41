/*
 * ref_me_static.c -- reaches the reference's file-static mode-decision functions.  TEST
 * INFRASTRUCTURE, built by oracle/build_ref.sh into oracle/_ref/libschro_ref_me.so (its own
 * shared object: this translation unit is the reference's schromotionest.c, included where it
 * lies under /root/reference and compiled unmodified, plus the entry point below; every other
 * reference function comes from libschro_ref.so).
 *
 * ref_split2_pass: schro_do_split2 (schroedinger/schromotionest.c:1601-1802) followed by
 * schro_motion_copy_to (:1511-1523) for every superblock in raster order -- the first step of
 * schro_mode_decision (:2587-2685) for each superblock, i.e. what that function produces when the
 * split-1 and split-0 candidates never win.
 */
#include "schroedinger/schromotionest.c"

static SchroFrame *
me_load_frame (SchroFrameFormat fmt, int width, int height, void **data, const int *stride)
{
  SchroFrame *f = schro_frame_new_and_alloc_full (NULL, fmt, width, height, 32, TRUE);
  int k, y;
  for (k = 0; k < 3; k++) {
    SchroFrameData *c = &f->components[k];
    for (y = 0; y < c->height; y++)
      memcpy (SCHRO_FRAME_DATA_GET_LINE (c, y), (uint8_t *) data[k] + (size_t) stride[k] * y, c->width);
  }
  schro_frame_mc_edgeextend (f);
  return f;
}

/* q: width, height, xbsep, ybsep, mv_precision, num_refs.  fields: the sub-pel fields of the two
 * references (split2_mf of the SchroMe).  motion_out: x_num_blocks * y_num_blocks vectors;
 * sb_error / sb_entropy: one int per superblock. */
void
ref_split2_pass (const int *q, double lambda, void **src, const int *src_stride, void **ref0, const int *ref0_stride,
    void **ref1, const int *ref1_stride, SchroMotionVector *field0, SchroMotionVector *field1,
    SchroMotionVector *motion_out, int *sb_error, int *sb_entropy, int *x_num_blocks, int *y_num_blocks)
{
  static int inited;
  SchroVideoFormat vf;
  SchroParams params;
  SchroFrameFormat fmt;
  struct _SchroMe me;
  struct SchroMeElement el[2];
  SchroMotion motion;
  SchroFrameData fd[2];
  SchroMotionField mf[2];
  const int num_refs = q[5];
  int r, i, j, n, block_size;

  if (!inited) {
    schro_init ();
    schro_frame_unref (schro_frame_new ());
    inited = 1;
  }
  memset (&vf, 0, sizeof (vf));
  vf.width = q[0];
  vf.height = q[1];
  vf.chroma_format = SCHRO_CHROMA_420;
  fmt = schro_params_get_frame_format (8, SCHRO_CHROMA_420);
  memset (&params, 0, sizeof (params));
  params.video_format = &vf;
  params.xbsep_luma = params.xblen_luma = q[2];
  params.ybsep_luma = params.yblen_luma = q[3];
  params.num_refs = num_refs;
  params.mv_precision = q[4];
  schro_params_calculate_mc_sizes (&params);
  n = params.x_num_blocks * params.y_num_blocks;
  *x_num_blocks = params.x_num_blocks;
  *y_num_blocks = params.y_num_blocks;

  memset (&me, 0, sizeof (me));
  memset (el, 0, sizeof (el));
  memset (&motion, 0, sizeof (motion));
  me.src = me_load_frame (fmt, q[0], q[1], src, src_stride);
  me.params = &params;
  me.lambda = lambda;
  me.motion = &motion;
  motion.params = &params;
  motion.motion_vectors = calloc ((size_t) n, sizeof (SchroMotionVector));
  for (r = 0; r < num_refs; r++) {
    el[r].ref = me_load_frame (fmt, q[0], q[1], r ? ref1 : ref0, r ? ref1_stride : ref0_stride);
    schro_upsampled_frame_upsample (el[r].ref);
    mf[r].x_num_blocks = params.x_num_blocks;
    mf[r].y_num_blocks = params.y_num_blocks;
    mf[r].motion_vectors = r ? field1 : field0;
    el[r].split2_mf = &mf[r];
    me.meElement[r] = &el[r];
  }
  /* the scratch blocks of schro_mode_decision (:2599-2610) */
  block_size = 16 * params.xbsep_luma * params.ybsep_luma;
  fd[0].data = fd[1].data = NULL;
  if (1 < params.mv_precision) {
    for (r = 0; r < num_refs; r++) {
      fd[r].data = schro_malloc (block_size * sizeof (uint8_t));
      fd[r].stride = fd[r].width = params.xbsep_luma << 2;
      fd[r].height = params.ybsep_luma << 2;
      fd[r].length = block_size * sizeof (uint8_t);
      fd[r].h_shift = fd[r].v_shift = 0;
      fd[r].format = SCHRO_FRAME_FORMAT_U8_420;
    }
  }
  for (j = 0; j < params.y_num_blocks; j += 4)
    for (i = 0; i < params.x_num_blocks; i += 4) {
      SchroBlock block = { 0 };
      const int sb = (j / 4) * (params.x_num_blocks / 4) + i / 4;
      schro_do_split2 (&me, i, j, &block, fd);
      schro_motion_copy_to (&motion, i, j, &block);
      sb_error[sb] = block.error;
      sb_entropy[sb] = block.entropy;
    }
  memcpy (motion_out, motion.motion_vectors, sizeof (SchroMotionVector) * (size_t) n);
  if (1 < params.mv_precision)
    for (r = 0; r < num_refs; r++) schro_free (fd[r].data);
  for (r = 0; r < num_refs; r++) schro_frame_unref (el[r].ref);
  schro_frame_unref (me.src);
  free (motion.motion_vectors);
}
