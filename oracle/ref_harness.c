/*
 * ref_harness.c -- flat C-ABI over the UNMODIFIED reference, linked into
 * oracle/_ref/libschro_ref.so by oracle/build_ref.sh.  TEST INFRASTRUCTURE.
 * It only marshals plain pointers into the reference's own structs
 * (SchroFrameData, SchroFrame, SchroMotion, SchroHierBm ...) and calls the
 * reference's public entry points; no algorithm lives here.
 * The function shapes mirror oracle.h one to one (ref_* <-> oracle_*).
 */
#include <schroedinger/schro.h>
#include <schroedinger/schroframe.h>
#include <schroedinger/schrowavelet.h>
#include <schroedinger/schromotion.h>
#include <schroedinger/schromotionest.h>
#include <schroedinger/schrometric.h>
#include <schroedinger/schroencoder.h>
#include <schroedinger/schrodebug.h>
#include <stdlib.h>
#include <string.h>

static int ref_inited;
static void
ref_init (void)
{
  if (!ref_inited) {
    schro_init ();
    /* the reference creates its frame mutex in the first schro_frame_new (schroframe.c:36);
     * the frames built on the stack below never go through it */
    schro_frame_unref (schro_frame_new ());
    ref_inited = 1;
  }
}

int oracle_ref_harness_version (void) { return 3; }

static void
fill_fd (SchroFrameData *fd, void *data, int stride, int width, int height,
    int is_s32)
{
  memset (fd, 0, sizeof (*fd));
  fd->format = is_s32 ? SCHRO_FRAME_FORMAT_S32_444 : SCHRO_FRAME_FORMAT_S16_444;
  fd->data = data;
  fd->stride = stride;
  fd->width = width;
  fd->height = height;
}

/* schro_wavelet_transform_2d (schroedinger/schrowaveletorc.c:60) */
void
ref_wavelet_fwd (void *data, int stride, int width, int height, int is_s32,
    int filter)
{
  SchroFrameData fd;
  void *tmp = malloc ((size_t) (2 * width + 64) * 8);
  ref_init ();
  fill_fd (&fd, data, stride, width, height, is_s32);
  schro_wavelet_transform_2d (&fd, filter, tmp);
  free (tmp);
}

/* schro_wavelet_inverse_transform_2d (schroedinger/schrowaveletorc.c:121), dest == src */
void
ref_wavelet_inv (void *data, int stride, int width, int height, int is_s32,
    int filter)
{
  SchroFrameData fd;
  void *tmp = malloc ((size_t) (2 * width + 64) * 8);
  ref_init ();
  fill_fd (&fd, data, stride, width, height, is_s32);
  schro_wavelet_inverse_transform_2d (&fd, &fd, filter, tmp);
  free (tmp);
}

/* level loop exactly as schro_frame_iwt_transform (schroedinger/schroframe.c:1192-1228) */
void
ref_iwt_fwd (void *data, int stride, int width, int height, int is_s32,
    int filter, int depth)
{
  int level;
  void *tmp = malloc ((size_t) (2 * width + 64) * 8);
  ref_init ();
  for (level = 0; level < depth; level++) {
    SchroFrameData fd;
    fill_fd (&fd, data, stride << level, width >> level, height >> level, is_s32);
    schro_wavelet_transform_2d (&fd, filter, tmp);
  }
  free (tmp);
}

/* level loop exactly as schro_decoder_inverse_iwt_transform
 * (schroedinger/schrodecoder.c:1809-1853) */
void
ref_iwt_inv (void *data, int stride, int width, int height, int is_s32,
    int filter, int depth)
{
  int level;
  void *tmp = malloc ((size_t) (2 * width + 64) * 8);
  ref_init ();
  for (level = depth - 1; level >= 0; level--) {
    SchroFrameData fd;
    fill_fd (&fd, data, stride << level, width >> level, height >> level, is_s32);
    schro_wavelet_inverse_transform_2d (&fd, &fd, filter, tmp);
  }
  free (tmp);
}

/* ---- frame operations ---------------------------------------------------- */
/* A SchroFrame whose three components all alias one caller-owned plane (the operations
 * below are idempotent per component, so running them three times is harmless). */
static void
fake_frame (SchroFrame *f, uint8_t *data, int stride, int width, int height, int ext,
    int upsampled)
{
  int k;
  memset (f, 0, sizeof (*f));
  f->refcount = 1;
  f->format = SCHRO_FRAME_FORMAT_U8_444;
  f->width = width;
  f->height = height;
  f->extension = ext;
  f->is_upsampled = upsampled;
  for (k = 0; k < 3; k++) {
    f->components[k].format = SCHRO_FRAME_FORMAT_U8_444;
    f->components[k].data = data;
    f->components[k].stride = stride;
    f->components[k].width = width;
    f->components[k].height = height;
    f->components[k].length = stride * (height + 2 * ext);
  }
}

/* schro_frame_mc_edgeextend (schroedinger/schroframe.c:1986) */
void
ref_mc_edgeextend (uint8_t *data, int stride, int width, int height, int ext)
{
  SchroFrame f;
  ref_init ();
  fake_frame (&f, data, stride, width, height, ext, 0);
  schro_frame_mc_edgeextend (&f);
}

/* schro_upsampled_frame_upsample (schroedinger/schroframe.c:2000) */
void
ref_upsample (uint8_t *data, int stride, int width, int height, int ext)
{
  SchroFrame f;
  ref_init ();
  fake_frame (&f, data, stride, width, height, ext, 1);
  schro_upsampled_frame_upsample (&f);
}

/* schro_frame_downsample (schroedinger/schroframe.c:1505) */
void
ref_downsample (uint8_t *dest, int dstride, int dwidth, int dheight,
    const uint8_t *src, int sstride, int swidth, int sheight)
{
  SchroFrame fd, fs;
  ref_init ();
  fake_frame (&fd, dest, dstride, dwidth, dheight, 0, 0);
  fake_frame (&fs, (uint8_t *) src, sstride, swidth, sheight, 0, 0);
  schro_frame_downsample (&fd, &fs);
}

/* ---- OBMC ---------------------------------------------------------------- */
typedef struct {
  int width, height;            /* luma */
  int chroma_format;            /* SchroChromaFormat */
  int xbsep, ybsep, xblen, yblen;
  int x_num_blocks, y_num_blocks;
  int mv_precision, weight1, weight2, weight_bits, num_refs;
} RefMotionParams;

static void
fake_frame3 (SchroFrame *f, SchroFrameFormat fmt, void **data, const int *stride, int width,
    int height, int ext, int upsampled)
{
  int k;
  int hs = SCHRO_FRAME_FORMAT_H_SHIFT (fmt), vs = SCHRO_FRAME_FORMAT_V_SHIFT (fmt);
  memset (f, 0, sizeof (*f));
  f->refcount = 1;
  f->format = fmt;
  f->width = width;
  f->height = height;
  f->extension = ext;
  f->is_upsampled = upsampled;
  f->upsample_done = upsampled;
  for (k = 0; k < 3; k++) {
    f->components[k].format = fmt;
    f->components[k].data = data[k];
    f->components[k].stride = stride[k];
    f->components[k].width = k ? ROUND_UP_SHIFT (width, hs) : width;
    f->components[k].height = k ? ROUND_UP_SHIFT (height, vs) : height;
    f->components[k].h_shift = k ? hs : 0;
    f->components[k].v_shift = k ? vs : 0;
  }
}

/* global-motion model for the next ref_motion_render calls (NULL: off): 2 x 10 ints, SchroGlobalMotion's
 * fields in order.  With it set, the call goes through the dispatcher schro_motion_render, which switches
 * to the per-pixel renderer (schroedinger/schromotion.c:113-121). */
static const int *g_global_motion;
void ref_set_global_motion (const int *gm) { g_global_motion = gm; }

/* schro_motion_render_u8 (schroedinger/schromotion8.c:700) or, with use_ref_renderer,
 * the golden schro_motion_render_ref (schroedinger/schromotionref.c:245) */
void
ref_motion_render (const RefMotionParams *mp, const SchroMotionVector *mvs,
    void **ref0, const int *ref0_stride, void **ref1, const int *ref1_stride,
    void **acc, const int *acc_stride, void **residual, const int *res_stride,
    int res_is_s32, int add, void **out, const int *out_stride, int use_ref_renderer)
{
  SchroVideoFormat vf;
  SchroParams params;
  SchroMotion *motion;
  SchroFrame r0, r1, dest, addframe, output;
  int cf = mp->chroma_format;
  int n = mp->x_num_blocks * mp->y_num_blocks;

  ref_init ();
  memset (&vf, 0, sizeof (vf));
  vf.width = mp->width;
  vf.height = mp->height;
  vf.chroma_format = cf;
  memset (&params, 0, sizeof (params));
  params.video_format = &vf;
  params.num_refs = mp->num_refs;
  params.xbsep_luma = mp->xbsep;
  params.ybsep_luma = mp->ybsep;
  params.xblen_luma = mp->xblen;
  params.yblen_luma = mp->yblen;
  params.mv_precision = mp->mv_precision;
  params.picture_weight_1 = mp->weight1;
  params.picture_weight_2 = mp->weight2;
  params.picture_weight_bits = mp->weight_bits;
  params.x_num_blocks = mp->x_num_blocks;
  params.y_num_blocks = mp->y_num_blocks;
  params.x_offset = (mp->xblen - mp->xbsep) / 2;
  params.y_offset = (mp->yblen - mp->ybsep) / 2;

  {
    SchroFrameFormat u8 = schro_params_get_frame_format (8, cf);
    SchroFrameFormat s16 = schro_params_get_frame_format (16, cf);
    SchroFrameFormat s32 = schro_params_get_frame_format (32, cf);
    fake_frame3 (&r0, u8, ref0, ref0_stride, mp->width, mp->height, 32, 1);
    if (ref1) fake_frame3 (&r1, u8, ref1, ref1_stride, mp->width, mp->height, 32, 1);
    fake_frame3 (&dest, s16, acc, acc_stride, mp->width, mp->height, 0, 0);
    fake_frame3 (&addframe, res_is_s32 ? s32 : s16, residual, res_stride, mp->width, mp->height, 0, 0);
    if (out) fake_frame3 (&output, u8, out, out_stride, mp->width, mp->height, 0, 0);
  }
  if (g_global_motion) {
    params.have_global_motion = TRUE;
    memcpy (&params.global_motion[0], g_global_motion, sizeof (int) * 10);
    memcpy (&params.global_motion[1], g_global_motion + 10, sizeof (int) * 10);
  }
  motion = schro_motion_new (&params, &r0, ref1 ? &r1 : NULL);
  memcpy (motion->motion_vectors, mvs, sizeof (SchroMotionVector) * n);
  if (g_global_motion)
    schro_motion_render (motion, &dest, &addframe, add, out ? &output : NULL);
  else if (use_ref_renderer)
    schro_motion_render_ref (motion, &dest, &addframe, add, out ? &output : NULL);
  else
    schro_motion_render_u8 (motion, &dest, &addframe, add, out ? &output : NULL);
  schro_motion_free (motion);
}

/* ---- hierarchical block matching ------------------------------------------ */
typedef struct {
  int width, height;            /* luma */
  int chroma_format;
  int xbsep, ybsep;
  int levels;                   /* encoder->downsample_levels */
  int use_chroma;               /* encoder->enable_chroma_me */
  int ref_index;                /* 0 or 1 */
  int level0_range;             /* >0: also run scan_hint(0, range) as schromotionest.c:123-127 */
} RefHbmParams;

static SchroFrame *
load_frame (SchroFrameFormat fmt, int width, int height, void **data, const int *stride)
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

/* Builds both pyramids with the reference's schro_encoder_frame_downsample
 * (schroedinger/schroanalysis.c:9-28), then schro_hbm_new / schro_hbm_scan
 * (schroedinger/schrohierbm.c:25-172) and optionally the level-0 refinement.
 * fields: (levels+1) consecutive arrays of x_num_blocks*y_num_blocks vectors, index = level.
 * pyr_out (optional): levels*3 pointers receiving the src pyramid planes (level 1.., dense). */
void
ref_hbm_run (const RefHbmParams *hp, void **src, const int *src_stride, void **ref,
    const int *ref_stride, SchroMotionVector *fields, int *x_num_blocks, int *y_num_blocks,
    void **pyr_out)
{
  SchroEncoder *enc = calloc (1, sizeof (SchroEncoder));
  SchroEncoderFrame *fs = calloc (1, sizeof (SchroEncoderFrame));
  SchroEncoderFrame *fr = calloc (1, sizeof (SchroEncoderFrame));
  SchroVideoFormat vf;
  SchroFrameFormat fmt;
  SchroHierBm *hbm;
  int i, n, l;

  ref_init ();
  memset (&vf, 0, sizeof (vf));
  vf.width = hp->width;
  vf.height = hp->height;
  vf.chroma_format = hp->chroma_format;
  fmt = schro_params_get_frame_format (8, hp->chroma_format);
  enc->downsample_levels = hp->levels;
  enc->enable_chroma_me = hp->use_chroma;
  for (i = 0; i < 2; i++) {
    SchroEncoderFrame *f = i ? fr : fs;
    f->encoder = enc;
    f->params.video_format = &vf;
    f->params.xbsep_luma = hp->xbsep;
    f->params.ybsep_luma = hp->ybsep;
    f->params.xblen_luma = hp->xbsep;
    f->params.yblen_luma = hp->ybsep;
    f->params.num_refs = 1;
    schro_params_calculate_mc_sizes (&f->params);
  }
  fs->filtered_frame = load_frame (fmt, hp->width, hp->height, src, src_stride);
  fr->filtered_frame = load_frame (fmt, hp->width, hp->height, ref, ref_stride);
  schro_encoder_frame_downsample (fs);
  schro_encoder_frame_downsample (fr);
  fs->ref_frame[hp->ref_index] = fr;

  hbm = schro_hbm_new (fs, hp->ref_index);
  schro_hbm_scan (hbm);
  if (hp->level0_range > 0)
    schro_hierarchical_bm_scan_hint (hbm, 0, hp->level0_range);

  n = fs->params.x_num_blocks * fs->params.y_num_blocks;
  *x_num_blocks = fs->params.x_num_blocks;
  *y_num_blocks = fs->params.y_num_blocks;
  for (l = 0; l <= hp->levels; l++) {
    SchroMotionField *mf = schro_hbm_motion_field (hbm, l);
    if (mf) memcpy (fields + (size_t) l * n, mf->motion_vectors, sizeof (SchroMotionVector) * n);
    else memset (fields + (size_t) l * n, 0, sizeof (SchroMotionVector) * n);
  }
  if (pyr_out) {
    for (l = 0; l < hp->levels; l++) {
      int k, y;
      for (k = 0; k < 3; k++) {
        SchroFrameData *c = &fs->downsampled_frames[l]->components[k];
        uint8_t *d = pyr_out[l * 3 + k];
        if (!d) continue;
        for (y = 0; y < c->height; y++)
          memcpy (d + (size_t) c->width * y, SCHRO_FRAME_DATA_GET_LINE (c, y), c->width);
      }
    }
  }
  schro_hbm_unref (hbm);
  for (i = 0; i < hp->levels; i++) {
    schro_frame_unref (fs->downsampled_frames[i]);
    schro_frame_unref (fr->downsampled_frames[i]);
  }
  schro_frame_unref (fs->filtered_frame);
  schro_frame_unref (fr->filtered_frame);
  free (fs);
  free (fr);
  free (enc);
}

/* ---- rough (bigblock) motion search: schro_rough_me_heirarchical_scan
 * (schroedinger/schroroughmotion.c:46-60) = _nohint (levels, 12) then _hint (l, 4) for l = levels-1 .. 1;
 * with other distances the two functions are called directly in that order.
 * fields: (levels+1) arrays of x_num_blocks*y_num_blocks vectors, index = level (level 0 stays zero). */
void
ref_rough_run (const RefHbmParams *hp, int nohint_distance, int hint_distance, void **src, const int *src_stride,
    void **ref, const int *ref_stride, SchroMotionVector *fields, int *x_num_blocks, int *y_num_blocks)
{
  SchroEncoder *enc = calloc (1, sizeof (SchroEncoder));
  SchroEncoderFrame *fs = calloc (1, sizeof (SchroEncoderFrame));
  SchroEncoderFrame *fr = calloc (1, sizeof (SchroEncoderFrame));
  SchroVideoFormat vf;
  SchroFrameFormat fmt;
  SchroRoughME *rme;
  int i, n, l;

  ref_init ();
  memset (&vf, 0, sizeof (vf));
  vf.width = hp->width;
  vf.height = hp->height;
  vf.chroma_format = hp->chroma_format;
  fmt = schro_params_get_frame_format (8, hp->chroma_format);
  enc->downsample_levels = hp->levels;
  for (i = 0; i < 2; i++) {
    SchroEncoderFrame *f = i ? fr : fs;
    f->encoder = enc;
    f->params.video_format = &vf;
    f->params.xbsep_luma = hp->xbsep;
    f->params.ybsep_luma = hp->ybsep;
    f->params.xblen_luma = hp->xbsep;
    f->params.yblen_luma = hp->ybsep;
    f->params.num_refs = 1;
    schro_params_calculate_mc_sizes (&f->params);
    f->have_downsampling = TRUE;
  }
  fs->filtered_frame = load_frame (fmt, hp->width, hp->height, src, src_stride);
  fr->filtered_frame = load_frame (fmt, hp->width, hp->height, ref, ref_stride);
  schro_encoder_frame_downsample (fs);
  schro_encoder_frame_downsample (fr);
  fs->ref_frame[hp->ref_index] = fr;

  rme = schro_rough_me_new (fs, fr);
  if (nohint_distance == 12 && hint_distance == 4) {
    schro_rough_me_heirarchical_scan (rme);
  } else {
    schro_rough_me_heirarchical_scan_nohint (rme, hp->levels, nohint_distance);
    for (l = hp->levels - 1; l >= 1; l--) schro_rough_me_heirarchical_scan_hint (rme, l, hint_distance);
  }
  n = fs->params.x_num_blocks * fs->params.y_num_blocks;
  *x_num_blocks = fs->params.x_num_blocks;
  *y_num_blocks = fs->params.y_num_blocks;
  for (l = 0; l <= hp->levels; l++) {
    SchroMotionField *mf = rme->motion_fields[l];
    if (mf) memcpy (fields + (size_t) l * n, mf->motion_vectors, sizeof (SchroMotionVector) * n);
    else memset (fields + (size_t) l * n, 0, sizeof (SchroMotionVector) * n);
  }
  schro_rough_me_free (rme);
  for (i = 0; i < hp->levels; i++) {
    schro_frame_unref (fs->downsampled_frames[i]);
    schro_frame_unref (fr->downsampled_frames[i]);
  }
  schro_frame_unref (fs->filtered_frame);
  schro_frame_unref (fr->filtered_frame);
  free (fs);
  free (fr);
  free (enc);
}

/* ---- sub-pel refinement: schro_encoder_motion_predict_subpel_deep (schroedinger/schromotionest.c:246-355)
 * through a SchroMe built by the reference's own schro_me_new (:2758-2774).  The reference pictures are
 * edge-extended and upsampled by the reference; fields[r] (x_num_blocks*y_num_blocks vectors) are refined
 * in place.  q: width, height, xbsep, ybsep, mv_precision, num_refs. */
void
ref_subpel_run (const int *q, double lambda, void **src, const int *src_stride, void **ref0, const int *ref0_stride,
    void **ref1, const int *ref1_stride, SchroMotionVector *field0, SchroMotionVector *field1,
    int *x_num_blocks, int *y_num_blocks)
{
  SchroEncoder *enc = calloc (1, sizeof (SchroEncoder));
  SchroEncoderFrame *fs = calloc (1, sizeof (SchroEncoderFrame));
  SchroEncoderFrame *fr[2];
  SchroHierBm fake_hbm[2];
  SchroVideoFormat vf;
  SchroFrameFormat fmt;
  SchroMe *me;
  int r, n;
  const int num_refs = q[5];

  ref_init ();
  memset (&vf, 0, sizeof (vf));
  vf.width = q[0];
  vf.height = q[1];
  vf.chroma_format = SCHRO_CHROMA_420;
  fmt = schro_params_get_frame_format (8, SCHRO_CHROMA_420);
  fs->encoder = enc;
  fs->params.video_format = &vf;
  fs->params.xbsep_luma = q[2];
  fs->params.ybsep_luma = q[3];
  fs->params.xblen_luma = q[2];
  fs->params.yblen_luma = q[3];
  fs->params.num_refs = num_refs;
  fs->params.mv_precision = q[4];
  schro_params_calculate_mc_sizes (&fs->params);
  fs->frame_me_lambda = lambda;
  fs->filtered_frame = load_frame (fmt, q[0], q[1], src, src_stride);
  memset (fake_hbm, 0, sizeof (fake_hbm));
  for (r = 0; r < num_refs; r++) {
    fr[r] = calloc (1, sizeof (SchroEncoderFrame));
    fr[r]->upsampled_original_frame = load_frame (fmt, q[0], q[1], r ? ref1 : ref0, r ? ref1_stride : ref0_stride);
    schro_upsampled_frame_upsample (fr[r]->upsampled_original_frame);
    fs->ref_frame[r] = fr[r];
    fake_hbm[r].ref_count = 2;          /* schro_me_new refs it, schro_me_free unrefs it: never freed */
    fs->hier_bm[r] = &fake_hbm[r];
  }
  n = fs->params.x_num_blocks * fs->params.y_num_blocks;
  *x_num_blocks = fs->params.x_num_blocks;
  *y_num_blocks = fs->params.y_num_blocks;
  me = schro_me_new (fs);
  for (r = 0; r < num_refs; r++) {
    SchroMotionField *mf = schro_motion_field_new (fs->params.x_num_blocks, fs->params.y_num_blocks);
    memcpy (mf->motion_vectors, r ? field1 : field0, sizeof (SchroMotionVector) * n);
    schro_me_set_subpel_mf (me, mf, r);
  }
  schro_encoder_motion_predict_subpel_deep (me);
  for (r = 0; r < num_refs; r++)
    memcpy (r ? field1 : field0, schro_me_subpel_mf (me, r)->motion_vectors, sizeof (SchroMotionVector) * n);
  schro_me_free (me);                   /* frees the fields */
  for (r = 0; r < num_refs; r++) {
    schro_frame_unref (fr[r]->upsampled_original_frame);
    free (fr[r]);
  }
  schro_frame_unref (fs->filtered_frame);
  free (fs);
  free (enc);
}

/* schro_metric_absdiff_u8 (schroedinger/schrometric.c:10) */
/* schro_metric_scan_setup / _do_scan / _get_min and schro_metric_fast_block
 * (schroedinger/schrometric.c:31-214, 380-414) on two 4:2:0 u8 pictures loaded into frames with a
 * 32-pixel border.  q: x, y, block_width, block_height, dx, dy, dist, use_chroma.  out: ref_x, ref_y,
 * scan_width, scan_height, min dx, min dy, min metric, min chroma metric, fast_block metric. */
void
ref_metric_scan (int width, int height, void **src, const int *src_stride, void **ref, const int *ref_stride,
    const int *q, int *out, uint32_t *metrics, uint32_t *chroma_metrics)
{
  SchroFrameFormat fmt;
  SchroMetricScan *scan = calloc (1, sizeof (SchroMetricScan));
  SchroMetricInfo info;
  uint32_t chroma = 0;
  int dx = q[4], dy = q[5];
  ref_init ();
  fmt = schro_params_get_frame_format (8, SCHRO_CHROMA_420);
  scan->frame = load_frame (fmt, width, height, src, src_stride);
  scan->ref_frame = load_frame (fmt, width, height, ref, ref_stride);
  scan->x = q[0];
  scan->y = q[1];
  scan->block_width = q[2];
  scan->block_height = q[3];
  scan->gravity_x = dx;
  scan->gravity_y = dy;
  schro_metric_scan_setup (scan, dx, dy, q[6], q[7]);
  schro_metric_scan_do_scan (scan);
  out[6] = schro_metric_scan_get_min (scan, &dx, &dy, &chroma);
  out[0] = scan->ref_x;
  out[1] = scan->ref_y;
  out[2] = scan->scan_width;
  out[3] = scan->scan_height;
  out[4] = dx;
  out[5] = dy;
  out[7] = (int) chroma;
  memcpy (metrics, scan->metrics, sizeof (scan->metrics));
  memcpy (chroma_metrics, scan->chroma_metrics, sizeof (scan->chroma_metrics));
  schro_metric_info_init (&info, scan->frame, scan->ref_frame, q[2], q[3]);
  out[8] = schro_metric_fast_block (&info, q[0], q[1], q[4], q[5]);
  schro_frame_unref (scan->frame);
  schro_frame_unref (scan->ref_frame);
  free (scan);
}

uint32_t
ref_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int width, int height)
{
  ref_init ();
  return (uint32_t) schro_metric_absdiff_u8 ((uint8_t *) a, a_stride, (uint8_t *) b, b_stride, width, height);
}

/* ---- combine / convert glue --------------------------------------
 * depth: 0 u8, 1 s16, 2 s32; 4:2:0 frames described by three planes each. */
static SchroFrameFormat
fmt420 (int depth)
{
  return depth == 0 ? SCHRO_FRAME_FORMAT_U8_420 : depth == 1 ? SCHRO_FRAME_FORMAT_S16_420
      : SCHRO_FRAME_FORMAT_S32_420;
}

/* schro_frame_convert (schroedinger/schroframe.c:870) */
void
ref_frame_convert (void **dst, const int *dstride, int ddepth, int dwidth, int dheight,
    void **src, const int *sstride, int sdepth, int swidth, int sheight)
{
  SchroFrame d, s;
  ref_init ();
  fake_frame3 (&d, fmt420 (ddepth), dst, dstride, dwidth, dheight, 0, 0);
  fake_frame3 (&s, fmt420 (sdepth), src, sstride, swidth, sheight, 0, 0);
  s.refcount = 1000;                 /* schro_frame_convert refs / unrefs its source */
  schro_frame_convert (&d, &s);
}

/* schro_frame_add / schro_frame_subtract (schroedinger/schroframe.c:1012, 1062) */
void
ref_frame_add (void **dst, const int *dstride, int dwidth, int dheight,
    void **src, const int *sstride, int sdepth, int swidth, int sheight, int subtract)
{
  SchroFrame d, s;
  ref_init ();
  fake_frame3 (&d, fmt420 (1), dst, dstride, dwidth, dheight, 0, 0);
  fake_frame3 (&s, fmt420 (sdepth), src, sstride, swidth, sheight, 0, 0);
  if (subtract) schro_frame_subtract (&d, &s);
  else schro_frame_add (&d, &s);
}

/* schro_frame_shift_left / _right (schroedinger/schroframe.c:1238-1291) and schro_frame_md5 (:1817-1861) */
void
ref_frame_shift (void **planes, const int *stride, int depth, int width, int height, int shift, int right)
{
  SchroFrame f;
  ref_init ();
  fake_frame3 (&f, fmt420 (depth), planes, stride, width, height, 0, 0);
  if (right) schro_frame_shift_right (&f, shift);
  else schro_frame_shift_left (&f, shift);
}

void
ref_frame_md5 (void **planes, const int *stride, int depth, int width, int height, uint32_t *state)
{
  SchroFrame f;
  ref_init ();
  fake_frame3 (&f, fmt420 (depth), planes, stride, width, height, 0, 0);
  schro_frame_md5 (&f, state);
}

/* ---- low-delay slice decoder: schro_decoder_decode_lowdelay_transform_data (schroedinger/schrolowdelay.c:745-761)
 * on a SchroPicture that carries what the function reads: params, lowdelay_buffer, transform_frame.
 * q: luma width, luma height (iwt sizes; chroma = half), transform_depth, n_horiz_slices, n_vert_slices,
 * slice_bytes_num, slice_bytes_denom, is_s32, path (0: the dispatcher, 1: force _slow, 2: force _fast). */
#include <schroedinger/schrodecoder.h>
void schro_decoder_decode_lowdelay_transform_data_slow (SchroPicture * picture);
void schro_decoder_decode_lowdelay_transform_data_fast (SchroPicture * picture);
void
ref_lowdelay_decode (const int *q, const int *quant_matrix, const uint8_t *data, int data_bytes, void **planes,
    const int *strides)
{
  SchroPicture *pic = calloc (1, sizeof (SchroPicture));
  SchroBuffer buf;
  SchroFrame f;
  int i;
  ref_init ();
  memset (&buf, 0, sizeof (buf));
  buf.data = (unsigned char *) data;
  buf.length = (unsigned) data_bytes;
  buf.ref_count = 1;
  fake_frame3 (&f, q[7] ? SCHRO_FRAME_FORMAT_S32_420 : SCHRO_FRAME_FORMAT_S16_420, planes, strides, q[0], q[1], 0, 0);
  pic->params.transform_depth = q[2];
  pic->params.iwt_luma_width = q[0];
  pic->params.iwt_luma_height = q[1];
  pic->params.iwt_chroma_width = q[0] / 2;
  pic->params.iwt_chroma_height = q[1] / 2;
  pic->params.n_horiz_slices = q[3];
  pic->params.n_vert_slices = q[4];
  pic->params.slice_bytes_num = q[5];
  pic->params.slice_bytes_denom = q[6];
  pic->params.is_lowdelay = TRUE;
  for (i = 0; i < 1 + 3 * q[2]; i++) pic->params.quant_matrix[i] = quant_matrix[i];
  pic->lowdelay_buffer = &buf;
  pic->transform_frame = &f;
  if (q[8] == 1) schro_decoder_decode_lowdelay_transform_data_slow (pic);
  else if (q[8] == 2) schro_decoder_decode_lowdelay_transform_data_fast (pic);
  else schro_decoder_decode_lowdelay_transform_data (pic);
  free (pic);
}

/* ---- dequantisation: the reference's own subband geometry
 * (schro_subband_get_frame_data, schro_subband_get_position) and Orc kernels, driven codeblock by
 * codeblock the way schro_decoder_decode_subband does (schrodecoder.c:3559-3576, 3395-3448). */
#include <schroedinger/schroorc.h>
void
ref_dequantise_plane (void *data, int stride, int width, int height, int is_s32,
    int transform_depth, const int *hcb, const int *vcb, const int32_t *quant)
{
  SchroFrame f;
  SchroParams params;
  void *planes[3];
  int strides[3], index;
  ref_init ();
  memset (&params, 0, sizeof (params));
  params.transform_depth = transform_depth;
  params.iwt_luma_width = width;
  params.iwt_luma_height = height;
  planes[0] = planes[1] = planes[2] = data;
  strides[0] = strides[1] = strides[2] = stride;
  fake_frame3 (&f, is_s32 ? SCHRO_FRAME_FORMAT_S32_444 : SCHRO_FRAME_FORMAT_S16_444, planes, strides, width, height, 0, 0);
  for (index = 0; index <= 3 * transform_depth; index++) {
    const int position = schro_subband_get_position (index);
    const int nh = position == 0 ? hcb[0] : hcb[SCHRO_SUBBAND_SHIFT (position) + 1];
    const int nv = position == 0 ? vcb[0] : vcb[SCHRO_SUBBAND_SHIFT (position) + 1];
    SchroFrameData fd;
    int x, y;
    schro_subband_get_frame_data (&fd, &f, 0, position, &params);
    for (y = 0; y < nv; y++) {
      const int ymin = (fd.height * y) / nv, ymax = (fd.height * (y + 1)) / nv;
      int xmin = 0, acc = 0;
      const int cbw = fd.width / nh, inc = fd.width - nh * cbw;
      for (x = 0; x < nh; x++) {
        const int x0 = xmin;
        xmin += cbw;
        acc += inc;
        if (acc >= nh) { acc -= nh; xmin++; }
        if (xmin > x0 && ymax > ymin) {
          if (is_s32)
            orc_dequantise_s32_ip_2d (SCHRO_FRAME_DATA_GET_PIXEL_S32 (&fd, x0, ymin), fd.stride, quant[0], quant[1],
                xmin - x0, ymax - ymin);
          else
            orc_dequantise_s16_ip_2d (SCHRO_FRAME_DATA_GET_PIXEL_S16 (&fd, x0, ymin), fd.stride, quant[0], quant[1],
                xmin - x0, ymax - ymin);
        }
        quant += 2;
      }
    }
  }
}

/* the reference's quantiser tables (schroedinger/schrotables.c): 61 entries each */
void
ref_quant_tables (uint32_t *factor, uint32_t *offset_intra, uint32_t *offset_inter)
{
  int i;
  for (i = 0; i < 61; i++) {
    factor[i] = schro_table_quant[i];
    offset_intra[i] = schro_table_offset_1_2[i];
    offset_inter[i] = schro_table_offset_3_8[i];
  }
}
