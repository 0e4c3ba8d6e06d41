/*
 * schro_host_rough.c -- the rough ("bigblock") motion search behind the reference's entry points.
 *
 *   schro_rough_me_new / _free                       schroedinger/schroroughmotion.c:21-45
 *   schro_rough_me_heirarchical_scan                 schroedinger/schroroughmotion.c:46-60
 *   schro_rough_me_heirarchical_scan_nohint / _hint  schroedinger/schroroughmotion.c:62-300
 *
 * Pyramid levels are used in place when they live in a CUDA domain and uploaded once otherwise; the
 * fields stay on the device between levels and a copy of each comes back into page-locked memory
 * behind its kernel, because the reference's callers read rme->motion_fields[] directly.
 */
#include "schro_host.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
  SchroRoughME pub;                      /* the reference's struct, first */
  SchroParams *params;
  int ref, levels;
  SchroFrame *src[SCHRO_MAX_HIER_LEVELS + 1], *rf[SCHRO_MAX_HIER_LEVELS + 1];
  void *dev_src[SCHRO_MAX_HIER_LEVELS + 1], *dev_ref[SCHRO_MAX_HIER_LEVELS + 1];
  void *dev_field[SCHRO_MAX_HIER_LEVELS + 1];
  void *dev_ws;
  size_t ws_bytes;
} Sb2hRoughME;

SchroRoughME *
schro_rough_me_new_from_frames (struct _SchroEncoderFrame *frame, struct _SchroEncoderFrame *ref_frame,
    SchroParams *params, int ref, int levels, SchroFrame **src_frames, SchroFrame **ref_frames)
{
  Sb2hRoughME *r = calloc (1, sizeof (Sb2hRoughME));
  int i;
  SB2H_ASSERT (params && levels >= 1 && levels < SCHRO_MAX_HIER_LEVELS && (ref == 0 || ref == 1));
  r->pub.encoder_frame = frame;
  r->pub.ref_frame = ref_frame;
  r->params = params;
  r->ref = ref;
  r->levels = levels;
  for (i = 0; i <= levels; i++) {
    SB2H_ASSERT (src_frames[i] && ref_frames[i]);
    r->src[i] = schro_frame_ref (src_frames[i]);
    r->rf[i] = schro_frame_ref (ref_frames[i]);
  }
  return &r->pub;
}

void
schro_rough_me_free (SchroRoughME *rme)
{
  Sb2hRoughME *r = (Sb2hRoughME *) rme;
  int i;
  for (i = 0; i < SCHRO_MAX_HIER_LEVELS; i++)
    if (rme->motion_fields[i]) schro_motion_field_free (rme->motion_fields[i]);
  for (i = 0; i <= r->levels; i++) {
    schro_frame_unref (r->src[i]);
    schro_frame_unref (r->rf[i]);
    sb2h_pool_free (r->dev_src[i]);
    sb2h_pool_free (r->dev_ref[i]);
    sb2h_pool_free (r->dev_field[i]);
  }
  sb2h_pool_free (r->dev_ws);
  free (r);
}

/* one level: kernel(s) + the copy of its field, all on the thread's stream; no wait */
static void
enqueue_level (Sb2hContext *cx, Sb2hRoughME *r, int shift, int distance, int hint)
{
  SchroParams *params = r->params;
  const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;
  sb2_slab ss, rs;
  sb2_hbm_params p;
  SchroMotionField *mf;
  SB2H_ASSERT (shift >= 0 && shift <= r->levels);
  SB2H_ASSERT (params->x_num_blocks != 0 && params->y_num_blocks != 0);
  sb2h_level_slab (cx, r->src[shift], &r->dev_src[shift], &ss);
  sb2h_level_slab (cx, r->rf[shift], &r->dev_ref[shift], &rs);
  if (!r->dev_field[shift]) r->dev_field[shift] = sb2h_pool_alloc (n * sizeof (SchroMotionVector));
  memset (&p, 0, sizeof (p));
  p.xbsep = params->xbsep_luma;
  p.ybsep = params->ybsep_luma;
  p.x_num_blocks = params->x_num_blocks;
  p.y_num_blocks = params->y_num_blocks;
  p.ref_index = r->ref;
  if (hint) {
    SB2H_ASSERT (shift < r->levels && r->dev_field[shift + 1]);     /* rme->motion_fields[shift + 1] (:169) */
    if (!r->dev_ws) {
      r->ws_bytes = sb2_rough_workspace_bytes (params->x_num_blocks, params->y_num_blocks, 1);
      r->dev_ws = sb2h_pool_alloc (r->ws_bytes);
    }
    SB2H_CHECK (sb2_rough_scan_hint (&p, &ss, &rs, r->src[shift]->extension, shift, distance, r->dev_field[shift + 1],
            r->dev_field[shift], n, r->dev_ws, r->ws_bytes, cx->stream), "sb2_rough_scan_hint");
  } else {
    SB2H_CHECK (sb2_rough_scan_nohint (&p, &ss, &rs, r->src[shift]->extension, shift, distance, r->dev_field[shift], n,
            cx->stream), "sb2_rough_scan_nohint");
  }
  mf = r->pub.motion_fields[shift];
  if (!mf) {
    mf = malloc (sizeof (SchroMotionField));
    mf->x_num_blocks = params->x_num_blocks;
    mf->y_num_blocks = params->y_num_blocks;
    mf->motion_vectors = sb2h_pinned_pool_alloc (n * sizeof (SchroMotionVector));
    r->pub.motion_fields[shift] = mf;
  }
  SB2H_CUDA (cudaMemcpyAsync (mf->motion_vectors, r->dev_field[shift], n * sizeof (SchroMotionVector),
          cudaMemcpyDefault, cx->stream));
  cx->dirty = 1;
}

void
schro_rough_me_heirarchical_scan_nohint (SchroRoughME *rme, int shift, int distance)
{
  Sb2hContext *cx = sb2h_context ();
  enqueue_level (cx, (Sb2hRoughME *) rme, shift, distance, 0);
  sb2h_sync (cx);
}

void
schro_rough_me_heirarchical_scan_hint (SchroRoughME *rme, int shift, int distance)
{
  Sb2hContext *cx = sb2h_context ();
  enqueue_level (cx, (Sb2hRoughME *) rme, shift, distance, 1);
  sb2h_sync (cx);
}

void
schro_rough_me_heirarchical_scan (SchroRoughME *rme)
{
  Sb2hContext *cx = sb2h_context ();
  Sb2hRoughME *r = (Sb2hRoughME *) rme;
  int i;
  /* (:46-60) the whole chain is enqueued, the host waits once */
  enqueue_level (cx, r, r->levels, 12, 0);
  for (i = r->levels - 1; i >= 1; i--) enqueue_level (cx, r, i, 4, 1);
  sb2h_sync (cx);
}
