/*
 * schro_host_hbm.c -- hierarchical block matching entry points on top of the CUDA layer.
 *
 *   schro_motion_field_new / _free          schroedinger/schromotionest.c:395-415
 *   schro_hbm_new (as _new_from_frames)     schroedinger/schrohierbm.c:25-64
 *   schro_hbm_ref / _unref                  schroedinger/schrohierbm.c:66-116
 *   schro_hbm_motion_field                  schroedinger/schrohierbm.c:122-128
 *   schro_hbm_scan                          schroedinger/schrohierbm.c:158-172
 *   schro_hierarchical_bm_scan_hint         schroedinger/schrohierbm.c:174-383
 *
 * The pyramid levels are uploaded once per SchroHierBm (whole frame regions, borders
 * included) and the motion fields stay on the device between levels.  A level's field comes
 * back to the host when it is asked for (schro_hbm_motion_field): the reference's callers read
 * level 0 only (schromotionest.c:123-155), the coarser levels exist to seed the next one, and
 * copying all of them back cost 13 MB per 2160p picture on a link that is the bottleneck.
 */
#include "schro_host.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
  SchroHierBm pub;                       /* the reference's struct, first */
  void *dev_src[9], *dev_ref[9];         /* device copies of host pyramid levels (NULL: zero-copy) */
  void *dev_field[9];                    /* device motion field per level */
  int on_host[9];                        /* pub.downsampled_mf[level] holds the device field's current content */
  void *dev_ws;
  size_t ws_bytes;
  cudaEvent_t ev_done;                   /* position of the last enqueued level kernel (any thread may wait on it) */
  int have_done;
} Sb2hHierBm;

SchroMotionField *
schro_motion_field_new (int x_num_blocks, int y_num_blocks)
{
  SchroMotionField *mf = calloc (1, sizeof (SchroMotionField));
  mf->x_num_blocks = x_num_blocks;
  mf->y_num_blocks = y_num_blocks;
  mf->motion_vectors = calloc ((size_t) x_num_blocks * y_num_blocks, sizeof (SchroMotionVector));
  return mf;
}

void
schro_motion_field_free (SchroMotionField *field)
{
  /* fields that came back from the GPU live in pooled page-locked memory */
  if (!sb2h_pinned_pool_free (field->motion_vectors)) free (field->motion_vectors);
  free (field);
}

SchroHierBm *
schro_hbm_new_from_frames (SchroParams *params, int ref, int hierarchy_levels, schro_bool use_chroma,
    SchroFrame **src_frames, SchroFrame **ref_frames)
{
  Sb2hHierBm *h = calloc (1, sizeof (Sb2hHierBm));
  int i;
  SB2H_ASSERT (hierarchy_levels >= 1 && hierarchy_levels <= 8);
  h->pub.ref_count = 1;
  h->pub.ref = ref;
  h->pub.hierarchy_levels = hierarchy_levels;
  h->pub.params = params;
  h->pub.use_chroma = use_chroma ? 1 : 0;
  h->pub.downsampled_src = calloc ((size_t) hierarchy_levels + 1, sizeof (SchroFrame *));
  h->pub.downsampled_ref = calloc ((size_t) hierarchy_levels + 1, sizeof (SchroFrame *));
  h->pub.downsampled_mf = calloc ((size_t) hierarchy_levels + 1, sizeof (SchroMotionField *));
  for (i = 0; i <= hierarchy_levels; i++) {
    SB2H_ASSERT (src_frames[i] && ref_frames[i]);
    h->pub.downsampled_src[i] = schro_frame_ref (src_frames[i]);
    h->pub.downsampled_ref[i] = schro_frame_ref (ref_frames[i]);
  }
  return &h->pub;
}

SchroHierBm *
schro_hbm_ref (SchroHierBm *src)
{
  SB2H_ASSERT (src && src->ref_count > 0);
  ++src->ref_count;
  return src;
}

void
schro_hbm_unref (SchroHierBm *hbm)
{
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  int i;
  if (--hbm->ref_count > 0) return;
  if (h->have_done) {
    /* the searches may have been enqueued by another thread: the blocks go back to the pool
     * behind this thread's stream, which therefore waits for them first */
    Sb2hContext *cx = sb2h_context ();
    SB2H_CUDA (cudaStreamWaitEvent (cx->stream, h->ev_done, 0));
  }
  for (i = 0; i <= hbm->hierarchy_levels; i++) {
    if (hbm->downsampled_src[i]) schro_frame_unref (hbm->downsampled_src[i]);
    if (hbm->downsampled_ref[i]) schro_frame_unref (hbm->downsampled_ref[i]);
    if (hbm->downsampled_mf[i]) schro_motion_field_free (hbm->downsampled_mf[i]);
    sb2h_pool_free (h->dev_src[i]);
    sb2h_pool_free (h->dev_ref[i]);
    sb2h_pool_free (h->dev_field[i]);
  }
  sb2h_pool_free (h->dev_ws);
  if (h->ev_done) cudaEventDestroy (h->ev_done);
  free (hbm->downsampled_mf);
  free (hbm->downsampled_ref);
  free (hbm->downsampled_src);
  free (h);
}

/* schrohierbm.c:122-128.  The field is fetched from the device on first use after a search of
 * that level (page-locked host memory, one DMA, one wait). */
SchroMotionField *
schro_hbm_motion_field (SchroHierBm *hbm, int level)
{
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  SB2H_ASSERT (hbm && hbm->ref_count > 0 && level >= 0 && level <= hbm->hierarchy_levels);
  if (h->dev_field[level] && !h->on_host[level]) {
    Sb2hContext *cx = sb2h_context ();
    SchroParams *params = hbm->params;
    const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;
    SchroMotionField *mf = hbm->downsampled_mf[level];
    if (!mf) {
      mf = malloc (sizeof (SchroMotionField));
      mf->x_num_blocks = params->x_num_blocks;
      mf->y_num_blocks = params->y_num_blocks;
      mf->motion_vectors = sb2h_pinned_pool_alloc (n * sizeof (SchroMotionVector));
      hbm->downsampled_mf[level] = mf;
    }
    if (h->have_done) SB2H_CUDA (cudaStreamWaitEvent (cx->stream, h->ev_done, 0));
    SB2H_CUDA (cudaMemcpyAsync (mf->motion_vectors, h->dev_field[level], n * sizeof (SchroMotionVector),
            cudaMemcpyDefault, cx->stream));
    sb2h_sync (cx);
    h->on_host[level] = 1;
  }
  return hbm->downsampled_mf[level];
}

/* returns 1 when it enqueued an upload out of page-locked host memory (the caller then waits
 * before returning: the DMA engine reads that memory after the copy call) */
int
sb2h_level_slab (Sb2hContext *cx, SchroFrame *f, void **cache, sb2_slab *slab)
{
  const size_t bytes = (size_t) f->components[0].length + f->components[1].length + f->components[2].length;
  char *base;
  int k, staged = 0;
  if (SCHRO_FRAME_FORMAT_DEPTH (f->format) != SCHRO_FRAME_FORMAT_DEPTH_U8)
    sb2h_fatal (__func__, "block matching needs u8 frames");
  if (sb2h_mem_kind (f->regions[0]) == SB2H_MEM_DEVICE) {
    base = f->regions[0];
    sb2h_frame_use (cx, base);
  } else {
    if (!*cache) {
      *cache = sb2h_pool_alloc (bytes);
      SB2H_CUDA (cudaMemcpyAsync (*cache, f->regions[0], bytes, cudaMemcpyDefault, cx->stream));
      staged = sb2h_mem_kind (f->regions[0]) == SB2H_MEM_PINNED;
    }
    base = *cache;
  }
  memset (slab, 0, sizeof (*slab));
  slab->base = base;
  slab->picture_pitch = bytes;
  slab->count = 1;
  slab->ncomp = 3;
  for (k = 0; k < 3; k++) {
    slab->offset[k] = (size_t) ((char *) f->components[k].data - (char *) f->regions[0]);
    slab->stride[k] = f->components[k].stride;
    slab->width[k] = f->components[k].width;
    slab->height[k] = f->components[k].height;
  }
  return staged;
}

typedef struct {
  sb2_slab ss, rs;
  int staged;
} LevelIn;

/* Everything of a level that touches the thread's ordering stream: uploads of host pyramid
 * levels, waits on other threads' writes of device frames, allocations.  Done for every level
 * of a search before the one fork, so that the levels then chain on the priority stream alone. */
static void
prepare_level (Sb2hContext *cx, SchroHierBm *hbm, int shift, LevelIn *in)
{
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  SchroParams *params = hbm->params;
  const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;
  SB2H_ASSERT (shift >= 0 && shift <= hbm->hierarchy_levels);
  in->staged = sb2h_level_slab (cx, hbm->downsampled_src[shift], &h->dev_src[shift], &in->ss);
  in->staged |= sb2h_level_slab (cx, hbm->downsampled_ref[shift], &h->dev_ref[shift], &in->rs);
  if (!h->dev_ws) {
    h->ws_bytes = sb2_hbm_workspace_bytes (params->x_num_blocks, params->y_num_blocks, 1);
    h->dev_ws = sb2h_pool_alloc (h->ws_bytes);
  }
  if (!h->dev_field[shift])
    h->dev_field[shift] = sb2h_pool_alloc (n * sizeof (SchroMotionVector));
}

/* The block-matching kernels are dependency-latency bound: a block row makes progress only
 * while every row above it is resident.  They run on the thread's highest-priority stream so
 * that their CTAs are placed before those of the bandwidth kernels other threads have queued.
 * cx->stream stays the ordering backbone: one fork before the first level ... */
static void
fork_priority_stream (Sb2hContext *cx)
{
  SB2H_CUDA (cudaEventRecord (cx->ev_fork, cx->stream));
  SB2H_CUDA (cudaStreamWaitEvent (cx->stream_hi, cx->ev_fork, 0));
}

/* ... and one level: the kernel on the priority stream.  Its field stays on the device (the
 * next level reads it there); schro_hbm_motion_field brings it to the host on demand. */
static void
launch_level (Sb2hContext *cx, SchroHierBm *hbm, int shift, int h_range, const LevelIn *in)
{
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  SchroParams *params = hbm->params;
  SchroFrame *fs = hbm->downsampled_src[shift];
  sb2_hbm_params p;
  const size_t n = (size_t) params->x_num_blocks * params->y_num_blocks;

  memset (&p, 0, sizeof (p));
  p.xbsep = params->xbsep_luma;
  p.ybsep = params->ybsep_luma;
  p.x_num_blocks = params->x_num_blocks;
  p.y_num_blocks = params->y_num_blocks;
  p.ref_index = hbm->ref;
  p.use_chroma = hbm->use_chroma;
  p.chroma_h_shift = SCHRO_FRAME_FORMAT_H_SHIFT (fs->format);
  p.chroma_v_shift = SCHRO_FRAME_FORMAT_V_SHIFT (fs->format);
  SB2H_CHECK (sb2_hbm_scan_hint (&p, &in->ss, &in->rs, fs->extension, shift, h_range,
          shift < hbm->hierarchy_levels ? h->dev_field[shift + 1] : NULL, h->dev_field[shift], n,
          h->dev_ws, h->ws_bytes, cx->stream_hi), "sb2_hbm_scan_hint");
  h->on_host[shift] = 0;
  cx->dirty = 1;
}

/* after the last level of a call: the ordering stream rejoins the priority stream, and the
 * object remembers that position for whichever thread reads a field or frees the object */
static void
join_priority_stream (Sb2hContext *cx, Sb2hHierBm *h)
{
  SB2H_CUDA (cudaEventRecord (cx->ev_join, cx->stream_hi));
  SB2H_CUDA (cudaStreamWaitEvent (cx->stream, cx->ev_join, 0));
  if (!h->ev_done) SB2H_CUDA (cudaEventCreateWithFlags (&h->ev_done, cudaEventDisableTiming));
  SB2H_CUDA (cudaEventRecord (h->ev_done, cx->stream));
  h->have_done = 1;
}

void
schro_hierarchical_bm_scan_hint (SchroHierBm *hbm, int shift, int h_range)
{
  Sb2hContext *cx = sb2h_context ();
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  LevelIn in;
  if (h->have_done) SB2H_CUDA (cudaStreamWaitEvent (cx->stream, h->ev_done, 0));   /* earlier levels, any thread */
  prepare_level (cx, hbm, shift, &in);
  fork_priority_stream (cx);
  launch_level (cx, hbm, shift, h_range, &in);
  join_priority_stream (cx, h);
  if (in.staged) sb2h_sync (cx);
}

void
schro_hbm_scan (SchroHierBm *hbm)
{
  Sb2hContext *cx = sb2h_context ();
  Sb2hHierBm *h = (Sb2hHierBm *) hbm;
  LevelIn in[9];
  int i, half_scan_range = 20, staged = 0;
  const int n_levels = hbm->hierarchy_levels;
  SB2H_ASSERT (n_levels > 0);
  if (h->have_done) SB2H_CUDA (cudaStreamWaitEvent (cx->stream, h->ev_done, 0));
  for (i = n_levels; 1 <= i; --i) {
    prepare_level (cx, hbm, i, &in[i]);
    staged |= in[i].staged;
  }
  fork_priority_stream (cx);
  /* the levels chain on the priority stream; nothing waits: the fields are fetched on demand */
  launch_level (cx, hbm, n_levels, half_scan_range, &in[n_levels]);
  half_scan_range >>= 1;
  for (i = n_levels - 1; 1 <= i; --i, half_scan_range >>= 1)
    launch_level (cx, hbm, i, half_scan_range > 3 ? half_scan_range : 3, &in[i]);
  join_priority_stream (cx, h);
  if (staged) sb2h_sync (cx);
}
