/*
 * schro_b200_compat.h -- the reference's C API for the picture core, as exported
 * by libschro_b200.so (host layer in schroedinger_b200/host/, C, calling the
 * sb2_* CUDA layer of schro_b200.h).
 *
 * The structs below are layout-compatible with the reference's (same field
 * order and types; tests/test_abi_layout.py compiles both headers and compares
 * sizeof/offsetof where /root/reference is available), so a libschroedinger
 * build can route these symbols here without touching callers.  Each
 * declaration cites the reference declaration it stands in for.
 *
 * Frames may live in
 *   - ordinary host memory           (malloc; copies are staged through pinned buffers),
 *   - pinned host memory             (schro_memory_domain_new_pinned; direct DMA),
 *   - device memory                  (schro_memory_domain_new_cuda; zero-copy),
 * and every entry point accepts all three; the arithmetic always runs on the GPU.
 * Errors follow the reference: log + abort (schroedinger/schrodebug.h:55-60).
 */
#ifndef SCHRO_B200_COMPAT_H
#define SCHRO_B200_COMPAT_H

#include <limits.h>
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int schro_bool;          /* schroedinger/schroutils.h:26 */

/* schroedinger/schroframe.h:22-44 */
typedef enum _SchroFrameFormat {
  SCHRO_FRAME_FORMAT_U8_444 = 0x00,
  SCHRO_FRAME_FORMAT_U8_422 = 0x01,
  SCHRO_FRAME_FORMAT_U8_420 = 0x03,
  SCHRO_FRAME_FORMAT_S16_444 = 0x04,
  SCHRO_FRAME_FORMAT_S16_422 = 0x05,
  SCHRO_FRAME_FORMAT_S16_420 = 0x07,
  SCHRO_FRAME_FORMAT_S32_444 = 0x08,
  SCHRO_FRAME_FORMAT_S32_422 = 0x09,
  SCHRO_FRAME_FORMAT_S32_420 = 0x0b
} SchroFrameFormat;

#define SCHRO_FRAME_FORMAT_DEPTH(format) ((format) & 0xc)
#define SCHRO_FRAME_FORMAT_DEPTH_U8 0x00
#define SCHRO_FRAME_FORMAT_DEPTH_S16 0x04
#define SCHRO_FRAME_FORMAT_DEPTH_S32 0x08
#define SCHRO_FRAME_FORMAT_H_SHIFT(format) ((format) & 0x1)
#define SCHRO_FRAME_FORMAT_V_SHIFT(format) (((format)>>1) & 0x1)
#define SCHRO_FRAME_CACHE_SIZE 32

typedef struct _SchroFrame SchroFrame;
typedef struct _SchroFrameData SchroFrameData;
typedef struct _SchroMemoryDomain SchroMemoryDomain;
typedef void (*SchroFrameFreeFunc) (SchroFrame *frame, void *priv);

/* schroedinger/schroframe.h:58-67 */
struct _SchroFrameData {
  SchroFrameFormat format;
  void *data;
  int stride;
  int width;
  int height;
  int length;
  int h_shift;
  int v_shift;
};

/* schroedinger/schroframe.h:69-94 */
struct _SchroFrame {
  int refcount;
  SchroFrameFreeFunc free;
  SchroMemoryDomain *domain;
  void *regions[3];
  void *priv;

  SchroFrameFormat format;
  int width;
  int height;

  SchroFrameData components[3];

  int is_virtual;
  int cached_lines[3][SCHRO_FRAME_CACHE_SIZE];
  SchroFrame *virt_frame1;
  SchroFrame *virt_frame2;
  void (*render_line) (SchroFrame *frame, void *dest, int component, int i);
  void *virt_priv;
  void *virt_priv2;

  int extension;
  int cache_offset[3];
  int is_upsampled;
  schro_bool upsample_done;
};

/* schroedinger/schrodomain.h:13-28.  Only `flags`, `alloc`, `free` are used here;
 * the slot cache is kept so the struct has the reference's size. */
#define SCHRO_MEMORY_DOMAIN_SLOTS 1000
struct _SchroMemoryDomain {
  void *mutex;
  unsigned int flags;
  void *(*alloc) (int size);
  void *(*alloc_2d) (int depth, int width, int height);
  void (*free) (void *ptr, int size);
  struct {
    unsigned int flags;
    void *ptr;
    int size;
    void *priv;
  } slots[SCHRO_MEMORY_DOMAIN_SLOTS];
};
#define SCHRO_MEMORY_DOMAIN_CPU 0x0001
#define SCHRO_MEMORY_DOMAIN_CUDA 0x0002
#define SCHRO_MEMORY_DOMAIN_PINNED 0x0100      /* new: page-locked host memory */
#define SCHRO_MEMORY_DOMAIN_SLOT_ALLOCATED 0x0001
#define SCHRO_MEMORY_DOMAIN_SLOT_IN_USE 0x0002

/* schroedinger/schrobitstream.h:79-86 */
typedef enum _SchroChromaFormat {
  SCHRO_CHROMA_444 = 0,
  SCHRO_CHROMA_422,
  SCHRO_CHROMA_420
} SchroChromaFormat;
#define SCHRO_CHROMA_FORMAT_H_SHIFT(format) (((format) == SCHRO_CHROMA_444)?0:1)
#define SCHRO_CHROMA_FORMAT_V_SHIFT(format) (((format) == SCHRO_CHROMA_420)?1:0)

/* schroedinger/schrovideoformat.h (struct _SchroVideoFormat) */
typedef struct _SchroVideoFormat {
  int index;
  int width;
  int height;
  SchroChromaFormat chroma_format;
  schro_bool interlaced;
  schro_bool top_field_first;
  int frame_rate_numerator;
  int frame_rate_denominator;
  int aspect_ratio_numerator;
  int aspect_ratio_denominator;
  int clean_width;
  int clean_height;
  int left_offset;
  int top_offset;
  int luma_offset;
  int luma_excursion;
  int chroma_offset;
  int chroma_excursion;
  int colour_primaries;
  int colour_matrix;
  int transfer_function;
  int interlaced_coding;
  int unused0;
  int unused1;
  int unused2;
} SchroVideoFormat;

#define SCHRO_LIMIT_TRANSFORM_DEPTH 6     /* schroedinger/schrolimits.h:21 */
#define SCHRO_LIMIT_BLOCK_SIZE 64         /* schroedinger/schrolimits.h:67 */

/* schroedinger/schroparams.h:18-29 */
typedef struct _SchroGlobalMotion {
  int b0, b1, a_exp, a00, a01, a10, a11, c_exp, c0, c1;
} SchroGlobalMotion;

/* schroedinger/schroparams.h:31-77 */
typedef struct _SchroParams {
  SchroVideoFormat *video_format;
  int is_noarith;
  int wavelet_filter_index;
  int transform_depth;
  int horiz_codeblocks[SCHRO_LIMIT_TRANSFORM_DEPTH + 1];
  int vert_codeblocks[SCHRO_LIMIT_TRANSFORM_DEPTH + 1];
  int codeblock_mode_index;
  int num_refs;
  int have_global_motion;
  int xblen_luma;
  int yblen_luma;
  int xbsep_luma;
  int ybsep_luma;
  int mv_precision;
  SchroGlobalMotion global_motion[2];
  int picture_pred_mode;
  int picture_weight_bits;
  int picture_weight_1;
  int picture_weight_2;
  int is_lowdelay;
  int n_horiz_slices;
  int n_vert_slices;
  int slice_bytes_num;
  int slice_bytes_denom;
  int quant_matrix[3 * SCHRO_LIMIT_TRANSFORM_DEPTH + 1];
  int iwt_chroma_width;
  int iwt_chroma_height;
  int iwt_luma_width;
  int iwt_luma_height;
  int x_num_blocks;
  int y_num_blocks;
  int x_offset;
  int y_offset;
} SchroParams;

/* schroedinger/schromotion.h:20-37 (20 bytes) */
typedef struct _SchroMotionVector {
  unsigned int pred_mode : 2;
  unsigned int using_global : 1;
  unsigned int split : 2;
  unsigned int unused : 3;
  unsigned int scan : 8;
  uint32_t metric;
  uint32_t chroma_metric;
  union {
    struct { int16_t dx[2]; int16_t dy[2]; } vec;
    struct { int16_t dc[3]; } dc;
  } u;
} SchroMotionVector;

/* schroedinger/schromotion.h:39-43 */
typedef struct _SchroMotionField {
  int x_num_blocks;
  int y_num_blocks;
  SchroMotionVector *motion_vectors;
} SchroMotionField;

/* schroedinger/schromotion.h:53-88.  Only src1, src2, motion_vectors and params
 * are inputs; the rest is the reference renderer's scratch state, kept for layout. */
typedef struct _SchroMotion {
  SchroFrame *src1;
  SchroFrame *src2;
  SchroMotionVector *motion_vectors;
  SchroParams *params;
  int ref_weight_precision;
  int ref1_weight;
  int ref2_weight;
  int mv_precision;
  int xoffset;
  int yoffset;
  int xbsep;
  int ybsep;
  int xblen;
  int yblen;
  SchroFrameData block;
  SchroFrameData alloc_block;
  SchroFrameData obmc_weight;
  SchroFrameData alloc_block_ref[2];
  SchroFrameData block_ref[2];
  int weight_x[SCHRO_LIMIT_BLOCK_SIZE];
  int weight_y[SCHRO_LIMIT_BLOCK_SIZE];
  int width;
  int height;
  int max_fast_x;
  int max_fast_y;
  schro_bool simple_weight;
  schro_bool oneref_noscale;
} SchroMotion;

/* schroedinger/schromotionest.h:22-31 */
typedef struct _SchroHierBm {
  int ref_count;
  int ref;
  int hierarchy_levels;
  SchroParams *params;
  SchroFrame **downsampled_src;
  SchroFrame **downsampled_ref;
  SchroMotionField **downsampled_mf;
  schro_bool use_chroma;
} SchroHierBm;

/* ---- library / domains -------------------------------------------------- */
void schro_init (void);                                   /* schroedinger/schro.c:23 */
/* new: the GPU this process's picture core runs on (default: the device current on the
 * first thread that calls into the library); worker threads inherit it */
void schro_b200_set_device (int device);
/* new: release the calling thread's stream / staging buffers / device-block pool (call
 * before a worker thread exits; long-lived workers never need it) */
void schro_b200_thread_release (void);
/* new: wait until the work the calling thread has left in flight on device frames is done
 * (calls that hand a result to the host already wait; this is for timing and for teardown) */
void schro_b200_thread_sync (void);
/* schroedinger/schrocuda.h:9 (schro_memory_domain_new_cuda) */
SchroMemoryDomain *schro_memory_domain_new_cuda (void);
SchroMemoryDomain *schro_memory_domain_new_pinned (void); /* new: cudaHostAlloc'd frames */
void schro_memory_domain_free (SchroMemoryDomain *domain); /* schroedinger/schrodomain.h:43 */
void *schro_memory_domain_alloc (SchroMemoryDomain *domain, int size);      /* :45 */
void schro_memory_domain_memfree (SchroMemoryDomain *domain, void *ptr);    /* :48 */

/* ---- frames (schroedinger/schroframe.h:96-160) ---------------------------- */
SchroFrame *schro_frame_new (void);
SchroFrame *schro_frame_new_and_alloc (SchroMemoryDomain *domain,
    SchroFrameFormat format, int width, int height);
SchroFrame *schro_frame_new_and_alloc_extended (SchroMemoryDomain *domain,
    SchroFrameFormat format, int width, int height, int extension);
SchroFrame *schro_frame_new_and_alloc_full (SchroMemoryDomain *domain,
    SchroFrameFormat format, int width, int height, int extension, int upsampled);
SchroFrame *schro_frame_ref (SchroFrame *frame);
void schro_frame_unref (SchroFrame *frame);
/* schroedinger/schroframe.c:699-724: a new frame of the same domain / format / size with the given
 * extension and layout, filled by schro_frame_convert */
SchroFrame *schro_frame_dup (SchroFrame *frame);
SchroFrame *schro_frame_dup_extended (SchroFrame *frame, int extension);
SchroFrame *schro_frame_dup_full (SchroFrame *frame, int extension, int is_upsampled);
/* schroedinger/schroframe.h: in-place shifts (schroframe.c:1238-1291) and the frame checksum (:1817-1861) */
void schro_frame_shift_left (SchroFrame *frame, int shift);
void schro_frame_shift_right (SchroFrame *frame, int shift);
void schro_frame_md5 (SchroFrame *frame, uint32_t *state);
/* schroedinger/schrocuda.h:14-16 / schrogpuframe.h:14-15: move a frame between domains */
void schro_frame_to_gpu (SchroFrame *dest, SchroFrame *src);
void schro_gpuframe_to_cpu (SchroFrame *dest, SchroFrame *src);

void schro_upsampled_frame_get_framedata (SchroFrame *upframe,
    SchroFrameData *fd, int up_index, int component);      /* schroframe.c:1917 */

/* ---- wavelets (schroedinger/schrowavelet.h:12-14) ------------------------- */
void schro_wavelet_transform_2d (SchroFrameData *fd, int type, int16_t *tmp);
void schro_wavelet_inverse_transform_2d (SchroFrameData *fd_dest,
    SchroFrameData *fd_src, int type, int16_t *tmp);
/* frame-granular drivers: schroedinger/schroframe.c:1192 and the decoder's
 * schro_decoder_inverse_iwt_transform (schroedinger/schrodecoder.c:1809), named as
 * the reference's own testsuite/cuda/cuda.c:96 calls it */
/* new: dequantise a frame of quantised coefficients in place on the GPU;
 * pairs = (quant_factor, quant_offset + 2) per codeblock, see sb2_dequantise in schro_b200.h.
 * Replaces the orc_dequantise_* calls of schro_decoder_decode_subband (schrodecoder.c:3395-3448). */
void schro_b200_frame_dequantise (SchroFrame *frame, SchroParams *params, const int32_t *pairs);
/* the same with an s16 source of quantised coefficients and an s32 destination (>8-bit streams whose
 * quantised values fit 16 bits): half the upload of the s32 coefficient frame */
void schro_b200_frame_dequantise_widen (SchroFrame *dest, SchroFrame *src, SchroParams *params,
    const int32_t *pairs);
void schro_frame_iwt_transform (SchroFrame *frame, SchroParams *params);
void schro_frame_inverse_iwt_transform (SchroFrame *frame, SchroParams *params);

/* ---- upsample / downsample / edge extension (schroedinger/schroframe.h:130-160) */
void schro_frame_downsample (SchroFrame *dest, SchroFrame *src);
/* schroedinger/schroframe.h:129-131 (schroframe.c:870, 1012, 1062): planar frames of equal chroma
 * format; packed formats and chroma resampling abort with a message */
void schro_frame_convert (SchroFrame *dest, SchroFrame *src);
void schro_frame_add (SchroFrame *dest, SchroFrame *src);
void schro_frame_subtract (SchroFrame *dest, SchroFrame *src);
void schro_frame_upsample_horiz (SchroFrameData *dest, SchroFrameData *src);
void schro_frame_upsample_vert (SchroFrameData *dest, SchroFrameData *src);
void schro_frame_mc_edgeextend (SchroFrame *frame);
void schro_upsampled_frame_upsample (SchroFrame *df);

/* ---- OBMC (schroedinger/schromotion.h:90-102) ------------------------------ */
SchroMotion *schro_motion_new (SchroParams *params, SchroFrame *ref1, SchroFrame *ref2);
void schro_motion_free (SchroMotion *motion);
void schro_motion_render (SchroMotion *motion, SchroFrame *dest,
    SchroFrame *addframe, int add, SchroFrame *output_frame);
void schro_motion_render_u8 (SchroMotion *motion, SchroFrame *dest,
    SchroFrame *addframe, int add, SchroFrame *output_frame);
/* schroedinger/schromotionref.c:245: the per-pixel renderer; schro_motion_render takes it when params->have_global_motion */
void schro_motion_render_ref (SchroMotion *motion, SchroFrame *dest, SchroFrame *addframe, int add,
    SchroFrame *output_frame);
void schro_motion_init_obmc_weight (SchroMotion *motion);

/* schroedinger/schrometric.h:16-53 */
#define SCHRO_LIMIT_METRIC_SCAN 42
#define SCHRO_METRIC_INVALID INT_MAX
typedef struct _SchroMetricInfo SchroMetricInfo;
struct _SchroMetricInfo {
  SchroFrame *frame;
  SchroFrame *ref_frame;
  int block_width[3];
  int block_height[3];
  int h_shift[3];
  int v_shift[3];
  int (*metric) (SchroMetricInfo *info, int ref_x, int ref_y, int dx, int dy);
  int (*metric_right) (SchroMetricInfo *info, int ref_x, int ref_y, int dx, int dy);
  int (*metric_bottom) (SchroMetricInfo *info, int ref_x, int ref_y, int dx, int dy);
  int (*metric_corner) (SchroMetricInfo *info, int ref_x, int ref_y, int dx, int dy);
};
typedef struct _SchroMetricScan {
  SchroFrame *frame;
  SchroFrame *ref_frame;
  int block_width;
  int block_height;
  int x, y;
  int ref_x, ref_y;
  int scan_width;
  int scan_height;
  int gravity_scale;
  int gravity_x, gravity_y;
  int use_chroma;
  /* output */
  uint32_t metrics[SCHRO_LIMIT_METRIC_SCAN * SCHRO_LIMIT_METRIC_SCAN];
  uint32_t chroma_metrics[SCHRO_LIMIT_METRIC_SCAN * SCHRO_LIMIT_METRIC_SCAN];
} SchroMetricScan;

/* ---- SAD primitives (schroedinger/schrometric.h:57-86) ---------------------- */
int schro_metric_absdiff_u8 (uint8_t *a, int a_stride, uint8_t *b, int b_stride,
    int width, int height);
int schro_metric_get (SchroFrameData *src1, SchroFrameData *src2, int width, int height);
int schro_metric_get_dc (SchroFrameData *src, int value, int width, int height);
int schro_metric_get_biref (SchroFrameData *fd, SchroFrameData *src1, int weight1,
    SchroFrameData *src2, int weight2, int shift, int width, int height);
/* the scan of one block (schroedinger/schrometric.c:31-214): window set-up (host arithmetic), the grid
 * of SADs on the GPU (one launch, one wait -- callers with many blocks use sb2_metric_scan, which takes
 * n scans per launch), arg-min with the reference's tie-break */
void schro_metric_scan_setup (SchroMetricScan *scan, int dx, int dy, int dist, int use_chroma);
void schro_metric_scan_do_scan (SchroMetricScan *scan);
int schro_metric_scan_get_min (SchroMetricScan *scan, int *dx, int *dy, uint32_t *chroma_metric);
/* 3-component block SAD of the candidate ranking (schroedinger/schrometric.c:332-414) */
void schro_metric_info_init (SchroMetricInfo *info, SchroFrame *frame, SchroFrame *ref_frame,
    int block_width, int block_height);
int schro_metric_fast_block (SchroMetricInfo *info, int x, int y, int dx, int dy);

/* ---- hierarchical block matching (schroedinger/schromotionest.h:112-120) ---- */
SchroMotionField *schro_motion_field_new (int x_num_blocks, int y_num_blocks);
void schro_motion_field_free (SchroMotionField *field);
/* schro_hbm_new (schroedinger/schrohierbm.c:25-64) takes a SchroEncoderFrame; this is
 * the same constructor with the five fields it reads passed explicitly (INTEGRATION.md
 * shows the one-line shim). frames[0] = full-resolution picture, frames[i] = level i. */
SchroHierBm *schro_hbm_new_from_frames (SchroParams *params, int ref,
    int hierarchy_levels, schro_bool use_chroma,
    SchroFrame **src_frames, SchroFrame **ref_frames);
SchroHierBm *schro_hbm_ref (SchroHierBm *schro_hbm);
void schro_hbm_unref (SchroHierBm *schro_hbm);
void schro_hbm_scan (SchroHierBm *schro_hbm);
void schro_hierarchical_bm_scan_hint (SchroHierBm *schro_hbm, int shift, int h_range);
SchroMotionField *schro_hbm_motion_field (SchroHierBm *schro_hbm, int level);

/* ---- rough (bigblock) motion search (schroedinger/schromotionest.h:50-76, schroroughmotion.c) ----
 * SchroRoughME as the reference lays it out; callers read motion_fields[1] / [2] directly
 * (schromotionest.c:535, 742; schroglobalest.c:25), so every level function brings its field
 * back to the host before it returns.  schro_rough_me_new (SchroEncoderFrame *, SchroEncoderFrame *)
 * reads the encoder structure; the library exports the constructor with those things passed
 * explicitly and compat/schro_rough_me_new.c is the reference-side half (as for schro_hbm_new). */
#define SCHRO_MAX_HIER_LEVELS 8                         /* schroedinger/schromotionest.h:20 */
struct _SchroEncoderFrame;
typedef struct _SchroRoughME {
  struct _SchroEncoderFrame *encoder_frame;
  struct _SchroEncoderFrame *ref_frame;
  SchroMotionField *motion_fields[SCHRO_MAX_HIER_LEVELS];
} SchroRoughME;
/* frames[0] = the filtered picture, frames[i] = pyramid level i, i <= levels
 * (encoder->downsample_levels); ref = which of ref_frame[0] / [1] `ref_frame` is */
SchroRoughME *schro_rough_me_new_from_frames (struct _SchroEncoderFrame *frame,
    struct _SchroEncoderFrame *ref_frame, SchroParams *params, int ref, int levels,
    SchroFrame **src_frames, SchroFrame **ref_frames);
void schro_rough_me_free (SchroRoughME *rme);
void schro_rough_me_heirarchical_scan (SchroRoughME *rme);
void schro_rough_me_heirarchical_scan_nohint (SchroRoughME *rme, int shift, int distance);
void schro_rough_me_heirarchical_scan_hint (SchroRoughME *rme, int shift, int distance);


/* ---- sub-pel refinement (schroedinger/schromotionest.c:246-355) ----
 * schro_encoder_motion_predict_subpel_deep (SchroMe *) reads a structure private to schromotionest.c
 * through five accessors; this is the same function with those five things passed explicitly
 * (compat/schro_subpel_deep.c keeps the reference's symbol on top of it):
 *   params             schro_me_params: xbsep_luma, ybsep_luma, x/y_num_blocks, num_refs, mv_precision
 *   lambda             schro_me_lambda
 *   orig_frame         schro_me_src: the (filtered) source picture, u8
 *   upsampled_refs[r]  schro_me_ref (me, r): the upsampled reference pictures (upsampled here if not yet)
 *   subpel_mfs[r]      schro_me_subpel_mf (me, r): refined in place */
void schro_b200_motion_predict_subpel_deep (SchroParams *params, double lambda, SchroFrame *orig_frame,
    SchroFrame **upsampled_refs, SchroMotionField **subpel_mfs);

/* ---- split-2 pass of the mode decision (schroedinger/schromotionest.c:1601-1802, 1511-1523) ----
 * schro_do_split2 (static, takes the private SchroMe) + schro_motion_copy_to for every superblock in raster
 * order: what schro_mode_decision (:2587-2685) produces when the split-1 / split-0 candidates never win.
 * The things schro_do_split2 reads through the SchroMe accessors are passed explicitly:
 *   params, lambda, orig_frame, upsampled_refs[r]   as for the sub-pel refinement above
 *   split2_mfs[r]      schro_me_split2_mf (me, r): the references' sub-pel fields (read only)
 *   motion             schro_me_motion: motion->motion_vectors receives the decided blocks
 *   sb_error / sb_entropy   SchroBlock.error / .entropy per superblock ((x_num_blocks / 4) * (y_num_blocks / 4)
 *                      ints, raster order; either may be NULL); SchroBlock.score = entropy + lambda * error */
void schro_b200_mode_decision_split2 (SchroParams *params, double lambda, SchroFrame *orig_frame,
    SchroFrame **upsampled_refs, SchroMotionField **split2_mfs, SchroMotion *motion, int *sb_error, int *sb_entropy);

/* ---- inverse transform + combine (schroedinger/schrodecoder.c:1809-1853 + 2054-2061) ----
 * new: schro_frame_inverse_iwt_transform (frame, params) + schro_frame_shift_right (frame, shift) +
 * schro_frame_convert (output, frame) -- the decoder's path for a non-reference intra picture -- as ONE call;
 * the 8-bit picture is written by the last wavelet level, the coefficient frame is left as it was. */
void schro_b200_frame_inverse_iwt_combine (SchroFrame *output, SchroFrame *frame, SchroParams *params, int shift);

/* new: a batch of n independent low-delay intra pictures (one `params`, n slice buffers of `length` bytes, n u8
 * output frames of one layout in host memory) decoded with one launch per stage and one wait:
 * slices -> coefficients -> inverse transform with the shift / conversion to 8 bits fused in.  The batched
 * counterpart of schro_b200_decode_lowdelay_transform_data + schro_b200_frame_inverse_iwt_combine. */
void schro_b200_decode_lowdelay_pictures (SchroParams *params, int n, const uint8_t *const *data, int length,
    SchroFrame *const *outputs, int is_s32, int shift);

/* ---- low-delay slices (schroedinger/schrolowdelay.c:745-761) ----
 * schro_decoder_decode_lowdelay_transform_data (SchroPicture *) with the three things it reads from the picture
 * passed explicitly (compat/schro_lowdelay.c keeps the reference's symbol): picture->params,
 * picture->lowdelay_buffer->data / ->length, picture->transform_frame (s16 or s32, host or CUDA domain).
 * The compressed slices are what goes to the GPU; the dequantised, DC-predicted coefficients are left in the
 * frame, ready for schro_frame_inverse_iwt_transform. */
void schro_b200_decode_lowdelay_transform_data (SchroParams *params, const uint8_t *data, int length,
    SchroFrame *transform_frame);

#ifdef __cplusplus
}
#endif
#endif
