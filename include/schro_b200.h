/*
 * schro_b200.h -- device-level C ABI of the B200 picture core.
 *
 * This is the thin CUDA layer that the C host code (schroedinger_b200/host/,
 * mirroring the reference's own API) and any FFI binding call into.  Plain
 * pointers and sizes only; every pointer marked "device" must be GPU memory,
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * All functions return 0 on success or a negative sb2 error / cudaError code;
 * sb2_last_error() gives a message.  Nothing here falls back to the CPU.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to the dschleef/schroedinger tree).
 */
#ifndef SCHRO_B200_H
#define SCHRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB2_MAX_COMPONENTS 4

enum {
  SB2_OK = 0,
  SB2_ERR_ARG = -1,
  SB2_ERR_WORKSPACE = -2,
  SB2_ERR_CUDA = -3,
  SB2_ERR_UNSUPPORTED = -4
};

/* wavelet filter ids == SchroWaveletIndex, schroedinger/schrobitstream.h:124-132 */
enum {
  SB2_WAVELET_DESLAURIERS_DUBUC_9_7 = 0,
  SB2_WAVELET_LE_GALL_5_3 = 1,
  SB2_WAVELET_DESLAURIERS_DUBUC_13_7 = 2,
  SB2_WAVELET_HAAR_0 = 3,
  SB2_WAVELET_HAAR_1 = 4,
  SB2_WAVELET_FIDELITY = 5,
  SB2_WAVELET_DAUBECHIES_9_7 = 6
};

const char *sb2_last_error (void);
int sb2_version (void);
/* number of kernels this library has launched in the calling process */
unsigned long long sb2_launch_count (void);

/* Per-launch device timing (CUDA events on the launching stream): enable, run,
 * synchronise the stream(s), then read (tag, milliseconds, algorithmic bytes). */
void sb2_profile_enable (int on);
void sb2_profile_reset (void);
int sb2_profile_count (void);
int sb2_profile_get (int index, char *tag, int tag_len, float *ms, double *bytes);

/* ------------------------------------------------------------------------
 * Picture slabs.  A slab is `count` pictures laid out `picture_pitch` bytes
 * apart in one device allocation; every picture has `ncomp` component planes
 * at fixed byte offsets (the layout schro_frame_new_and_alloc_full gives one
 * frame, schroedinger/schroframe.c:60-191, repeated at a fixed pitch so that a
 * whole batch is one launch).  offset[] points at pixel (0,0) of the plane
 * (for extended / upsampled frames: of phase 0), stride[] is in bytes.
 * A call takes at most 65535 pictures (they ride on the launch grid's y dimension).
 * ---------------------------------------------------------------------- */
typedef struct {
  void *base;                       /* device */
  size_t picture_pitch;             /* bytes between consecutive pictures */
  int count;                        /* pictures in the slab */
  int ncomp;                        /* 1..SB2_MAX_COMPONENTS */
  size_t offset[SB2_MAX_COMPONENTS];
  int stride[SB2_MAX_COMPONENTS];
  int width[SB2_MAX_COMPONENTS];
  int height[SB2_MAX_COMPONENTS];
} sb2_slab;

/* ---- wavelets --------------------------------------------------------- */

/* Workspace (device bytes) needed by sb2_iwt_forward / sb2_iwt_inverse for
 * this slab shape.  in_place != 0 when dst aliases src. */
size_t sb2_iwt_workspace_bytes (const sb2_slab *slab, int is_s32, int depth,
    int in_place);

/* Multi-level forward transform of every component of every picture:
 * replaces the level loop of schro_frame_iwt_transform
 * (schroedinger/schroframe.c:1192-1228) / schro_encoder_iwt_transform
 * (schroedinger/schroencoder.c:2391-2427), i.e. `depth` calls of
 * schro_wavelet_transform_2d (schroedinger/schrowaveletorc.c:60) per component.
 * width/height of each component must be multiples of 1<<depth.
 * dst may alias src (same slab) -- costs one extra device copy. */
int sb2_iwt_forward (const sb2_slab *src, const sb2_slab *dst, int is_s32,
    int filter, int depth, void *workspace, size_t workspace_bytes,
    void *stream);

/* Multi-level inverse: replaces schro_decoder_inverse_iwt_transform
 * (schroedinger/schrodecoder.c:1809-1853) / schro_encoder_inverse_iwt_transform
 * (schroedinger/schroencoder.c:2645-2689), i.e. `depth` calls of
 * schro_wavelet_inverse_transform_2d (schroedinger/schrowaveletorc.c:121). */
int sb2_iwt_inverse (const sb2_slab *src, const sb2_slab *dst, int is_s32,
    int filter, int depth, void *workspace, size_t workspace_bytes,
    void *stream);

/* A level is run by register-chunk kernels where a component's size allows (half-width and
 * half-height multiples of 8 or 16) and by a generic tile kernel otherwise.  on != 0 sends every
 * component to the generic kernel (tests run both on the same input; also SB2_IWT_GENERIC=1). */
/* Inverse transform with the decoder's combine step fused into its last level: what
 * schro_decoder_x_combine does for a non-reference intra picture (schroedinger/schrodecoder.c:2054-2061) --
 * schro_frame_shift_right (frame, shift) when the stream is deeper than the output, then schro_frame_convert
 * to the 8-bit picture (crop included) -- happens in the epilogue of the level-0 kernel, so the picture leaves
 * the GPU's registers as u8 and the s16 / s32 plane is never written.  dst_u8: u8 slab of the same component
 * count, each component no larger than the transform's area.  Needs planes whose half sizes are multiples of 8
 * (the register-chunk kernels); otherwise SB2_ERR_UNSUPPORTED: call sb2_iwt_inverse + sb2_frame_shift +
 * sb2_frame_convert.  Workspace as for sb2_iwt_inverse (not in place). */
int sb2_iwt_inverse_convert (const sb2_slab *src, const sb2_slab *dst_u8, int is_s32, int filter, int depth,
    int shift, void *workspace, size_t workspace_bytes, void *stream);

void sb2_iwt_force_generic (int on);
/* on != 0: the inverse transform runs its last two levels (1 and 0) as ONE fused launch -- level 1's
 * output stays in shared memory, never in HBM -- for the filters DD 9/7, DD 13/7 and Daubechies 9/7 on
 * planes the register-chunk kernels cover.  Off by default: measured on B200 the fused launch moves a
 * quarter less DRAM traffic but takes 1.36 ms against 1.00 ms per 32 2160p pictures, because the halo of
 * level 1 is recomputed per tile on an issue-bound kernel (tools/time_wavelet_fused.py).
 * The environment variable SB2_IWT_FUSED=1 turns it on as well.  Results are identical either way. */
void sb2_iwt_enable_fused (int on);

/* ---- reference-frame preparation (u8) ------------------------------------ */

/* Replicate the picture edge into the `extension` border pixels of every plane:
 * schro_frame_mc_edgeextend (schroedinger/schroframe.c:1986-1997).  For frames in the
 * 4-phase "upsampled" layout `phase` selects the phase plane (0 = the picture). */
int sb2_mc_edgeextend (const sb2_slab *frames, int extension, int phase, void *stream);

/* Fill half-pel phases 1..3 (and all their borders) of upsampled frames whose phase 0 is
 * already edge-extended: schro_upsampled_frame_upsample (schroedinger/schroframe.c:2000-2030),
 * i.e. schro_frame_upsample_vert / _horiz (:1557-1645) plus the border passes.
 * stride[] is the 4-phase row stride; phase p starts (stride>>2)*p bytes into each row. */
int sb2_upsample (const sb2_slab *frames, int extension, void *stream);

/* sb2_mc_edgeextend (phase 0) + sb2_upsample in one launch: what the decoder does to every
 * reconstructed reference picture (schroedinger/schrodecoder.c:2068-2087, 2120-2141). */
int sb2_edgeextend_upsample (const sb2_slab *frames, int extension, void *stream);

/* Two kernels implement the upsampler: 1 = whole words and dp4a (needs 4-byte aligned rows in every
 * phase plane and an extension that is a multiple of 4: the codec's layouts), 2 = one pixel at a time
 * (anything).  0 picks by alignment; tests force each (also: environment variable SB2_UPSAMPLE_KERNEL). */
void sb2_upsample_force_kernel (int which);
/* which of the two the calling thread's last sb2_upsample / sb2_edgeextend_upsample launched */
int sb2_upsample_last_kernel (void);

/* dst = half-resolution src, (6,26,26,6) twice with an 8-bit intermediate:
 * schro_frame_downsample (schroedinger/schroframe.c:1505-1513).  dst sizes must be
 * (w+1)/2 x (h+1)/2 per component. */
int sb2_downsample (const sb2_slab *src, const sb2_slab *dst, void *stream);

/* sb2_downsample followed by sb2_mc_edgeextend (dst, dst_extension) in one launch: one level
 * of schro_encoder_frame_downsample (schroedinger/schroanalysis.c:17-27). */
int sb2_downsample_edgeextend (const sb2_slab *src, const sb2_slab *dst, int dst_extension,
    void *stream);

/* As for the upsampler: 1 = whole words (4-byte aligned rows on both sides), 2 = one pixel at a time;
 * 0 picks by alignment (also: environment variable SB2_DOWNSAMPLE_KERNEL). */
void sb2_downsample_force_kernel (int which);
int sb2_downsample_last_kernel (void);

/* One-direction half-pel filter of a bare plane, no borders:
 * schro_frame_upsample_horiz / schro_frame_upsample_vert (schroedinger/schroframe.c:1557, 1612) */
int sb2_upsample_plane_1d (uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride,
    int width, int height, int vertical, void *stream);

/* ---- OBMC motion compensation ------------------------------------------- */

/* The subset of SchroParams (schroedinger/schroparams.h:31-77) the renderer reads */
typedef struct {
  int xbsep, ybsep, xblen, yblen;       /* luma block separation / length */
  int x_num_blocks, y_num_blocks;
  int mv_precision;                     /* 0..3 */
  int picture_weight_1, picture_weight_2, picture_weight_bits;
  int chroma_h_shift, chroma_v_shift;   /* applied to components >= 1 (sizes and vectors) */
} sb2_obmc_params;

/* schro_motion_render / schro_motion_render_u8 (schroedinger/schromotion.c:95,
 * schroedinger/schromotion8.c:700) for every picture of the slabs.
 *   motion_vectors  device array of SchroMotionVector (20 bytes each,
 *                   schroedinger/schromotion.h:20-37), x_num_blocks*y_num_blocks per picture,
 *                   consecutive pictures `mv_picture_pitch` VECTORS apart
 *   ref0 / ref1     upsampled (4-phase), edge-extended (extension >= 32) u8 references;
 *                   ref1 may be NULL when no block uses it
 *   acc             optional s16 slab receiving what the reference leaves in `dest`.  Its plane
 *                   sizes are the rendered area (schromotion8.c:722-751 takes them from dest);
 *                   residual / out may be larger (an iwt-padded addframe), never smaller.
 *                   Without acc the area is that of out (add) or residual (subtract).
 *   add != 0        out(u8) = clamp(residual + ((acc+32)>>6)); residual s16 or s32
 *   add == 0        t = (acc-8160)>>6; acc := t; residual(s16) -= t */
int sb2_obmc_render (const sb2_obmc_params *params, const void *motion_vectors,
    size_t mv_picture_pitch, const sb2_slab *ref0, const sb2_slab *ref1, const sb2_slab *acc,
    const sb2_slab *residual, int residual_is_s32, int add, const sb2_slab *out, void *stream);

/* schro_motion_render_ref (schroedinger/schromotionref.c:245-330): the per-pixel renderer the reference
 * switches to whenever the picture has global motion (schroedinger/schromotion.c:113-121).  Same arguments as
 * sb2_obmc_render, plus global_motion: 2 x 10 ints (SchroGlobalMotion's fields b0 b1 a_exp a00 a01 a10 a11 c_exp c0
 * c1 for each reference; NULL: all zero) used by the blocks whose using_global bit is set.  Semantics differ
 * from the block renderer where the reference's do: acc receives clamp (prediction, 0, 255) - 128 and the
 * prediction is clamped before the residual (s16 only) is added or subtracted.  Blocks at most twice their
 * separation (a pixel is covered by at most four). */
int sb2_obmc_render_ref (const sb2_obmc_params *params, const int *global_motion, const void *motion_vectors,
    size_t mv_picture_pitch, const sb2_slab *ref0, const sb2_slab *ref1, const sb2_slab *acc,
    const sb2_slab *residual, int add, const sb2_slab *out, void *stream);

/* Three kernels implement the renderer: 1 = one block per warp pass out of TMA-staged reference regions
 * (blocks of at most 32 row x 8-pixel lanes, 32-pixel borders: every Dirac preset), 2 = block-major
 * scatter out of global memory with shared-memory atomics (the default where it applies: measured
 * faster), 3 = one thread per pixel (any geometry).  0 picks by geometry; tests force each (also:
 * environment variable SB2_OBMC_KERNEL). */
void sb2_obmc_force_kernel (int which);
/* which of the three the calling thread's last sb2_obmc_render launched */
int sb2_obmc_last_kernel (void);

/* ---- SAD / hierarchical block matching ----------------------------------- */

typedef struct {
  int xbsep, ybsep;                     /* luma block size used for matching (= separation) */
  int x_num_blocks, y_num_blocks;
  int ref_index;                        /* which of dx[2]/dy[2] is written (SchroHierBm.ref) */
  int use_chroma;                       /* encoder->enable_chroma_me (4:2:0 only) */
  int chroma_h_shift, chroma_v_shift;
} sb2_hbm_params;

size_t sb2_hbm_workspace_bytes (int x_num_blocks, int y_num_blocks, int count);

/* Two kernels implement a level: the skewed multi-row wavefront (hbm_wave.cu) for 8x8 blocks,
 * 4:2:0 and a luma-only scan -- the codec's defaults -- and a generic one for everything else.
 * on != 0 forces the generic kernel for every geometry (tests compare the two); the
 * environment variable SB2_HBM_GENERIC=1 does the same. */
void sb2_hbm_force_generic (int on);

/* One level of hierarchical block matching for `count` independent (picture, reference)
 * pairs: schro_hierarchical_bm_scan_hint (schroedinger/schrohierbm.c:174-383), including
 * the candidate ranking (schro_metric_fast_block, schroedinger/schrometric.c:332-414) and
 * the scan (schro_metric_scan_setup / _do_scan / _get_min, :31-214).
 *   src_level / ref_level   pyramid level `shift` of both pictures: three u8 components,
 *                           edge-extended by `extension` (32 at level 0, max(xbsep,ybsep) above)
 *   parent_field            device field of level shift+1 (NULL for the coarsest level)
 *   out_field               device field (x_num_blocks*y_num_blocks SchroMotionVector per
 *                           pair, pairs `field_picture_pitch` vectors apart); every entry is
 *                           initialised as schro_motion_field_set does, entries on the
 *                           1<<shift grid receive the search result */
int sb2_hbm_scan_hint (const sb2_hbm_params *params, const sb2_slab *src_level,
    const sb2_slab *ref_level, int extension, int shift, int h_range,
    const void *parent_field, void *out_field, size_t field_picture_pitch,
    void *workspace, size_t workspace_bytes, void *stream);

/* The rough ("bigblock") motion search, one pyramid level for `count` independent (picture,
 * reference) pairs.  Luma only; sb2_hbm_params.use_chroma / chroma shifts are
 * ignored.  Fields as in sb2_hbm_scan_hint; every entry is initialised as schro_motion_field_set
 * (mf, 0, 1) does, entries on the 1<<shift grid receive the search result.
 * sb2_rough_scan_nohint = schro_rough_me_heirarchical_scan_nohint (schroedinger/schroroughmotion.c:62-143):
 *   full search of +-distance around every block (no dependency between blocks).
 * sb2_rough_scan_hint = schro_rough_me_heirarchical_scan_hint (:145-300): candidates zero / four
 *   nearest parents (field of level shift+1) / left, up, up-left of this level, then a +-distance scan.
 * A block that lies entirely outside its level's frame keeps the zero vector at a hint level (the
 * reference reads stale memory there, oracle/oracle_rough.c). */
size_t sb2_rough_workspace_bytes (int x_num_blocks, int y_num_blocks, int count);
/* the full search stages each block's reference window in shared memory; on != 0 makes it read every row
 * segment from global memory instead (the path partial blocks and unaligned layouts take) -- tests run both */
void sb2_rough_force_unstaged (int on);
int sb2_rough_scan_nohint (const sb2_hbm_params *params, const sb2_slab *src_level,
    const sb2_slab *ref_level, int extension, int shift, int distance, void *out_field,
    size_t field_picture_pitch, void *stream);
int sb2_rough_scan_hint (const sb2_hbm_params *params, const sb2_slab *src_level,
    const sb2_slab *ref_level, int extension, int shift, int distance, const void *parent_field,
    void *out_field, size_t field_picture_pitch, void *workspace, size_t workspace_bytes, void *stream);

/* Sub-pel refinement of one reference's motion field, in place, for `count` independent pictures
 * (SURVEY.md 8f rank 3): schro_encoder_motion_predict_subpel_deep
 * (schroedinger/schromotionest.c:246-355) for one value of `ref`.
 *   orig        the source pictures (u8; the luma plane is used); orig_extension = that frame's
 *               `extension`, which only enters the reference's probe range test (:306-312)
 *   upref       the upsampled reference pictures (four half-pel phases, edge-extended by
 *               upref_extension = 32 as schro_upsampled_frame_upsample leaves them)
 *   field       device field (x_num_blocks*y_num_blocks SchroMotionVector per picture): on entry the
 *               integer-pel vectors and metrics (the level-0 field of hierarchical block matching,
 *               schroedinger/schroencoder.c:2306-2315), on return vectors in units of
 *               2^-mv_precision pixels and the SADs of the probes that won
 *   lambda      schro_me_lambda: score = entropy + lambda * SAD in double precision */
typedef struct {
  int xblen, yblen;                     /* params->xbsep_luma, ybsep_luma */
  int x_num_blocks, y_num_blocks;       /* at most 1024 block rows */
  int mv_precision;                     /* passes mvprec = 1 .. mv_precision */
  int ref_index;
  int orig_extension;
  double lambda;
} sb2_subpel_params;
size_t sb2_subpel_workspace_bytes (int x_num_blocks, int y_num_blocks, int count);
/* full 8 x 8 blocks take a word-wide probe path; on != 0 sends every block through the per-pixel path (tests run both) */
void sb2_subpel_force_generic (int on);
int sb2_subpel_refine (const sb2_subpel_params *params, const sb2_slab *orig, const sb2_slab *upref,
    int upref_extension, void *field, size_t field_picture_pitch, void *workspace, size_t workspace_bytes,
    void *stream);

/* The split-2 pass of the encoder's mode decision for every picture of the slabs: schro_do_split2 +
 * schro_motion_copy_to (schroedinger/schromotionest.c:1601-1802, 1511-1523) for every superblock, the first
 * step of schro_mode_decision (:2587-2685).  Per block the candidates are each reference's sub-pel vector
 * (luma SAD from the field + chroma SADs, schro_get_split2_metric :1527-1594), both together
 * (schro_metric_get_biref) and a DC block (schro_block_average :481-516); cost = entropy + lambda * error.
 *   orig            the source picture, three u8 components
 *   upref0/upref1   the upsampled (4-phase), edge-extended references, three components (upref1 and field1
 *                   may be NULL with one reference)
 *   field0/field1   device fields of the two references at mv_precision (the sub-pel refinement's output,
 *                   copied into split2_mf at schroedinger/schroencoder.c:2340-2349), pictures
 *                   `field_picture_pitch` vectors apart
 *   motion          receives x_num_blocks * y_num_blocks decided SchroMotionVectors per picture
 *   sb_error / sb_entropy   SchroBlock.error / .entropy of every superblock, (x_num_blocks / 4) *
 *                   (y_num_blocks / 4) ints per picture, pictures packed; SchroBlock.score =
 *                   entropy + lambda * error
 * Results are the reference's bit for bit, including where its behaviour is accidental (a single-reference
 * winner records the luma metric only; with mv_precision >= 2 the bi-reference metrics are taken on a scratch
 * block the three components share) -- see csrc/split2.cu. */
typedef struct {
  int xblen, yblen;                     /* params->xbsep_luma, ybsep_luma */
  int x_num_blocks, y_num_blocks;       /* multiples of 4; at most 1024 block rows */
  int mv_precision;                     /* 0..3 */
  int num_refs;                         /* 1 or 2 */
  int chroma_h_shift, chroma_v_shift;
  int orig_extension;                   /* extension of the source frame (enters the bi-reference range test) */
  double lambda;                        /* schro_me_lambda */
} sb2_split2_params;
size_t sb2_split2_workspace_bytes (int x_num_blocks, int y_num_blocks, int count);
/* full 8 x 8 blocks take a fixed lane map; on != 0 sends every block through the per-pixel path (tests run both) */
void sb2_split2_force_generic (int on);
int sb2_split2_decide (const sb2_split2_params *params, const sb2_slab *orig, const sb2_slab *upref0,
    const sb2_slab *upref1, int upref_extension, const void *field0, const void *field1,
    size_t field_picture_pitch, void *motion, size_t motion_picture_pitch, int *sb_error, int *sb_entropy,
    void *workspace, size_t workspace_bytes, void *stream);

/* schro_metric_absdiff_u8 (schroedinger/schrometric.c:10-29) for `n` independent block
 * pairs: sad[i] = SAD(a + a_offset[i], b + b_offset[i]) over width x height. */
int sb2_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b, int b_stride,
    const int64_t *a_offset, const int64_t *b_offset, int n, int width, int height,
    uint32_t *sad, void *stream);

/* The metric-scan entry points on their own, batched: n independent scans / block SADs per launch.
 * sb2_metric_scan = schro_metric_scan_do_scan (schroedinger/schrometric.c:31-116):
 *   metrics[k][i * scan_height + j] = luma SAD of the block at (x, y) against (ref_x + i, ref_y + j);
 *   chroma_metrics likewise (zero unless use_chroma; 42 x 42 entries per scan as SchroMetricScan,
 *   schroedinger/schrometric.h:38-53).  The window comes from schro_metric_scan_setup (:174-214), which is
 *   host arithmetic (host layer).
 * sb2_metric_block_sad3 = schro_metric_fast_block / schro_metric_block_sad_slow (:332-414): Y + U + V SAD
 *   of the block at (x, y) against (x + dx, y + dy), INT_MAX when a block leaves frame +- extension. */
typedef struct {
  int picture;                          /* index into the slabs */
  int x, y, block_width, block_height;
  int ref_x, ref_y, scan_width, scan_height;
} sb2_metric_scan_desc;
typedef struct {
  int picture;
  int x, y, dx, dy;
} sb2_metric_block_desc;
int sb2_metric_scan (const sb2_slab *src, const sb2_slab *ref, int chroma_h_shift, int chroma_v_shift,
    int use_chroma, const sb2_metric_scan_desc *descs /* device */, int n, uint32_t *metrics /* device */,
    uint32_t *chroma_metrics /* device or NULL */, void *stream);
int sb2_metric_block_sad3 (const sb2_slab *src, const sb2_slab *ref, int extension, int block_width,
    int block_height, int chroma_h_shift, int chroma_v_shift, const sb2_metric_block_desc *descs /* device */,
    int n, int *metric /* device */, void *stream);

/* schro_metric_get_dc (schroedinger/schrometric.c:253) and schro_metric_get_biref (:272);
 * one block, result in device memory */
int sb2_sad_dc_u8 (const uint8_t *a, int a_stride, int value, int width, int height, int *sad,
    void *stream);
int sb2_sad_biref_u8 (const uint8_t *a, int a_stride, const uint8_t *src1, int src1_stride,
    int weight1, const uint8_t *src2, int src2_stride, int weight2, int shift, int width,
    int height, int *sad, void *stream);

/* ------------------------------------------------------------------------
 * Combine / convert glue between the stages.
 * Depth codes: 0 = u8, 1 = s16, 2 = s32.  Planar slabs of equal component count.
 *
 * sb2_frame_convert: the planar, equal-chroma-format part of schro_frame_convert
 *   (schroedinger/schroframe.c:870-978): depth conversion with the Orc programs' wrap and
 *   saturation points (schroorc-dist.c: orc_offsetconvert_* / orc_convert_*), then crop or
 *   edge extension (schrovirtframe.c:1824-1960): dst(x,y) = conv(src(min(x,sw-1), min(y,sh-1))).
 * sb2_frame_add: schro_frame_add / schro_frame_subtract (schroedinger/schroframe.c:1012-1182):
 *   dst (s16) +-= src (u8 or s16) over the common area of each component, 16-bit wrap.
 * ---------------------------------------------------------------------- */
int sb2_frame_convert (const sb2_slab *src, int src_depth, const sb2_slab *dst, int dst_depth,
    void *stream);
int sb2_frame_add (const sb2_slab *dst, const sb2_slab *src, int src_depth, int subtract,
    void *stream);
/* schro_frame_shift_left / schro_frame_shift_right (schroedinger/schroframe.c:1238-1291), in place:
 * right == 0: x << shift with 16-bit wrap (s16 only); right != 0: (x + ((1 << shift) >> 1)) >> shift,
 * the add wrapping at the sample width (depth 1: s16, 2: s32). */
int sb2_frame_shift (const sb2_slab *frames, int depth, int shift, int right, void *stream);

/* ------------------------------------------------------------------------
 * Dequantisation of a coefficient frame in place: what
 * schro_decoder_decode_subband applies codeblock by codeblock
 * (schroedinger/schrodecoder.c:3395-3448, 3559-3576) with orc_dequantise_s16_ip_2d /
 * _s32_ip_2d (schroedinger/schroorc.orc:1154-1168, 2148-2162), on the in-place subband layout
 * (schroedinger/schroparams.c:319-352).
 * quant (device): per picture `quant_picture_pitch` pairs apart; for every component, for band
 * index 0..3*depth (schro_subband_get_position order), codeblock rows then columns: int32 pairs
 * (quant_factor, quant_offset + 2) -- the values the decoder passes to the Orc kernel, e.g.
 * schro_table_quant[i], schro_table_offset_1_2[i] + 2.  sb2_dequant_table_pairs = pairs per picture.
 * horiz/vert_codeblocks: [0] for the LL band, [i + 1] for the bands of level i (SchroParams).
 * ---------------------------------------------------------------------- */
#define SB2_DEQUANT_MAX_LEVELS 6
typedef struct {
  int transform_depth;
  int horiz_codeblocks[SB2_DEQUANT_MAX_LEVELS + 1];
  int vert_codeblocks[SB2_DEQUANT_MAX_LEVELS + 1];
} sb2_dequant_params;
size_t sb2_dequant_table_pairs (const sb2_dequant_params *p, int ncomp);
int sb2_dequantise (const sb2_slab *coeffs, int is_s32, const sb2_dequant_params *p,
    const int32_t *quant, size_t quant_picture_pitch, void *stream);

/* The same for a >8-bit stream whose quantised coefficients fit 16 bits (what an entropy decoder
 * produces for any practical quantiser): the input slab is s16, the arithmetic and the output slab
 * are s32 -- exactly orc_dequantise_s32_ip_2d (schroedinger/schroorc.orc:2148-2162) on the
 * sign-extended values -- so the host uploads half the bytes of the s32 coefficient frame. */
int sb2_dequantise_widen (const sb2_slab *quantised_s16, const sb2_slab *coeffs_s32,
    const sb2_dequant_params *p, const int32_t *quant, size_t quant_picture_pitch, void *stream);

/* ------------------------------------------------------------------------
 * Low-delay slice decoder (second half): schro_decoder_decode_lowdelay_transform_data
 * (schroedinger/schrolowdelay.c:99-761) + DC prediction of the LL band (schroedinger/schrodecoder.c:3219-3277)
 * for `count` pictures: the compressed slices go in, dequantised coefficients in the in-place subband
 * layout come out (then sb2_iwt_inverse).  One thread per slice.
 *   slices          device: per picture the slices back to back (SchroPicture.lowdelay_buffer->data),
 *                   pictures `picture_pitch` bytes apart, `picture_bytes` valid bytes each
 *   coeffs          three-component coefficient frames (s16, or s32 with is_s32), sizes = the iwt sizes
 *   table_quant / table_offset   the reference's schro_table_quant / schro_table_offset_1_2 (61 entries)
 * Which dequantiser runs follows the reference's dispatcher (:745-761): the 16-bit Orc program for s16
 * frames whose chroma LL band divides by the slice grid, the plain-C one otherwise. */
typedef struct {
  int transform_depth;
  int n_horiz_slices, n_vert_slices;
  int slice_bytes_num, slice_bytes_denom;
  int quant_matrix[1 + 3 * SB2_DEQUANT_MAX_LEVELS];
  uint32_t table_quant[61], table_offset[61];
} sb2_lowdelay_params;
/* the slice kernel stages each CTA's slices in shared memory; on != 0 makes it read them from global memory
 * (what slices too large to stage get) -- tests run both */
void sb2_lowdelay_force_unstaged (int on);
int sb2_lowdelay_decode (const sb2_lowdelay_params *params, const uint8_t *slices, size_t picture_bytes,
    size_t picture_pitch, const sb2_slab *coeffs, int is_s32, void *stream);

#ifdef __cplusplus
}
#endif
#endif
