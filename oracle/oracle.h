/*
 * oracle.h -- CPU restatement of the schroedinger picture core.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load liboracle.so.  The product
 * (schroedinger_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against
 * the unmodified reference compiled into oracle/_ref/libschro_ref.so
 * (oracle/build_ref.sh) by tests/test_oracle_vs_ref.py, and against the
 * committed golden vectors in tests/golden/ (generated from that same
 * reference build by tests/golden/make_golden.py).
 *
 * All functions use a flat C ABI (plain pointers, strides in BYTES).
 * The identically shaped ref_* functions in oracle/ref_harness.c drive the
 * real reference; the sb2_* functions in include/schro_b200.h are the product.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Filter ids: reference schroedinger/schrobitstream.h:124-132 */
enum {
  ORACLE_WAVELET_DESLAURIERS_DUBUC_9_7 = 0,
  ORACLE_WAVELET_LE_GALL_5_3 = 1,
  ORACLE_WAVELET_DESLAURIERS_DUBUC_13_7 = 2,
  ORACLE_WAVELET_HAAR_0 = 3,
  ORACLE_WAVELET_HAAR_1 = 4,
  ORACLE_WAVELET_FIDELITY = 5,
  ORACLE_WAVELET_DAUBECHIES_9_7 = 6
};

/* ---- wavelets (oracle_wavelet.c) ---- */
/* one level, in place; follows schro_wavelet_transform_2d
 * (schroedinger/schrowaveletorc.c:60-117) */
void oracle_wavelet_fwd (void *data, int stride, int width, int height,
    int is_s32, int filter);
/* one level, in place (dest == src); follows schro_wavelet_inverse_transform_2d
 * (schroedinger/schrowaveletorc.c:121-188) */
void oracle_wavelet_inv (void *data, int stride, int width, int height,
    int is_s32, int filter);
/* multi-level drivers on one component plane; follow
 * schro_frame_iwt_transform (schroedinger/schroframe.c:1192-1228) and
 * schro_decoder_inverse_iwt_transform (schroedinger/schrodecoder.c:1809-1853) */
void oracle_iwt_fwd (void *data, int stride, int width, int height,
    int is_s32, int filter, int depth);
void oracle_iwt_inv (void *data, int stride, int width, int height,
    int is_s32, int filter, int depth);

/* ---- upsampled reference frames (oracle_frame.c) ---- */
/* `data` points at pixel (0,0) of phase 0 of ONE component of an upsampled
 * frame: 4 phase planes side by side in each row, phase p at
 * data + (stride>>2)*p, every phase surrounded by `ext` border pixels
 * (schroedinger/schroframe.c:60-191, 1917-1925). */
/* schro_frame_mc_edgeextend on one plane (schroedinger/schroframe.c:1940-1997) */
void oracle_mc_edgeextend (uint8_t *data, int stride, int width, int height,
    int ext);
/* schro_upsampled_frame_upsample on one component, phase 0 already
 * edge-extended (schroedinger/schroframe.c:2000-2030) */
void oracle_upsample (uint8_t *data, int stride, int width, int height,
    int ext);
/* schro_frame_downsample on one component (schroedinger/schroframe.c:1400-1513) */
void oracle_downsample (uint8_t *dest, int dstride, int dwidth, int dheight,
    const uint8_t *src, int sstride, int swidth, int sheight);

/* ---- OBMC (oracle_motion.c) ---- */
/* Same memory layout as SchroMotionVector (schroedinger/schromotion.h:20-37) */
typedef struct {
  uint32_t flags;               /* pred_mode:2 using_global:1 split:2 unused:3 scan:8 */
  uint32_t metric;
  uint32_t chroma_metric;
  int16_t v[4];                 /* vec: dx[0],dx[1],dy[0],dy[1]   dc: dc[0],dc[1],dc[2] */
} OracleMotionVector;

typedef struct {
  int xbsep, ybsep, xblen, yblen;       /* of THIS component */
  int x_num_blocks, y_num_blocks;
  int mv_precision;
  int weight1, weight2, weight_bits;    /* picture_weight_1/2/bits */
  int h_shift, v_shift;                 /* chroma shifts applied to vectors (0 for luma) */
  int comp;                             /* component index (selects dc[k]) */
} OracleObmcParams;

/* One component of schro_motion_render_u8 (schroedinger/schromotion8.c:700-929).
 * ref0/ref1: phase-0 pixel (0,0) of the upsampled references (ref1 may be NULL),
 * rstride their 4-phase row stride.  acc (int16, may be NULL) receives what the
 * reference leaves in `dest`.  add!=0: out = clamp_u8(residual + ((acc+32)>>6)),
 * residual is s16 (res_is_s32==0) or s32.  add==0: dest=acc:=(acc-8160)>>6 and
 * residual(s16) -= that. */
void oracle_obmc_render (const OracleObmcParams *p, const OracleMotionVector *mvs,
    const uint8_t *ref0, const uint8_t *ref1, int rstride,
    int width, int height,
    int16_t *acc, int acc_stride,
    void *residual, int res_stride, int res_is_s32,
    int add, uint8_t *out, int out_stride);

/* One component of schro_motion_render_ref (schroedinger/schromotionref.c:245-330), the per-pixel renderer
 * the reference switches to when global motion is on (schroedinger/schromotion.c:113-121).  global_motion:
 * 2 x 10 ints (b0 b1 a_exp a00 a01 a10 a11 c_exp c0 c1 per reference) used by blocks flagged using_global.
 * width / height: the component's size; strides of acc / residual in SAMPLES.  acc (may be NULL) receives
 * clamp (prediction) - 128; add != 0: out = clamp (residual + that + 128), else residual -= that. */
void oracle_obmc_render_ref (const OracleObmcParams *p, const int *global_motion, const OracleMotionVector *mvs,
    const uint8_t *ref0, const uint8_t *ref1, int rstride, int width, int height, int16_t *acc, int acc_stride,
    int16_t *residual, int res_stride, int add, uint8_t *out, int out_stride);

/* ---- SAD / hierarchical block matching (oracle_hbm.c) ---- */
/* schro_metric_absdiff_u8 (schroedinger/schrometric.c:10-29) */
uint32_t oracle_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b,
    int b_stride, int width, int height);

typedef struct {
  /* pixel (0,0) pointers of the three u8 components, edge-extended by `ext` */
  const uint8_t *data[3];
  int stride[3];
  int width, height;            /* luma size */
  int h_shift, v_shift;
  int ext;
} OraclePyrLevel;

/* One call of schro_hierarchical_bm_scan_hint (schroedinger/schrohierbm.c:174-383).
 * mf: output field (x_num_blocks*y_num_blocks), parent: field of level shift+1
 * or NULL. */
void oracle_hbm_scan_hint (const OraclePyrLevel *src, const OraclePyrLevel *ref,
    int xbsep, int ybsep, int x_num_blocks, int y_num_blocks, int ref_index,
    int shift, int h_range, int use_chroma,
    const OracleMotionVector *parent, OracleMotionVector *mf);

/* The metric-scan entry points on their own (schroedinger/schrometric.c:31-214, 380-414):
 * window set-up, the grid of SADs (metrics[i * scan_h + j], 42 x 42 arrays as SchroMetricScan has),
 * the arg-min with the reference's tie-break, and the 3-component block SAD of the candidate ranking. */
void oracle_metric_scan_setup (const OraclePyrLevel *f, int x, int y, int bw, int bh, int dx, int dy,
    int dist, int *ref_x, int *ref_y, int *scan_w, int *scan_h);
void oracle_metric_scan_do_scan (const OraclePyrLevel *src, const OraclePyrLevel *ref, int x, int y,
    int bw, int bh, int ref_x, int ref_y, int scan_w, int scan_h, int use_chroma, uint32_t *metrics,
    uint32_t *chroma_metrics);
uint32_t oracle_metric_scan_get_min (const uint32_t *metrics, const uint32_t *chroma_metrics, int x, int y,
    int ref_x, int ref_y, int scan_w, int scan_h, int gravity_x, int gravity_y, int use_chroma, int *dx,
    int *dy, uint32_t *chroma_error);
int oracle_metric_fast_block (const OraclePyrLevel *src, const OraclePyrLevel *ref, int bw, int bh, int x,
    int y, int dx, int dy);

/* Rough (bigblock) motion search (oracle_rough.c): one level of
 * schro_rough_me_heirarchical_scan_nohint / _hint (schroedinger/schroroughmotion.c:62-300).
 * mf: output field (x_num_blocks*y_num_blocks), parent: field of level shift+1. */
void oracle_rough_scan_nohint (const OraclePyrLevel *src, const OraclePyrLevel *ref, int xbsep, int ybsep,
    int x_num_blocks, int y_num_blocks, int ref_index, int shift, int distance, OracleMotionVector *mf);
void oracle_rough_scan_hint (const OraclePyrLevel *src, const OraclePyrLevel *ref, int xbsep, int ybsep,
    int x_num_blocks, int y_num_blocks, int ref_index, int shift, int distance,
    const OracleMotionVector *parent, OracleMotionVector *mf);

/* schro_pack_estimate_sint (schroedinger/schropack.c:204-226) and one pixel of a block fetched at sub-pel
 * precision 0..3 (schroedinger/schroframe.c:2458-2482), shared by oracle_subpel.c and oracle_split2.c */
int oracle_bits_sint (int value);
int oracle_subpel_sample (const uint8_t *ref, int rstride, int prec, int x, int y, int a, int b);

/* Split-2 pass of the mode decision (oracle_split2.c): schro_do_split2 + schro_motion_copy_to
 * (schroedinger/schromotionest.c:1601-1802, 1511-1523) for every superblock.  src: pixel (0,0) of the three
 * source planes; ref0 / ref1: phase-0 pixel (0,0) of the three planes of each upsampled reference (ref1
 * unused with one reference), rstride their 4-phase row strides; field0 / field1: the references' sub-pel
 * fields.  motion: x_num_blocks * y_num_blocks decided blocks; sb_error / sb_entropy: per superblock. */
typedef struct {
  int width, height;            /* luma */
  int h_shift, v_shift;
  int orig_ext;                 /* extension of the source frame (enters the bi-reference range test) */
  int xblen, yblen;             /* = xbsep_luma, ybsep_luma */
  int x_num_blocks, y_num_blocks;
  int mv_precision, num_refs;
  double lambda;
} OracleSplit2Params;
void oracle_split2_decide (const OracleSplit2Params *p, const uint8_t *const src[3], const int src_stride[3],
    const uint8_t *const ref0[3], const uint8_t *const ref1[3], const int rstride[3],
    const OracleMotionVector *field0, const OracleMotionVector *field1, OracleMotionVector *motion,
    int *sb_error, int *sb_entropy);

/* Sub-pel refinement of one reference's motion field in place (oracle_subpel.c):
 * schro_encoder_motion_predict_subpel_deep (schroedinger/schromotionest.c:246-355) for one reference.
 * orig: luma pixel (0,0) of the source picture (frame extension orig_ext -- it only enters the range
 * test); upref: phase-0 luma pixel (0,0) of the upsampled reference, rstride its 4-phase row stride. */
void oracle_subpel_refine (const uint8_t *orig, int orig_stride, int width, int height, int orig_ext,
    const uint8_t *upref, int rstride, int xblen, int yblen, int x_num_blocks, int y_num_blocks,
    int mv_precision, int ref_index, double lambda, OracleMotionVector *mf);

/* The low-delay slice decoder (oracle_lowdelay.c): schro_decoder_decode_lowdelay_transform_data
 * (schroedinger/schrolowdelay.c:99-761) + DC prediction (schroedinger/schrodecoder.c:3219-3277) for one picture.
 * data: the picture's slices back to back; planes / strides (in SAMPLES) / widths / heights: the three
 * coefficient planes (in-place subband layout); quant_matrix[1 + 3 * depth]; the reference's quantiser
 * tables (61 entries).  orc16 != 0: the s16 "fast" path's 16-bit dequantiser (and its length-field width). */
void oracle_lowdelay_decode (const uint8_t *data, int data_bytes, int slice_bytes_num, int slice_bytes_denom, int n_horiz_slices,
    int n_vert_slices, int transform_depth, const int *quant_matrix, const uint32_t *table_quant,
    const uint32_t *table_offset, void **planes, const int *strides, const int *widths, const int *heights,
    int is_s32, int orc16);

#ifdef __cplusplus
}
#endif

/* ---- combine / convert glue --------------------------------------
 * depth: 0 u8, 1 s16, 2 s32.  One plane at a time.
 * oracle_convert_plane: schro_frame_convert for planar frames of equal chroma format
 *   (schroedinger/schroframe.c:870-978): depth conversion with Orc's semantics
 *   (schroedinger/schroorc.orc:476-549), then crop or edge extension
 *   (schroedinger/schrovirtframe.c:1824-1960): dest(x,y) = conv(src(min(x,sw-1), min(y,sh-1))).
 * oracle_add_plane: schro_frame_add / schro_frame_subtract (schroedinger/schroframe.c:1012-1182):
 *   dest (s16) +-= src (s16 or u8) over the common area, 16-bit wrap. */
void oracle_convert_plane (void *dst, int dstride, int ddepth, int dwidth, int dheight,
    const void *src, int sstride, int sdepth, int swidth, int sheight);
void oracle_add_plane (int16_t *dst, int dstride, int dwidth, int dheight,
    const void *src, int sstride, int sdepth, int swidth, int sheight, int subtract);


/* ---- dequantisation of a coefficient frame in place ----------------
 * What the decoder does per codeblock (schroedinger/schrodecoder.c:3395-3448) with
 * orc_dequantise_s16_ip_2d / _s32_ip_2d (schroedinger/schroorc.orc:1154-1168, 2148-2162), for one
 * component plane in the in-place subband layout (schro_subband_get_frame_data,
 * schroedinger/schroparams.c:319-352).  quant: for band index 0..3*depth (schro_subband_get_position),
 * then codeblock row, then codeblock column: (quant_factor, quant_offset + 2) pairs.
 * hcb / vcb: codeblocks per band, [0] for the LL band, [i + 1] for the bands of level i. */
void oracle_dequantise_plane (void *data, int stride, int width, int height, int is_s32,
    int transform_depth, const int *hcb, const int *vcb, const int32_t *quant);
#endif
