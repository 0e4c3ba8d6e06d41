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

#ifdef __cplusplus
}
#endif
#endif
