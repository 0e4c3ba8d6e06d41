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
    ref_inited = 1;
  }
}

int oracle_ref_harness_version (void) { return 2; }

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
