/* compat/schro_lowdelay.c -- schro_decoder_decode_lowdelay_transform_data, reference side.
 *
 * The reference's function (schroedinger/schrolowdelay.c:745-761) takes the decoder's SchroPicture and
 * reads three things from it.  This file keeps the symbol and its signature and hands those three things to
 * libschro_b200 (schro_b200_decode_lowdelay_transform_data, include/schro_b200_compat.h), which uploads the
 * compressed slices and decodes them on the GPU.  Compiled AGAINST THE REFERENCE'S OWN HEADERS by
 * oracle/build_ref.sh; it replaces the body in schrolowdelay.c, so schro_decoder_x_decode_residual
 * (schroedinger/schrodecoder.c:1795-1807) is unchanged. */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif
#include <schroedinger/schro.h>
#include <schroedinger/schrodecoder.h>

void schro_b200_decode_lowdelay_transform_data (SchroParams * params, const uint8_t * data, int length,
    SchroFrame * transform_frame);

void
schro_decoder_decode_lowdelay_transform_data (SchroPicture * picture)
{
  schro_b200_decode_lowdelay_transform_data (&picture->params, picture->lowdelay_buffer->data,
      (int) picture->lowdelay_buffer->length, picture->transform_frame);
}
