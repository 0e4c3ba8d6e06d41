"""CPU: the structs in include/schro_b200_compat.h have the reference's layout
(sizeof + offsetof of every field), checked by compiling a probe against both header sets.
Needs the reference headers, so it only runs where /root/reference exists."""
import os
import subprocess
import tempfile

import pytest

from tests import helpers

REF = "/root/reference"
FIELDS = {
    "SchroFrameData": "format data stride width height length h_shift v_shift",
    "SchroFrame": "refcount free domain regions priv format width height components is_virtual "
                  "cached_lines virt_frame1 virt_frame2 render_line virt_priv virt_priv2 extension "
                  "cache_offset is_upsampled upsample_done",
    "SchroMemoryDomain": "mutex flags alloc alloc_2d free slots",
    "SchroVideoFormat": "index width height chroma_format interlaced clean_width luma_offset "
                        "colour_primaries interlaced_coding unused2",
    "SchroParams": "video_format is_noarith wavelet_filter_index transform_depth horiz_codeblocks "
                   "vert_codeblocks codeblock_mode_index num_refs have_global_motion xblen_luma yblen_luma "
                   "xbsep_luma ybsep_luma mv_precision global_motion picture_pred_mode picture_weight_bits "
                   "picture_weight_1 picture_weight_2 is_lowdelay quant_matrix iwt_chroma_width "
                   "iwt_chroma_height iwt_luma_width iwt_luma_height x_num_blocks y_num_blocks x_offset y_offset",
    "SchroMotionVector": "metric chroma_metric u",
    "SchroMotionField": "x_num_blocks y_num_blocks motion_vectors",
    "SchroMotion": "src1 src2 motion_vectors params ref_weight_precision ref1_weight ref2_weight "
                   "mv_precision xoffset yoffset xbsep ybsep xblen yblen block alloc_block obmc_weight "
                   "alloc_block_ref block_ref weight_x weight_y width height max_fast_x max_fast_y "
                   "simple_weight oneref_noscale",
    "SchroHierBm": "ref_count ref hierarchy_levels params downsampled_src downsampled_ref "
                   "downsampled_mf use_chroma",
}


def probe_source(includes):
    lines = ["#include <stdio.h>", "#include <stddef.h>"] + includes + ["int main(void){"]
    for st, fields in FIELDS.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fields.split():
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ["return 0;}"]
    return "\n".join(lines)


def run_probe(src, cflags):
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-std=gnu99", "-w"] + cflags + [c, "-o", exe])
        return subprocess.check_output([exe], text=True)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers not present on this box")
def test_compat_structs_match_reference_layout():
    gen = os.path.join(helpers.ROOT, "oracle", "_ref", "gen")
    if not os.path.isdir(gen):
        subprocess.check_call(["bash", os.path.join(helpers.ROOT, "oracle", "build_ref.sh")])
    ref = run_probe(probe_source(["#include <schroedinger/schro.h>",
                                  "#include <schroedinger/schromotionest.h>"]),
                    ["-DSCHRO_ENABLE_UNSTABLE_API", f"-I{helpers.ROOT}/oracle/refshim", f"-I{gen}",
                     f"-I{REF}"])
    ours = run_probe(probe_source(['#include "schro_b200_compat.h"']),
                     [f"-I{helpers.ROOT}/include"])
    assert ref == ours
