"""Device-side picture slabs and thin wrappers over the sb2_* C-ABI.

torch is used only for device memory, streams and host<->device copies
(plumbing); all arithmetic happens in libschro_b200.so.

Layout follows the reference's frame allocator
(schroedinger/schroframe.c:60-191): planar components in one region, each
plane surrounded by ``extension`` border pixels, row stride
``ROUND_UP_16((w + 2*ext) * bpp)`` and x4 for "upsampled" frames whose four
half-pel phase planes sit side by side in every row.  A slab repeats that
frame layout ``count`` times at a fixed pitch so a batch is one launch.
"""
import ctypes

import numpy as np
import torch

from ._lib import Slab, lib, check

_BPP = {"u8": 1, "s16": 2, "s32": 4}
_NP = {"u8": np.uint8, "s16": np.int16, "s32": np.int32}
_TORCH = {"u8": torch.uint8, "s16": torch.int16, "s32": torch.int32}


def _round_up(x, a):
    return (x + a - 1) // a * a


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("schroedinger_b200 needs a CUDA device: there is no CPU fallback")


class FrameLayout:
    """Byte layout of one frame (all components), reference-compatible."""

    def __init__(self, depth, comp_sizes, extension=0, upsampled=False):
        self.depth = depth
        self.bpp = _BPP[depth]
        self.comp_sizes = [(int(w), int(h)) for (w, h) in comp_sizes]
        self.extension = int(extension)
        self.upsampled = bool(upsampled)
        self.stride, self.length, self.offset = [], [], []
        pos = 0
        for (w, h) in self.comp_sizes:
            stride = _round_up((w + 2 * self.extension) * self.bpp, 16)
            if self.upsampled:
                stride *= 4
            length = stride * (h + 2 * self.extension)
            self.stride.append(stride)
            self.length.append(length)
            # pixel (0,0) of (phase 0 of) the plane
            self.offset.append(pos + stride * self.extension + self.bpp * self.extension)
            pos += length
        self.frame_bytes = pos
        self.pitch = _round_up(pos, 256)

    @staticmethod
    def yuv420(depth, width, height, extension=0, upsampled=False):
        cw, ch = (width + 1) // 2, (height + 1) // 2
        return FrameLayout(depth, [(width, height), (cw, ch), (cw, ch)], extension, upsampled)


class PictureSlab:
    """``count`` frames of one layout in a single device allocation."""

    def __init__(self, layout, count, device="cuda", zero=True):
        require_cuda()
        self.layout = layout
        self.count = int(count)
        alloc = torch.zeros if zero else torch.empty
        # 256 spare bytes: the byte-SIMD SAD path reads whole aligned words
        self.buf = alloc(layout.pitch * self.count + 256, dtype=torch.uint8, device=device)
        self.slab = self._make_slab()

    def _make_slab(self):
        L = self.layout
        s = Slab()
        s.base = self.buf.data_ptr()
        s.picture_pitch = L.pitch
        s.count = self.count
        s.ncomp = len(L.comp_sizes)
        for c, (w, h) in enumerate(L.comp_sizes):
            s.offset[c] = L.offset[c]
            s.stride[c] = L.stride[c]
            s.width[c] = w
            s.height[c] = h
        return s

    @property
    def nbytes(self):
        return self.layout.pitch * self.count

    def plane(self, pic, comp, phase=0, with_border=False):
        """Strided torch view of one plane (of one phase) of one picture."""
        L = self.layout
        w, h = L.comp_sizes[comp]
        ext = L.extension
        start = pic * L.pitch + L.offset[comp]
        if L.upsampled:
            start += (L.stride[comp] >> 2) * phase
        if with_border:
            start -= L.stride[comp] * ext + L.bpp * ext
            w, h = w + 2 * ext, h + 2 * ext
        return _as_typed(self.buf, start, h, w, L.stride[comp], L.depth)

    def upload(self, pic, comp, array, phase=0, with_border=False):
        dst = self.plane(pic, comp, phase, with_border)
        src = torch.from_numpy(np.ascontiguousarray(array, dtype=_NP[self.layout.depth]))
        dst.copy_(src.to(dst.device, non_blocking=False))

    def download(self, pic, comp, phase=0, with_border=False):
        return self.plane(pic, comp, phase, with_border).cpu().numpy().copy()


class SlabView:
    """The same device memory as another slab seen with different component sizes (e.g. the
    picture-size window of an iwt-size coefficient frame)."""

    def __init__(self, parent, comp_sizes):
        import copy
        self.layout = copy.copy(parent.layout)
        self.layout.comp_sizes = [(int(w), int(h)) for (w, h) in comp_sizes]
        self.count = parent.count
        self.buf = parent.buf
        self.slab = PictureSlab._make_slab(self)

    plane = PictureSlab.plane
    upload = PictureSlab.upload
    download = PictureSlab.download


def _as_typed(buf, start, h, w, stride, depth):
    bpp = _BPP[depth]
    assert start % bpp == 0 and stride % bpp == 0
    typed = buf.view(_TORCH[depth])
    # as_strided's offset is absolute in the storage: honour a buffer that is itself a view
    return torch.as_strided(typed, (h, w), (stride // bpp, 1), typed.storage_offset() + start // bpp)


def _stream_ptr(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(stream.cuda_stream)


class Workspace:
    """Grow-only device scratch buffer."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        if nbytes == 0:
            return ctypes.c_void_p(0), 0
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        return ctypes.c_void_p(self.buf.data_ptr()), self.buf.numel()


_default_ws = Workspace()


def iwt_workspace_bytes(slab, depth_bits, transform_depth, in_place):
    return lib.sb2_iwt_workspace_bytes(ctypes.byref(slab), 1 if depth_bits == "s32" else 0,
                                       transform_depth, 1 if in_place else 0)


def iwt_forward(src, dst, filter_index, transform_depth, workspace=None, stream=None):
    """Multi-level forward transform of every component of every picture of ``src``
    into ``dst`` (PictureSlab; may be the same slab)."""
    _iwt(lib.sb2_iwt_forward, "sb2_iwt_forward", src, dst, filter_index, transform_depth,
         workspace, stream)


def iwt_inverse(src, dst, filter_index, transform_depth, workspace=None, stream=None):
    _iwt(lib.sb2_iwt_inverse, "sb2_iwt_inverse", src, dst, filter_index, transform_depth,
         workspace, stream)


def _iwt(fn, name, src, dst, filter_index, transform_depth, workspace, stream):
    require_cuda()
    is_s32 = 1 if src.layout.depth == "s32" else 0
    assert src.layout.depth == dst.layout.depth and src.layout.depth in ("s16", "s32")
    in_place = src.buf.data_ptr() == dst.buf.data_ptr()
    ws = workspace or _default_ws
    need = lib.sb2_iwt_workspace_bytes(ctypes.byref(src.slab), is_s32, transform_depth,
                                       1 if in_place else 0)
    ptr, size = ws.get(need)
    check(fn(ctypes.byref(src.slab), ctypes.byref(dst.slab), is_s32, filter_index,
             transform_depth, ptr, size, _stream_ptr(stream)), name)


# ---------------------------------------------------------------------------------------
# reference-frame preparation
# ---------------------------------------------------------------------------------------
def mc_edgeextend(frames, phase=0, stream=None):
    """schro_frame_mc_edgeextend on every plane of every picture of an (extended) u8 slab."""
    require_cuda()
    check(lib.sb2_mc_edgeextend(ctypes.byref(frames.slab), frames.layout.extension, phase,
                                _stream_ptr(stream)), "sb2_mc_edgeextend")


def upsample(frames, stream=None):
    """schro_upsampled_frame_upsample: phases 1..3 + borders from edge-extended phase 0."""
    require_cuda()
    assert frames.layout.upsampled and frames.layout.depth == "u8"
    check(lib.sb2_upsample(ctypes.byref(frames.slab), frames.layout.extension, _stream_ptr(stream)),
          "sb2_upsample")


def edgeextend_upsample(frames, stream=None):
    """schro_frame_mc_edgeextend + schro_upsampled_frame_upsample fused into one launch."""
    require_cuda()
    assert frames.layout.upsampled and frames.layout.depth == "u8"
    check(lib.sb2_edgeextend_upsample(ctypes.byref(frames.slab), frames.layout.extension,
                                      _stream_ptr(stream)), "sb2_edgeextend_upsample")


def downsample_edgeextend(src, dst, stream=None):
    """schro_frame_downsample + schro_frame_mc_edgeextend(dst) fused into one launch."""
    require_cuda()
    check(lib.sb2_downsample_edgeextend(ctypes.byref(src.slab), ctypes.byref(dst.slab),
                                        dst.layout.extension, _stream_ptr(stream)),
          "sb2_downsample_edgeextend")


_DEPTH_CODE = {"u8": 0, "s16": 1, "s32": 2}


def frame_convert(src, dst, stream=None):
    """schro_frame_convert for planar slabs of equal chroma format: depth conversion (+-128 offset,
    Orc's wrap / saturation points) with crop or edge extension to dst's size."""
    require_cuda()
    check(lib.sb2_frame_convert(ctypes.byref(src.slab), _DEPTH_CODE[src.layout.depth], ctypes.byref(dst.slab),
                                _DEPTH_CODE[dst.layout.depth], _stream_ptr(stream)), "sb2_frame_convert")


def frame_add(dst, src, subtract=False, stream=None):
    """schro_frame_add / schro_frame_subtract: dst (s16) +-= src (u8 or s16) over the common area."""
    require_cuda()
    check(lib.sb2_frame_add(ctypes.byref(dst.slab), ctypes.byref(src.slab), _DEPTH_CODE[src.layout.depth],
                            int(bool(subtract)), _stream_ptr(stream)), "sb2_frame_add")


def dequantise(coeffs, depth, horiz_codeblocks, vert_codeblocks, quant, stream=None):
    """Dequantise a coefficient slab in place.  quant: int32 CUDA tensor, per picture
    sb2_dequant_table_pairs() (factor, offset + 2) pairs (see include/schro_b200.h)."""
    from ._lib import DequantParams
    require_cuda()
    p = DequantParams()
    p.transform_depth = depth
    for i in range(7):
        p.horiz_codeblocks[i] = horiz_codeblocks[i] if i < len(horiz_codeblocks) else 1
        p.vert_codeblocks[i] = vert_codeblocks[i] if i < len(vert_codeblocks) else 1
    pairs = lib.sb2_dequant_table_pairs(ctypes.byref(p), len(coeffs.layout.comp_sizes))
    assert quant.dtype == torch.int32 and quant.numel() == 2 * pairs * coeffs.count, (quant.numel(), pairs)
    check(lib.sb2_dequantise(ctypes.byref(coeffs.slab), 1 if coeffs.layout.depth == "s32" else 0, ctypes.byref(p),
                             ctypes.c_void_p(quant.data_ptr()), ctypes.c_size_t(pairs), _stream_ptr(stream)),
          "sb2_dequantise")


def dequantise_widen(quantised, coeffs, depth, horiz_codeblocks, vert_codeblocks, quant, stream=None):
    """s16 quantised coefficients -> dequantised s32 coefficients in a second slab (sb2_dequantise_widen)."""
    from ._lib import DequantParams
    require_cuda()
    assert quantised.layout.depth == "s16" and coeffs.layout.depth == "s32"
    p = DequantParams()
    p.transform_depth = depth
    for i in range(7):
        p.horiz_codeblocks[i] = horiz_codeblocks[i] if i < len(horiz_codeblocks) else 1
        p.vert_codeblocks[i] = vert_codeblocks[i] if i < len(vert_codeblocks) else 1
    pairs = lib.sb2_dequant_table_pairs(ctypes.byref(p), len(coeffs.layout.comp_sizes))
    assert quant.dtype == torch.int32 and quant.numel() == 2 * pairs * coeffs.count, (quant.numel(), pairs)
    check(lib.sb2_dequantise_widen(ctypes.byref(quantised.slab), ctypes.byref(coeffs.slab), ctypes.byref(p),
                                   ctypes.c_void_p(quant.data_ptr()), ctypes.c_size_t(pairs), _stream_ptr(stream)),
          "sb2_dequantise_widen")


def downsample(src, dst, stream=None):
    """schro_frame_downsample: dst = half-size src (per component)."""
    require_cuda()
    check(lib.sb2_downsample(ctypes.byref(src.slab), ctypes.byref(dst.slab), _stream_ptr(stream)),
          "sb2_downsample")


# ---------------------------------------------------------------------------------------
# OBMC
# ---------------------------------------------------------------------------------------
class ObmcParams(ctypes.Structure):
    """Mirror of sb2_obmc_params."""
    _fields_ = [(n, ctypes.c_int) for n in (
        "xbsep", "ybsep", "xblen", "yblen", "x_num_blocks", "y_num_blocks", "mv_precision",
        "picture_weight_1", "picture_weight_2", "picture_weight_bits", "chroma_h_shift",
        "chroma_v_shift")]


def obmc_render(params, mvs, ref0, ref1, residual, add, out=None, acc=None, stream=None):
    """schro_motion_render for every picture.  mvs: uint8 CUDA tensor holding
    count x (x_num_blocks*y_num_blocks) SchroMotionVector structs (20 bytes each)."""
    require_cuda()
    n = params.x_num_blocks * params.y_num_blocks
    assert mvs.is_cuda and mvs.dtype == torch.uint8 and mvs.numel() >= residual.count * n * 20
    null = ctypes.POINTER(Slab)()
    check(lib.sb2_obmc_render(
        ctypes.byref(params), ctypes.c_void_p(mvs.data_ptr()), ctypes.c_size_t(n),
        ctypes.byref(ref0.slab), ctypes.byref(ref1.slab) if ref1 is not None else null,
        ctypes.byref(acc.slab) if acc is not None else null, ctypes.byref(residual.slab),
        1 if residual.layout.depth == "s32" else 0, 1 if add else 0,
        ctypes.byref(out.slab) if out is not None else null, _stream_ptr(stream)), "sb2_obmc_render")


# ---------------------------------------------------------------------------------------
# hierarchical block matching
# ---------------------------------------------------------------------------------------
class HbmParams(ctypes.Structure):
    """Mirror of sb2_hbm_params."""
    _fields_ = [(n, ctypes.c_int) for n in (
        "xbsep", "ybsep", "x_num_blocks", "y_num_blocks", "ref_index", "use_chroma",
        "chroma_h_shift", "chroma_v_shift")]


def hbm_scan_hint(params, src_level, ref_level, shift, h_range, parent, out, workspace=None,
                  stream=None):
    """One schro_hierarchical_bm_scan_hint level for every (picture, reference) pair.
    parent / out: uint8 CUDA tensors of count x nblocks x 20 bytes (parent may be None)."""
    require_cuda()
    n = params.x_num_blocks * params.y_num_blocks
    ws = workspace or _default_ws
    need = lib.sb2_hbm_workspace_bytes(params.x_num_blocks, params.y_num_blocks, src_level.count)
    ptr, size = ws.get(need)
    check(lib.sb2_hbm_scan_hint(
        ctypes.byref(params), ctypes.byref(src_level.slab), ctypes.byref(ref_level.slab),
        src_level.layout.extension, shift, h_range,
        ctypes.c_void_p(parent.data_ptr()) if parent is not None else None,
        ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n), ptr, size, _stream_ptr(stream)),
        "sb2_hbm_scan_hint")


def hbm_level_ranges(levels, level0_range=3):
    """(level, half_range) pairs in the order schro_hbm_scan (schrohierbm.c:158-172) and the
    level-0 refinement (schromotionest.c:123-127) run them."""
    order = [(levels, 20)]
    r = 10
    for l in range(levels - 1, 0, -1):
        order.append((l, max(3, r)))
        r >>= 1
    if level0_range > 0:
        order.append((0, level0_range))
    return order


class Pyramid:
    """Downsampled copies of `count` u8 4:2:0 pictures: level 0 = the pictures (ext 32),
    level i+1 = sb2_downsample(level i) with extension max(xbsep, ybsep), edge-extended
    (schro_encoder_frame_downsample, schroedinger/schroanalysis.c:9-28)."""

    def __init__(self, width, height, count, levels, ext=8, level0=None):
        self.levels = levels
        self.slabs = [level0 if level0 is not None else
                      PictureSlab(FrameLayout.yuv420("u8", width, height, 32), count)]
        w, h = width, height
        cw, ch = (width + 1) // 2, (height + 1) // 2
        for _ in range(levels):
            w, h, cw, ch = (w + 1) // 2, (h + 1) // 2, (cw + 1) // 2, (ch + 1) // 2
            self.slabs.append(PictureSlab(FrameLayout("u8", [(w, h), (cw, ch), (cw, ch)], ext), count))

    def build(self, stream=None, fused=True):
        mc_edgeextend(self.slabs[0], stream=stream)
        for l in range(self.levels):
            if fused:
                downsample_edgeextend(self.slabs[l], self.slabs[l + 1], stream=stream)
            else:
                downsample(self.slabs[l], self.slabs[l + 1], stream=stream)
                mc_edgeextend(self.slabs[l + 1], stream=stream)


def hbm_scan(params, src_pyr, ref_pyr, level0_range=3, fields=None, workspace=None, stream=None):
    """schro_hbm_scan (+ level-0 refinement): returns a list fields[level] of uint8 CUDA tensors."""
    levels = src_pyr.levels
    count = src_pyr.slabs[0].count
    n = params.x_num_blocks * params.y_num_blocks
    if fields is None:
        fields = [torch.empty(count * n * 20, dtype=torch.uint8, device="cuda") for _ in range(levels + 1)]
    for (l, r) in hbm_level_ranges(levels, level0_range):
        hbm_scan_hint(params, src_pyr.slabs[l], ref_pyr.slabs[l], l, r,
                      fields[l + 1] if l < levels else None, fields[l], workspace, stream)
    return fields


def rough_scan_nohint(params, src_level, ref_level, shift, distance, out, stream=None):
    """schro_rough_me_heirarchical_scan_nohint for every (picture, reference) pair; out: uint8 CUDA
    tensor of count x nblocks x 20 bytes."""
    require_cuda()
    n = params.x_num_blocks * params.y_num_blocks
    check(lib.sb2_rough_scan_nohint(
        ctypes.byref(params), ctypes.byref(src_level.slab), ctypes.byref(ref_level.slab),
        src_level.layout.extension, shift, distance, ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n),
        _stream_ptr(stream)), "sb2_rough_scan_nohint")


def rough_scan_hint(params, src_level, ref_level, shift, distance, parent, out, workspace=None, stream=None):
    """schro_rough_me_heirarchical_scan_hint; parent = the field of level shift + 1."""
    require_cuda()
    n = params.x_num_blocks * params.y_num_blocks
    ws = workspace or _default_ws
    ptr, size = ws.get(lib.sb2_rough_workspace_bytes(params.x_num_blocks, params.y_num_blocks, src_level.count))
    check(lib.sb2_rough_scan_hint(
        ctypes.byref(params), ctypes.byref(src_level.slab), ctypes.byref(ref_level.slab),
        src_level.layout.extension, shift, distance, ctypes.c_void_p(parent.data_ptr()),
        ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n), ptr, size, _stream_ptr(stream)),
        "sb2_rough_scan_hint")


def rough_scan(params, src_pyr, ref_pyr, nohint_distance=12, hint_distance=4, fields=None, workspace=None,
               stream=None):
    """schro_rough_me_heirarchical_scan (schroedinger/schroroughmotion.c:46-60): fields[level] for
    level = levels .. 1 (fields[0] is not produced by the rough search)."""
    levels = src_pyr.levels
    count = src_pyr.slabs[0].count
    n = params.x_num_blocks * params.y_num_blocks
    if fields is None:
        fields = [None] + [torch.empty(count * n * 20, dtype=torch.uint8, device="cuda") for _ in range(levels)]
    rough_scan_nohint(params, src_pyr.slabs[levels], ref_pyr.slabs[levels], levels, nohint_distance,
                      fields[levels], stream)
    for l in range(levels - 1, 0, -1):
        rough_scan_hint(params, src_pyr.slabs[l], ref_pyr.slabs[l], l, hint_distance, fields[l + 1], fields[l],
                        workspace, stream)
    return fields


def subpel_refine(orig, upref, field, xblen, yblen, x_num_blocks, y_num_blocks, mv_precision, ref_index, lam,
                  workspace=None, stream=None):
    """schro_encoder_motion_predict_subpel_deep for one reference: `field` (uint8 CUDA tensor,
    count x nblocks x 20 bytes) is refined in place."""
    from ._lib import SubpelParams
    require_cuda()
    n = x_num_blocks * y_num_blocks
    p = SubpelParams(xblen, yblen, x_num_blocks, y_num_blocks, mv_precision, ref_index, orig.layout.extension, lam)
    ws = workspace or _default_ws
    ptr, size = ws.get(lib.sb2_subpel_workspace_bytes(x_num_blocks, y_num_blocks, orig.count))
    check(lib.sb2_subpel_refine(ctypes.byref(p), ctypes.byref(orig.slab), ctypes.byref(upref.slab),
                                upref.layout.extension, ctypes.c_void_p(field.data_ptr()), ctypes.c_size_t(n),
                                ptr, size, _stream_ptr(stream)), "sb2_subpel_refine")


def split2_decide(orig, uprefs, fields, xblen, yblen, x_num_blocks, y_num_blocks, mv_precision, lam,
                  workspace=None, stream=None, out=None):
    """The split-2 pass of schro_mode_decision (schro_do_split2 for every superblock) for every picture:
    uprefs / fields are lists of one or two upsampled reference slabs / uint8 CUDA tensors (count x nblocks x 20
    bytes, vectors at mv_precision).  Returns (motion, sb_error, sb_entropy) as CUDA tensors."""
    import torch
    from ._lib import Split2Params
    require_cuda()
    n = x_num_blocks * y_num_blocks
    nsb = (x_num_blocks // 4) * (y_num_blocks // 4)
    count = orig.count
    p = Split2Params(xblen, yblen, x_num_blocks, y_num_blocks, mv_precision, len(uprefs), 1, 1, orig.layout.extension, lam)
    if out is not None:                 # caller-provided (motion uint8 [count*n*20], sb_error / sb_entropy int32 [count*nsb])
        motion, sb_error, sb_entropy = out
        assert motion.numel() == count * n * 20 and sb_error.numel() == count * nsb and sb_entropy.numel() == count * nsb
    else:
        motion = torch.empty(count * n * 20, dtype=torch.uint8, device="cuda")
        sb_error = torch.empty(count * nsb, dtype=torch.int32, device="cuda")
        sb_entropy = torch.empty(count * nsb, dtype=torch.int32, device="cuda")
    ws = workspace or _default_ws
    ptr, size = ws.get(lib.sb2_split2_workspace_bytes(x_num_blocks, y_num_blocks, count))
    two = len(uprefs) > 1
    check(lib.sb2_split2_decide(ctypes.byref(p), ctypes.byref(orig.slab), ctypes.byref(uprefs[0].slab),
                                ctypes.byref(uprefs[1].slab) if two else None, uprefs[0].layout.extension,
                                ctypes.c_void_p(fields[0].data_ptr()), ctypes.c_void_p(fields[1].data_ptr()) if two else None,
                                ctypes.c_size_t(n), ctypes.c_void_p(motion.data_ptr()), ctypes.c_size_t(n),
                                ctypes.c_void_p(sb_error.data_ptr()), ctypes.c_void_p(sb_entropy.data_ptr()),
                                ptr, size, _stream_ptr(stream)), "sb2_split2_decide")
    return motion, sb_error.view(count, nsb), sb_entropy.view(count, nsb)


def lowdelay_decode(slices, picture_bytes, coeffs, depth, n_horiz_slices, n_vert_slices, slice_bytes_num,
                    slice_bytes_denom, quant_matrix, table_quant, table_offset, picture_pitch=None, stream=None):
    """schro_decoder_decode_lowdelay_transform_data + DC prediction for every picture of `coeffs`;
    slices: uint8 CUDA tensor holding the pictures' slice buffers `picture_pitch` bytes apart."""
    from ._lib import LowdelayParams
    require_cuda()
    p = LowdelayParams()
    p.transform_depth, p.n_horiz_slices, p.n_vert_slices = depth, n_horiz_slices, n_vert_slices
    p.slice_bytes_num, p.slice_bytes_denom = slice_bytes_num, slice_bytes_denom
    for i, v in enumerate(quant_matrix):
        p.quant_matrix[i] = int(v)
    for i in range(61):
        p.table_quant[i] = int(table_quant[i])
        p.table_offset[i] = int(table_offset[i])
    check(lib.sb2_lowdelay_decode(ctypes.byref(p), ctypes.c_void_p(slices.data_ptr()), ctypes.c_size_t(picture_bytes),
                                  ctypes.c_size_t(picture_pitch if picture_pitch is not None else picture_bytes),
                                  ctypes.byref(coeffs.slab), 1 if coeffs.layout.depth == "s32" else 0, _stream_ptr(stream)),
          "sb2_lowdelay_decode")


def iwt_inverse_convert(src, dst_u8, filter_index, transform_depth, shift=0, workspace=None, stream=None):
    """Inverse transform with shift + convert to 8 bits fused into its last level (sb2_iwt_inverse_convert)."""
    require_cuda()
    is_s32 = 1 if src.layout.depth == "s32" else 0
    ws = workspace or _default_ws
    ptr, size = ws.get(lib.sb2_iwt_workspace_bytes(ctypes.byref(src.slab), is_s32, transform_depth, 0))
    check(lib.sb2_iwt_inverse_convert(ctypes.byref(src.slab), ctypes.byref(dst_u8.slab), is_s32, filter_index,
                                      transform_depth, shift, ptr, size, _stream_ptr(stream)), "sb2_iwt_inverse_convert")


def obmc_render_ref(params, global_motion, mvs, ref0, ref1, residual, add, out=None, acc=None, stream=None):
    """schro_motion_render_ref (what schro_motion_render runs when the picture has global motion);
    global_motion: 20 ints or None."""
    require_cuda()
    n = params.x_num_blocks * params.y_num_blocks
    null = ctypes.POINTER(Slab)()
    gm = (ctypes.c_int * 20)(*[int(v) for v in global_motion]) if global_motion is not None else None
    check(lib.sb2_obmc_render_ref(
        ctypes.byref(params), gm, ctypes.c_void_p(mvs.data_ptr()), ctypes.c_size_t(n),
        ctypes.byref(ref0.slab), ctypes.byref(ref1.slab) if ref1 is not None else null,
        ctypes.byref(acc.slab) if acc is not None else null, ctypes.byref(residual.slab), 1 if add else 0,
        ctypes.byref(out.slab) if out is not None else null, _stream_ptr(stream)), "sb2_obmc_render_ref")
