"""Python mirror of the drop-in C API (include/schro_b200_compat.h).

These are ctypes views of the reference's own structs (SchroFrame, SchroFrameData,
SchroParams, SchroMotion ...) and thin helpers so that tests and bench.py can call the
``schro_*`` symbols of libschro_b200.so exactly as a C caller of libschroedinger would.
No arithmetic happens here.
"""
import ctypes

import numpy as np

from ._lib import lib

FORMAT_U8_444, FORMAT_U8_422, FORMAT_U8_420 = 0x00, 0x01, 0x03
FORMAT_S16_444, FORMAT_S16_422, FORMAT_S16_420 = 0x04, 0x05, 0x07
FORMAT_S32_444, FORMAT_S32_422, FORMAT_S32_420 = 0x08, 0x09, 0x0B
CHROMA_444, CHROMA_422, CHROMA_420 = 0, 1, 2
CACHE_SIZE = 32
LIMIT_TRANSFORM_DEPTH = 6
LIMIT_BLOCK_SIZE = 64


class SchroFrameData(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int), ("data", ctypes.c_void_p), ("stride", ctypes.c_int),
                ("width", ctypes.c_int), ("height", ctypes.c_int), ("length", ctypes.c_int),
                ("h_shift", ctypes.c_int), ("v_shift", ctypes.c_int)]


class SchroFrame(ctypes.Structure):
    pass


SchroFrame._fields_ = [
    ("refcount", ctypes.c_int), ("free", ctypes.c_void_p), ("domain", ctypes.c_void_p),
    ("regions", ctypes.c_void_p * 3), ("priv", ctypes.c_void_p),
    ("format", ctypes.c_int), ("width", ctypes.c_int), ("height", ctypes.c_int),
    ("components", SchroFrameData * 3),
    ("is_virtual", ctypes.c_int), ("cached_lines", (ctypes.c_int * CACHE_SIZE) * 3),
    ("virt_frame1", ctypes.c_void_p), ("virt_frame2", ctypes.c_void_p),
    ("render_line", ctypes.c_void_p), ("virt_priv", ctypes.c_void_p), ("virt_priv2", ctypes.c_void_p),
    ("extension", ctypes.c_int), ("cache_offset", ctypes.c_int * 3),
    ("is_upsampled", ctypes.c_int), ("upsample_done", ctypes.c_uint),
]


class SchroVideoFormat(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "index", "width", "height", "chroma_format", "interlaced", "top_field_first",
        "frame_rate_numerator", "frame_rate_denominator", "aspect_ratio_numerator",
        "aspect_ratio_denominator", "clean_width", "clean_height", "left_offset", "top_offset",
        "luma_offset", "luma_excursion", "chroma_offset", "chroma_excursion", "colour_primaries",
        "colour_matrix", "transfer_function", "interlaced_coding", "unused0", "unused1", "unused2")]


class SchroGlobalMotion(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("b0", "b1", "a_exp", "a00", "a01", "a10", "a11",
                                            "c_exp", "c0", "c1")]


class SchroParams(ctypes.Structure):
    _fields_ = [
        ("video_format", ctypes.POINTER(SchroVideoFormat)), ("is_noarith", ctypes.c_int),
        ("wavelet_filter_index", ctypes.c_int), ("transform_depth", ctypes.c_int),
        ("horiz_codeblocks", ctypes.c_int * (LIMIT_TRANSFORM_DEPTH + 1)),
        ("vert_codeblocks", ctypes.c_int * (LIMIT_TRANSFORM_DEPTH + 1)),
        ("codeblock_mode_index", ctypes.c_int),
        ("num_refs", ctypes.c_int), ("have_global_motion", ctypes.c_int),
        ("xblen_luma", ctypes.c_int), ("yblen_luma", ctypes.c_int),
        ("xbsep_luma", ctypes.c_int), ("ybsep_luma", ctypes.c_int), ("mv_precision", ctypes.c_int),
        ("global_motion", SchroGlobalMotion * 2), ("picture_pred_mode", ctypes.c_int),
        ("picture_weight_bits", ctypes.c_int), ("picture_weight_1", ctypes.c_int),
        ("picture_weight_2", ctypes.c_int),
        ("is_lowdelay", ctypes.c_int), ("n_horiz_slices", ctypes.c_int), ("n_vert_slices", ctypes.c_int),
        ("slice_bytes_num", ctypes.c_int), ("slice_bytes_denom", ctypes.c_int),
        ("quant_matrix", ctypes.c_int * (3 * LIMIT_TRANSFORM_DEPTH + 1)),
        ("iwt_chroma_width", ctypes.c_int), ("iwt_chroma_height", ctypes.c_int),
        ("iwt_luma_width", ctypes.c_int), ("iwt_luma_height", ctypes.c_int),
        ("x_num_blocks", ctypes.c_int), ("y_num_blocks", ctypes.c_int),
        ("x_offset", ctypes.c_int), ("y_offset", ctypes.c_int),
    ]


class SchroMotionVector(ctypes.Structure):
    """20 bytes; ``flags`` packs pred_mode:2 using_global:1 split:2 unused:3 scan:8."""
    _fields_ = [("flags", ctypes.c_uint32), ("metric", ctypes.c_uint32),
                ("chroma_metric", ctypes.c_uint32), ("v", ctypes.c_int16 * 4)]


MV_DTYPE = np.dtype([("flags", "<u4"), ("metric", "<u4"), ("chroma_metric", "<u4"),
                     ("v", "<i2", (4,))])
assert MV_DTYPE.itemsize == 20 and ctypes.sizeof(SchroMotionVector) == 20


class SchroMotionField(ctypes.Structure):
    _fields_ = [("x_num_blocks", ctypes.c_int), ("y_num_blocks", ctypes.c_int),
                ("motion_vectors", ctypes.POINTER(SchroMotionVector))]


class SchroMotion(ctypes.Structure):
    _fields_ = [
        ("src1", ctypes.POINTER(SchroFrame)), ("src2", ctypes.POINTER(SchroFrame)),
        ("motion_vectors", ctypes.POINTER(SchroMotionVector)), ("params", ctypes.POINTER(SchroParams)),
        ("ref_weight_precision", ctypes.c_int), ("ref1_weight", ctypes.c_int), ("ref2_weight", ctypes.c_int),
        ("mv_precision", ctypes.c_int), ("xoffset", ctypes.c_int), ("yoffset", ctypes.c_int),
        ("xbsep", ctypes.c_int), ("ybsep", ctypes.c_int), ("xblen", ctypes.c_int), ("yblen", ctypes.c_int),
        ("block", SchroFrameData), ("alloc_block", SchroFrameData), ("obmc_weight", SchroFrameData),
        ("alloc_block_ref", SchroFrameData * 2), ("block_ref", SchroFrameData * 2),
        ("weight_x", ctypes.c_int * LIMIT_BLOCK_SIZE), ("weight_y", ctypes.c_int * LIMIT_BLOCK_SIZE),
        ("width", ctypes.c_int), ("height", ctypes.c_int), ("max_fast_x", ctypes.c_int),
        ("max_fast_y", ctypes.c_int), ("simple_weight", ctypes.c_uint), ("oneref_noscale", ctypes.c_uint),
    ]


class SchroHierBm(ctypes.Structure):
    _fields_ = [("ref_count", ctypes.c_int), ("ref", ctypes.c_int), ("hierarchy_levels", ctypes.c_int),
                ("params", ctypes.POINTER(SchroParams)),
                ("downsampled_src", ctypes.POINTER(ctypes.POINTER(SchroFrame))),
                ("downsampled_ref", ctypes.POINTER(ctypes.POINTER(SchroFrame))),
                ("downsampled_mf", ctypes.POINTER(ctypes.POINTER(SchroMotionField))),
                ("use_chroma", ctypes.c_uint)]


FrameP = ctypes.POINTER(SchroFrame)
FrameDataP = ctypes.POINTER(SchroFrameData)
ParamsP = ctypes.POINTER(SchroParams)


def _proto(name, restype, argtypes):
    fn = getattr(lib, name, None)
    if fn is None:
        return
    fn.restype = restype
    fn.argtypes = argtypes


_proto("schro_init", None, [])
_proto("schro_b200_set_device", None, [ctypes.c_int])
_proto("schro_b200_thread_release", None, [])
_proto("schro_b200_thread_sync", None, [])
_proto("schro_memory_domain_new_cuda", ctypes.c_void_p, [])
_proto("schro_memory_domain_new_pinned", ctypes.c_void_p, [])
_proto("schro_memory_domain_free", None, [ctypes.c_void_p])
_proto("schro_frame_new_and_alloc_full", FrameP,
       [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_proto("schro_frame_unref", None, [FrameP])
_proto("schro_frame_ref", FrameP, [FrameP])
_proto("schro_frame_to_gpu", None, [FrameP, FrameP])
_proto("schro_gpuframe_to_cpu", None, [FrameP, FrameP])
_proto("schro_wavelet_transform_2d", None, [FrameDataP, ctypes.c_int, ctypes.c_void_p])
_proto("schro_wavelet_inverse_transform_2d", None, [FrameDataP, FrameDataP, ctypes.c_int, ctypes.c_void_p])
_proto("schro_b200_frame_dequantise", None, [FrameP, ParamsP, ctypes.c_void_p])
_proto("schro_b200_frame_dequantise_widen", None, [FrameP, FrameP, ParamsP, ctypes.c_void_p])
_proto("schro_frame_iwt_transform", None, [FrameP, ParamsP])
_proto("schro_frame_inverse_iwt_transform", None, [FrameP, ParamsP])
_proto("schro_frame_downsample", None, [FrameP, FrameP])
_proto("schro_frame_convert", None, [FrameP, FrameP])
_proto("schro_frame_add", None, [FrameP, FrameP])
_proto("schro_frame_subtract", None, [FrameP, FrameP])
_proto("schro_frame_upsample_horiz", None, [FrameDataP, FrameDataP])
_proto("schro_frame_upsample_vert", None, [FrameDataP, FrameDataP])
_proto("schro_frame_mc_edgeextend", None, [FrameP])
_proto("schro_upsampled_frame_upsample", None, [FrameP])
_proto("schro_motion_new", ctypes.POINTER(SchroMotion), [ParamsP, FrameP, FrameP])
_proto("schro_motion_free", None, [ctypes.POINTER(SchroMotion)])
_proto("schro_motion_render", None, [ctypes.POINTER(SchroMotion), FrameP, FrameP, ctypes.c_int, FrameP])
_proto("schro_motion_render_u8", None, [ctypes.POINTER(SchroMotion), FrameP, FrameP, ctypes.c_int, FrameP])
_proto("schro_metric_scan_setup", None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_proto("schro_metric_scan_do_scan", None, [ctypes.c_void_p])
_proto("schro_metric_scan_get_min", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p])
_proto("schro_metric_info_init", None, [ctypes.c_void_p, FrameP, FrameP, ctypes.c_int, ctypes.c_int])
_proto("schro_metric_fast_block", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_proto("schro_frame_dup_full", FrameP, [FrameP, ctypes.c_int, ctypes.c_int])
_proto("schro_metric_absdiff_u8", ctypes.c_int,
       [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_proto("schro_hbm_new_from_frames", ctypes.POINTER(SchroHierBm),
       [ParamsP, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.POINTER(FrameP), ctypes.POINTER(FrameP)])
_proto("schro_hbm_unref", None, [ctypes.POINTER(SchroHierBm)])
_proto("schro_hbm_scan", None, [ctypes.POINTER(SchroHierBm)])
_proto("schro_hierarchical_bm_scan_hint", None, [ctypes.POINTER(SchroHierBm), ctypes.c_int, ctypes.c_int])
_proto("schro_hbm_motion_field", ctypes.POINTER(SchroMotionField), [ctypes.POINTER(SchroHierBm), ctypes.c_int])



class SchroRoughME(ctypes.Structure):
    """schroedinger/schromotionest.h:50-55"""
    _fields_ = [("encoder_frame", ctypes.c_void_p), ("ref_frame", ctypes.c_void_p),
                ("motion_fields", ctypes.POINTER(SchroMotionField) * 8)]


_proto("schro_rough_me_new_from_frames", ctypes.POINTER(SchroRoughME),
       [ctypes.c_void_p, ctypes.c_void_p, ParamsP, ctypes.c_int, ctypes.c_int, ctypes.POINTER(FrameP),
        ctypes.POINTER(FrameP)])
_proto("schro_rough_me_free", None, [ctypes.POINTER(SchroRoughME)])
_proto("schro_rough_me_heirarchical_scan", None, [ctypes.POINTER(SchroRoughME)])
_proto("schro_rough_me_heirarchical_scan_nohint", None, [ctypes.POINTER(SchroRoughME), ctypes.c_int, ctypes.c_int])
_proto("schro_rough_me_heirarchical_scan_hint", None, [ctypes.POINTER(SchroRoughME), ctypes.c_int, ctypes.c_int])

_proto("schro_frame_shift_left", None, [FrameP, ctypes.c_int])
_proto("schro_frame_shift_right", None, [FrameP, ctypes.c_int])
_proto("schro_frame_md5", None, [FrameP, ctypes.POINTER(ctypes.c_uint32)])
_proto("schro_b200_frame_inverse_iwt_combine", None, [FrameP, FrameP, ParamsP, ctypes.c_int])
_proto("schro_b200_decode_lowdelay_pictures", None, [ParamsP, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                                     ctypes.POINTER(FrameP), ctypes.c_int, ctypes.c_int])
_proto("schro_b200_decode_lowdelay_transform_data", None, [ParamsP, ctypes.c_void_p, ctypes.c_int, FrameP])
_proto("schro_motion_field_new", ctypes.POINTER(SchroMotionField), [ctypes.c_int, ctypes.c_int])
_proto("schro_motion_field_free", None, [ctypes.POINTER(SchroMotionField)])
_proto("schro_b200_motion_predict_subpel_deep", None,
       [ParamsP, ctypes.c_double, FrameP, ctypes.POINTER(FrameP), ctypes.POINTER(ctypes.POINTER(SchroMotionField))])

_proto("schro_b200_mode_decision_split2", None,
       [ParamsP, ctypes.c_double, FrameP, ctypes.POINTER(FrameP), ctypes.POINTER(ctypes.POINTER(SchroMotionField)),
        ctypes.POINTER(SchroMotion), ctypes.c_void_p, ctypes.c_void_p])

_NP = {0x00: np.uint8, 0x04: np.int16, 0x08: np.int32}


def pinned_domain():
    return ctypes.c_void_p(lib.schro_memory_domain_new_pinned())


def cuda_domain():
    return ctypes.c_void_p(lib.schro_memory_domain_new_cuda())


def frame_new_and_alloc(domain, fmt, width, height, extension=0, upsampled=0):
    return lib.schro_frame_new_and_alloc_full(domain, fmt, width, height, extension, upsampled)


def frame_plane(frame, comp, phase=0, with_border=False):
    """numpy view of one plane of a HOST frame (pageable or pinned)."""
    f = frame.contents
    c = f.components[comp]
    dt = np.dtype(_NP[f.format & 0xC])
    ext = f.extension if with_border else 0
    w, h = c.width + 2 * ext, c.height + 2 * ext
    addr = c.data + ((c.stride >> 2) * phase if f.is_upsampled else 0)
    addr -= ext * c.stride + ext * dt.itemsize
    return _strided(addr, dt, h, w, c.stride)


def _strided(addr, dt, h, w, stride):
    nbytes = (h - 1) * stride + w * dt.itemsize
    raw = (ctypes.c_uint8 * nbytes).from_address(addr)
    return np.ndarray(shape=(h, w), dtype=dt, buffer=raw, strides=(stride, dt.itemsize))


def frame_data(array, fmt=None):
    """SchroFrameData describing a 2-D numpy array (host memory)."""
    fd = SchroFrameData()
    if fmt is None:
        fmt = {np.dtype(np.uint8): FORMAT_U8_444, np.dtype(np.int16): FORMAT_S16_444,
               np.dtype(np.int32): FORMAT_S32_444}[array.dtype]
    fd.format = fmt
    fd.data = array.ctypes.data
    fd.stride = array.strides[0]
    fd.width = array.shape[1]
    fd.height = array.shape[0]
    fd.length = array.strides[0] * array.shape[0]
    return fd


def make_video_format(width, height, chroma=CHROMA_420):
    vf = SchroVideoFormat()
    vf.width, vf.height, vf.chroma_format = width, height, chroma
    return vf


def make_params(width, height, wavelet_filter_index=0, transform_depth=4, iwt_luma_width=None,
                iwt_luma_height=None, chroma=CHROMA_420, **mc):
    """A SchroParams with the calculated sizes filled in as schro_params_calculate_iwt_sizes /
    _mc_sizes do (schroedinger/schroparams.c:125-181)."""
    p = SchroParams()
    vf = make_video_format(width, height, chroma)
    p._vf = vf  # keep alive
    p.video_format = ctypes.pointer(vf)
    p.wavelet_filter_index = wavelet_filter_index
    p.transform_depth = transform_depth
    rnd = (1 << transform_depth) - 1
    hs = 0 if chroma == CHROMA_444 else 1
    vs = 1 if chroma == CHROMA_420 else 0
    cw, ch = (width + hs) >> hs, (height + vs) >> vs
    p.iwt_luma_width = iwt_luma_width or ((width + rnd) & ~rnd)
    p.iwt_luma_height = iwt_luma_height or ((height + rnd) & ~rnd)
    p.iwt_chroma_width = (p.iwt_luma_width >> hs) if iwt_luma_width else ((cw + rnd) & ~rnd)
    p.iwt_chroma_height = (p.iwt_luma_height >> vs) if iwt_luma_height else ((ch + rnd) & ~rnd)
    p.num_refs = mc.get("num_refs", 1)
    p.xblen_luma = mc.get("xblen", 12)
    p.yblen_luma = mc.get("yblen", 12)
    p.xbsep_luma = mc.get("xbsep", 8)
    p.ybsep_luma = mc.get("ybsep", 8)
    p.mv_precision = mc.get("mv_precision", 2)
    p.picture_weight_bits = mc.get("weight_bits", 1)
    p.picture_weight_1 = mc.get("weight1", 1)
    p.picture_weight_2 = mc.get("weight2", 1)
    p.x_num_blocks = 4 * ((width + 4 * p.xbsep_luma - 1) // (4 * p.xbsep_luma))
    p.y_num_blocks = 4 * ((height + 4 * p.ybsep_luma - 1) // (4 * p.ybsep_luma))
    p.x_offset = (p.xblen_luma - p.xbsep_luma) // 2
    p.y_offset = (p.yblen_luma - p.ybsep_luma) // 2
    return p
