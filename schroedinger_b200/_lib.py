"""ctypes loader for libschro_b200.so (the C-ABI declared in include/schro_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB2_LIB") or os.path.join(_HERE, "libschro_b200.so")   # (SB2_LIB: development builds)

SB2_MAX_COMPONENTS = 4


class Sb2Error(RuntimeError):
    pass


class Slab(ctypes.Structure):
    """Mirror of ``sb2_slab`` (include/schro_b200.h)."""
    _fields_ = [
        ("base", ctypes.c_void_p),
        ("picture_pitch", ctypes.c_size_t),
        ("count", ctypes.c_int),
        ("ncomp", ctypes.c_int),
        ("offset", ctypes.c_size_t * SB2_MAX_COMPONENTS),
        ("stride", ctypes.c_int * SB2_MAX_COMPONENTS),
        ("width", ctypes.c_int * SB2_MAX_COMPONENTS),
        ("height", ctypes.c_int * SB2_MAX_COMPONENTS),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise Sb2Error(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
            "There is no CPU fallback for the picture core.")
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)


lib = _load()

lib.sb2_last_error.restype = ctypes.c_char_p
lib.sb2_version.restype = ctypes.c_int
lib.sb2_launch_count.restype = ctypes.c_ulonglong
lib.sb2_iwt_workspace_bytes.restype = ctypes.c_size_t
lib.sb2_iwt_workspace_bytes.argtypes = [ctypes.POINTER(Slab), ctypes.c_int, ctypes.c_int, ctypes.c_int]
for _n in ("sb2_iwt_forward", "sb2_iwt_inverse"):
    _f = getattr(lib, _n)
    _f.restype = ctypes.c_int
    _f.argtypes = [ctypes.POINTER(Slab), ctypes.POINTER(Slab), ctypes.c_int, ctypes.c_int,
                   ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]


def _opt(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes


_SP = ctypes.POINTER(Slab)
_opt("sb2_mc_edgeextend", ctypes.c_int, [_SP, ctypes.c_int, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_upsample", ctypes.c_int, [_SP, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_downsample", ctypes.c_int, [_SP, _SP, ctypes.c_void_p])
_opt("sb2_downsample_edgeextend", ctypes.c_int, [_SP, _SP, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_frame_convert", ctypes.c_int, [_SP, ctypes.c_int, _SP, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_frame_shift", ctypes.c_int, [_SP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_frame_add", ctypes.c_int, [_SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_void_p])


class DequantParams(ctypes.Structure):
    _fields_ = [("transform_depth", ctypes.c_int), ("horiz_codeblocks", ctypes.c_int * 7),
                ("vert_codeblocks", ctypes.c_int * 7)]


_opt("sb2_dequant_table_pairs", ctypes.c_size_t, [ctypes.POINTER(DequantParams), ctypes.c_int])
_opt("sb2_dequantise", ctypes.c_int, [_SP, ctypes.c_int, ctypes.POINTER(DequantParams), ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p])
_opt("sb2_dequantise_widen", ctypes.c_int, [_SP, _SP, ctypes.POINTER(DequantParams), ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.c_void_p])


class LowdelayParams(ctypes.Structure):
    """Mirror of sb2_lowdelay_params."""
    _fields_ = [("transform_depth", ctypes.c_int), ("n_horiz_slices", ctypes.c_int), ("n_vert_slices", ctypes.c_int),
                ("slice_bytes_num", ctypes.c_int), ("slice_bytes_denom", ctypes.c_int), ("quant_matrix", ctypes.c_int * 19),
                ("table_quant", ctypes.c_uint32 * 61), ("table_offset", ctypes.c_uint32 * 61)]


_opt("sb2_lowdelay_force_unstaged", None, [ctypes.c_int])
_opt("sb2_lowdelay_decode", ctypes.c_int, [ctypes.POINTER(LowdelayParams), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                          _SP, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_edgeextend_upsample", ctypes.c_int, [_SP, ctypes.c_int, ctypes.c_void_p])
_opt("sb2_obmc_render", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _SP, _SP, _SP,
                                      _SP, ctypes.c_int, ctypes.c_int, _SP, ctypes.c_void_p])
_opt("sb2_obmc_render_ref", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _SP, _SP, _SP,
                                          _SP, ctypes.c_int, _SP, ctypes.c_void_p])
_opt("sb2_obmc_force_kernel", None, [ctypes.c_int])
_opt("sb2_obmc_last_kernel", ctypes.c_int, [])
_opt("sb2_hbm_workspace_bytes", ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int])
_opt("sb2_hbm_force_generic", None, [ctypes.c_int])
_opt("sb2_iwt_inverse_convert", ctypes.c_int, [_SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])
_opt("sb2_iwt_force_generic", None, [ctypes.c_int])
_opt("sb2_iwt_enable_fused", None, [ctypes.c_int])
_opt("sb2_upsample_force_kernel", None, [ctypes.c_int])
_opt("sb2_upsample_last_kernel", ctypes.c_int, [])
_opt("sb2_downsample_force_kernel", None, [ctypes.c_int])
_opt("sb2_downsample_last_kernel", ctypes.c_int, [])
_opt("sb2_hbm_scan_hint", ctypes.c_int, [ctypes.c_void_p, _SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p])
_opt("sb2_rough_force_unstaged", None, [ctypes.c_int])
_opt("sb2_rough_workspace_bytes", ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int])
_opt("sb2_rough_scan_nohint", ctypes.c_int, [ctypes.c_void_p, _SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])
_opt("sb2_rough_scan_hint", ctypes.c_int, [ctypes.c_void_p, _SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p])


class SubpelParams(ctypes.Structure):
    """Mirror of sb2_subpel_params."""
    _fields_ = [("xblen", ctypes.c_int), ("yblen", ctypes.c_int), ("x_num_blocks", ctypes.c_int),
                ("y_num_blocks", ctypes.c_int), ("mv_precision", ctypes.c_int), ("ref_index", ctypes.c_int),
                ("orig_extension", ctypes.c_int), ("lambda_", ctypes.c_double)]


_opt("sb2_subpel_force_generic", None, [ctypes.c_int])
_opt("sb2_subpel_workspace_bytes", ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int])
_opt("sb2_subpel_refine", ctypes.c_int, [ctypes.POINTER(SubpelParams), _SP, _SP, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])

class Split2Params(ctypes.Structure):
    """Mirror of sb2_split2_params."""
    _fields_ = [("xblen", ctypes.c_int), ("yblen", ctypes.c_int), ("x_num_blocks", ctypes.c_int),
                ("y_num_blocks", ctypes.c_int), ("mv_precision", ctypes.c_int), ("num_refs", ctypes.c_int),
                ("chroma_h_shift", ctypes.c_int), ("chroma_v_shift", ctypes.c_int), ("orig_extension", ctypes.c_int),
                ("lambda_", ctypes.c_double)]


_opt("sb2_split2_force_generic", None, [ctypes.c_int])
_opt("sb2_split2_workspace_bytes", ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int])
_opt("sb2_split2_decide", ctypes.c_int, [ctypes.POINTER(Split2Params), _SP, _SP, _SP, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])
_opt("sb2_metric_scan", ctypes.c_int, [_SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p])
_opt("sb2_metric_block_sad3", ctypes.c_int, [_SP, _SP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p])
_opt("sb2_sad_u8", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_void_p])
_opt("sb2_profile_enable", None, [ctypes.c_int])
_opt("sb2_profile_reset", None, [])
_opt("sb2_profile_count", ctypes.c_int, [])
_opt("sb2_profile_get", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)])


def last_error():
    return lib.sb2_last_error().decode()


def check(rc, what="sb2 call"):
    if rc != 0:
        raise Sb2Error(f"{what} failed ({rc}): {last_error()}")
