"""GPU: the stand-alone metric-scan entry points -- sb2_metric_scan / sb2_metric_block_sad3 (batched, device
layer) and schro_metric_scan_setup / _do_scan / _get_min / schro_metric_info_init / schro_metric_fast_block
(drop-in layer) -- against the oracle and the golden vectors of the compiled reference, bit-exact."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests import test_oracle_metric_scan as T

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(helpers.GOLDEN_DIR, "metric_scan.npz"))


class ScanDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("picture", "x", "y", "block_width", "block_height", "ref_x", "ref_y",
                                            "scan_width", "scan_height")]


class BlockDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("picture", "x", "y", "dx", "dy")]


def test_metric_scan_batched_device(cuda):
    """all golden queries as ONE launch over a two-picture slab (the second picture swaps the roles)"""
    from schroedinger_b200 import device as dev, lib, check
    src = [GOLD[f"src{k}"] for k in range(3)]
    ref = [GOLD[f"ref{k}"] for k in range(3)]
    lay = dev.FrameLayout.yuv420("u8", T.W, T.H, 32)
    ss, rs = dev.PictureSlab(lay, 2), dev.PictureSlab(lay, 2)
    for c in range(3):
        ss.upload(0, c, src[c]); rs.upload(0, c, ref[c])
        ss.upload(1, c, ref[c]); rs.upload(1, c, src[c])
    dev.mc_edgeextend(ss); dev.mc_edgeextend(rs)
    for use_chroma in (0, 1):
        qs = [q for q in GOLD["queries"] if int(q[7]) == use_chroma]
        wants = [T.oracle_scan(src, ref, tuple(int(v) for v in q)) for q in qs]
        wants += [T.oracle_scan(ref, src, tuple(int(v) for v in q)) for q in qs]
        descs = (ScanDesc * len(wants))()
        bdescs = (BlockDesc * len(wants))()
        for k, (out, _, _) in enumerate(wants):
            q = qs[k % len(qs)]
            descs[k] = ScanDesc(k // len(qs), int(q[0]), int(q[1]), int(q[2]), int(q[3]), int(out[0]), int(out[1]),
                                int(out[2]), int(out[3]))
            bdescs[k] = BlockDesc(k // len(qs), int(q[0]), int(q[1]), int(q[4]), int(q[5]))
        dd = torch.from_numpy(np.frombuffer(bytes(descs), np.uint8).copy()).cuda()
        m = torch.zeros(len(wants) * 42 * 42, dtype=torch.int32, device="cuda")
        cm = torch.zeros(len(wants) * 42 * 42, dtype=torch.int32, device="cuda")
        check(lib.sb2_metric_scan(ctypes.byref(ss.slab), ctypes.byref(rs.slab), 1, 1, use_chroma,
                                  ctypes.c_void_p(dd.data_ptr()), len(wants), ctypes.c_void_p(m.data_ptr()),
                                  ctypes.c_void_p(cm.data_ptr()), None), "sb2_metric_scan")
        torch.cuda.synchronize()
        gm = m.cpu().numpy().view(np.uint32).reshape(len(wants), -1)
        gc = cm.cpu().numpy().view(np.uint32).reshape(len(wants), -1)
        for k, (out, wm, wc) in enumerate(wants):
            n = T.used(out)
            assert np.array_equal(gm[k][:n], wm[:n]), (use_chroma, k)
            assert np.array_equal(gc[k][:n], wc[:n]), (use_chroma, k, "chroma")
        # 3-component block SADs of the same blocks, grouped by block size (one launch per size)
        for (bw, bh) in sorted({(int(q[2]), int(q[3])) for q in qs}):
            idx = [k for k in range(len(wants)) if (int(qs[k % len(qs)][2]), int(qs[k % len(qs)][3])) == (bw, bh)]
            sub = (BlockDesc * len(idx))(*[bdescs[k] for k in idx])
            bd = torch.from_numpy(np.frombuffer(bytes(sub), np.uint8).copy()).cuda()
            res = torch.zeros(len(idx), dtype=torch.int32, device="cuda")
            check(lib.sb2_metric_block_sad3(ctypes.byref(ss.slab), ctypes.byref(rs.slab), 32, bw, bh, 1, 1,
                                            ctypes.c_void_p(bd.data_ptr()), len(idx), ctypes.c_void_p(res.data_ptr()),
                                            None), "sb2_metric_block_sad3")
            torch.cuda.synchronize()
            assert res.cpu().tolist() == [int(wants[k][0][8]) for k in idx], (bw, bh)


@pytest.mark.parametrize("domain_kind", ["malloc", "cuda"])
def test_metric_scan_drop_in(cuda, domain_kind):
    """schro_metric_scan_setup / _do_scan / _get_min and schro_metric_fast_block exactly as
    schro_hierarchical_bm_scan_hint drives them (schrohierbm.c:349-371), against the golden outputs"""
    from schroedinger_b200 import compat, lib

    class MetricScan(ctypes.Structure):
        _fields_ = [("frame", compat.FrameP), ("ref_frame", compat.FrameP), ("block_width", ctypes.c_int),
                    ("block_height", ctypes.c_int), ("x", ctypes.c_int), ("y", ctypes.c_int), ("ref_x", ctypes.c_int),
                    ("ref_y", ctypes.c_int), ("scan_width", ctypes.c_int), ("scan_height", ctypes.c_int),
                    ("gravity_scale", ctypes.c_int), ("gravity_x", ctypes.c_int), ("gravity_y", ctypes.c_int),
                    ("use_chroma", ctypes.c_int), ("metrics", ctypes.c_uint32 * (42 * 42)),
                    ("chroma_metrics", ctypes.c_uint32 * (42 * 42))]

    class MetricInfo(ctypes.Structure):
        _fields_ = [("frame", compat.FrameP), ("ref_frame", compat.FrameP), ("block_width", ctypes.c_int * 3),
                    ("block_height", ctypes.c_int * 3), ("h_shift", ctypes.c_int * 3), ("v_shift", ctypes.c_int * 3),
                    ("metric", ctypes.c_void_p), ("metric_right", ctypes.c_void_p), ("metric_bottom", ctypes.c_void_p),
                    ("metric_corner", ctypes.c_void_p)]

    dom = compat.cuda_domain() if domain_kind == "cuda" else None
    frames = []
    for name in ("src", "ref"):
        host = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, T.W, T.H, 32, 0)
        for c in range(3):
            compat.frame_plane(host, c)[...] = GOLD[f"{name}{c}"]
        lib.schro_frame_mc_edgeextend(host)
        if dom is not None:
            # schro_frame_dup_full: same domain as its source, so go through a CUDA-domain copy
            d = compat.frame_new_and_alloc(dom, compat.FORMAT_U8_420, T.W, T.H, 32, 0)
            lib.schro_frame_to_gpu(d, host)
            frames.append(d)
        else:
            frames.append(host)
    lib.schro_metric_scan_get_min.restype = ctypes.c_int
    for i, q in enumerate(GOLD["queries"]):
        x, y, bw, bh, dx, dy, dist, uc = (int(v) for v in q)
        want = GOLD[f"q{i}_out"]
        scan = MetricScan()
        scan.frame, scan.ref_frame = frames[0], frames[1]
        scan.block_width, scan.block_height, scan.x, scan.y = bw, bh, x, y
        scan.gravity_x, scan.gravity_y = dx, dy
        lib.schro_metric_scan_setup(ctypes.byref(scan), dx, dy, dist, uc)
        assert [scan.ref_x, scan.ref_y, scan.scan_width, scan.scan_height] == [int(v) for v in want[:4]], i
        lib.schro_metric_scan_do_scan(ctypes.byref(scan))
        n = scan.scan_width * scan.scan_height
        assert np.array_equal(np.ctypeslib.as_array(scan.metrics)[:n], GOLD[f"q{i}_metrics"]), i
        assert np.array_equal(np.ctypeslib.as_array(scan.chroma_metrics)[:n], GOLD[f"q{i}_chroma"]), i
        odx, ody, chroma = ctypes.c_int(dx), ctypes.c_int(dy), ctypes.c_uint32()
        best = lib.schro_metric_scan_get_min(ctypes.byref(scan), ctypes.byref(odx), ctypes.byref(ody), ctypes.byref(chroma))
        assert [odx.value, ody.value, best, chroma.value] == [int(v) for v in want[4:8]], i
        info = MetricInfo()
        lib.schro_metric_info_init(ctypes.byref(info), frames[0], frames[1], bw, bh)
        assert lib.schro_metric_fast_block(ctypes.byref(info), x, y, dx, dy) == int(want[8]), i
    for f in frames:
        lib.schro_frame_unref(f)


def test_frame_dup_full(cuda):
    """schro_frame_dup_full (schroframe.c:712): same domain / format / size, new extension and layout,
    content through schro_frame_convert"""
    from schroedinger_b200 import compat, lib
    lib.schro_frame_dup_full.restype = compat.FrameP
    rng = np.random.default_rng(1)
    w, h = 100, 60
    for dom in (None, compat.cuda_domain()):
        host = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, w, h, 0, 0)
        imgs = []
        for c in range(3):
            v = compat.frame_plane(host, c)
            v[...] = rng.integers(0, 256, size=v.shape)
            imgs.append(v.copy())
        src = host
        if dom is not None:
            src = compat.frame_new_and_alloc(dom, compat.FORMAT_U8_420, w, h, 0, 0)
            lib.schro_frame_to_gpu(src, host)
        dup = lib.schro_frame_dup_full(src, 32, 1)
        d = dup.contents
        assert (d.width, d.height, d.extension, d.is_upsampled, d.format) == (w, h, 32, 1, compat.FORMAT_U8_420)
        back = dup
        if dom is not None:
            back = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, w, h, 32, 1)
            lib.schro_gpuframe_to_cpu(back, dup)
        for c in range(3):
            assert np.array_equal(compat.frame_plane(back, c), imgs[c]), c
