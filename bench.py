#!/usr/bin/env python
"""bench.py -- throughput of the B200 picture core on BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference's CPU path, host cores

A "step" is one pass of the hot path over one batch of synthetic 2160p 4:2:0 pictures
(the workload names its stages).  `value` is frames/s with inputs resident in HBM,
`e2e` the same through the drop-in C API (schro_* symbols) with pinned HOST frames,
H2D + D2H inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# --------------------------------------------------------------------------------------
# workload definition (shapes from SURVEY.md section 8 / BASELINE.json configs)
# --------------------------------------------------------------------------------------
def workload_spec(name):
    if name == "picture_core_2160p":
        return dict(width=3840, height=2160, iwt_w=3840, iwt_h=2176, depth_name="s32", filter=6,
                    transform_depth=5, batch=32,
                    label="2160p 4:2:0 10-bit: inverse Daubechies 9/7 5-level s32 wavelet"
                          " + half-pel upsample + 1/4-pel OBMC render (2 refs, 12x12/8x8)"
                          " + 4-level hierarchical SAD block matching")
    if name == "wavelet_1080p_dd97":
        return dict(width=1920, height=1080, iwt_w=1920, iwt_h=1088, depth_name="s16", filter=0,
                    transform_depth=4, batch=64, full_core=False,
                    label="1080p 4:2:0 8-bit: inverse Deslauriers-Dubuc 9/7 4-level s16 only (BASELINE configs[1])")
    if name == "wavelet_1080p_legall":
        return dict(width=1920, height=1080, iwt_w=1920, iwt_h=1088, depth_name="s16", filter=1,
                    transform_depth=4, batch=64, full_core=False,
                    label="1080p 4:2:0 8-bit: inverse LeGall 5/3 4-level s16 only (BASELINE configs[0], inverse half)")
    if name == "picture_core_1080p":
        return dict(width=1920, height=1080, iwt_w=1920, iwt_h=1088, depth_name="s16", filter=0,
                    transform_depth=4, batch=64,
                    label="1080p 4:2:0 8-bit, all stages: inverse DD 9/7 4-level s16 + half-pel upsample + 1/4-pel OBMC "
                          "render (2 refs, 12x12/8x8; BASELINE configs[3]) + 4-level hierarchical SAD block matching")
    if name == "picture_core_cif":       # tiny, for CPU-side testing of bench.py itself
        return dict(width=352, height=288, iwt_w=352, iwt_h=288, depth_name="s32", filter=6,
                    transform_depth=5, batch=4, label="CIF test workload, all stages")
    raise SystemExit(f"unknown workload {name}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, window=None):
        """window = (t0, t1) in time.perf_counter() seconds: keep the samples that arrived while the
        timed region ran (nvidia-smi needs a few hundred ms to start, so it is launched before the
        warm-up and the samples are cut to the window afterwards)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for (when, line) in self.lines:
            if window and not (window[0] <= when <= window[1] + 0.12):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def make_coeff_frame(spec, rng):
    """One synthetic coefficient frame: uniform in [-512, 511]."""
    dt = np.int32 if spec["depth_name"] == "s32" else np.int16
    w, h = spec["iwt_w"], spec["iwt_h"]
    return [rng.integers(-512, 512, size=s, dtype=np.int64).astype(dt)
            for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]


def make_mv_field(nbx, nby, rng):
    """modes {0:12%,1:38%,2:25%,3:25%}, vectors uniform +-64 quarter-pel
    (+1% outliers +-4000), DC +-127."""
    n = nbx * nby
    mv = np.zeros(n, dtype=np.dtype([("flags", "<u4"), ("metric", "<u4"), ("chroma_metric", "<u4"),
                                     ("v", "<i2", (4,))]))
    mode = rng.choice(4, size=n, p=[0.12, 0.38, 0.25, 0.25])
    v = rng.integers(-64, 65, size=(n, 4))
    out = rng.random(n) < 0.01
    v[out] = rng.integers(-4000, 4001, size=(int(out.sum()), 4))
    dc = rng.integers(-127, 128, size=(n, 4))
    dc[:, 3] = 0
    mv["v"] = np.where((mode == 0)[:, None], dc, v)
    mv["flags"] = mode.astype(np.uint32)
    return mv


CONTENT = "natural"


_CANVAS = {}
CANVAS_MARGIN = 32


def _texture_canvas(width, height):
    """The noise-free luma texture on a canvas CANVAS_MARGIN pixels larger than the picture on every side
    (computed once per size: panned pictures are crops of it)."""
    key = (width, height, CONTENT)
    if key not in _CANVAS:
        m = CANVAS_MARGIN
        yy, xx = np.mgrid[-m:height + m, -m:width + m]
        xx = xx.astype(np.float32)
        yy = yy.astype(np.float32)
        if CONTENT == "periodic":
            y = 128 + 50 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 30 * np.sin((xx + 2 * yy) / 31.0)
        else:
            wrng = np.random.default_rng(424242)                 # the texture itself is the same for every frame
            lam = np.exp(wrng.uniform(np.log(24.0), np.log(600.0), 20))
            ang = wrng.uniform(0, 2 * np.pi, 20)
            ph = wrng.uniform(0, 2 * np.pi, 20)
            amp = lam ** 0.6
            amp *= 45.0 / np.sqrt(0.5 * np.sum(amp ** 2))
            y = np.full(xx.shape, 128.0, np.float32)
            for k in range(20):
                kx, ky = 2 * np.pi / lam[k] * np.cos(ang[k]), 2 * np.pi / lam[k] * np.sin(ang[k])
                y += np.float32(amp[k]) * np.sin(np.float32(kx) * xx + np.float32(ky) * yy + np.float32(ph[k]))
        _CANVAS[key] = y.astype(np.float32)
    return _CANVAS[key]


def textured_frame(width, height, rng, pan=(0, 0)):
    """One u8 4:2:0 picture: smooth texture + per-frame noise, optionally panned.

    "natural" (default): 20 plane waves of random direction, wavelengths log-uniform in 24..600 px and
    amplitude rising with wavelength -- no repeat inside any search window, so block matching finds the pan.
    "periodic": the round-1a texture, one 56 x 44 px cell repeated; its aliases at the coarse pyramid levels
    make the motion field incoherent (a worst case for the block matcher, kept for DESIGN.md's comparison)."""
    m = CANVAS_MARGIN
    assert abs(pan[0]) <= m and abs(pan[1]) <= m
    canvas = _texture_canvas(width, height)
    y = canvas[m + pan[1]:m + pan[1] + height, m + pan[0]:m + pan[0] + width]
    y = y + rng.integers(-12, 13, size=(height, width))
    y = np.clip(y, 0, 255).astype(np.uint8)
    c = y[::2, ::2]
    return [y, np.ascontiguousarray(255 - c), np.ascontiguousarray((c // 2 + 64).astype(np.uint8))]


def picture_pan(i):
    """Every picture of a batch moves differently against its reference: (5,3) +- (2,1) pixels."""
    return (5 + (i % 5) - 2, 3 + (i % 3) - 1)


HBM_LEVELS = 4
BLOCK = dict(xbsep=8, ybsep=8, xblen=12, yblen=12, prec=2)


def block_counts(width, height):
    return (4 * ((width + 4 * BLOCK["xbsep"] - 1) // (4 * BLOCK["xbsep"])),
            4 * ((height + 4 * BLOCK["ybsep"] - 1) // (4 * BLOCK["ybsep"])))


class Stages:
    """Device-resident stages of one step.  Each stage launches on the current stream."""

    def __init__(self, spec, torch, dev):
        self.spec, self.torch, self.dev = spec, torch, dev
        self.stages = []
        B = spec["batch"]
        W, H = spec["width"], spec["height"]
        rng = np.random.default_rng(2026)
        # ---- stage 1: inverse wavelet (BASELINE config 2/3) ----
        lay = dev.FrameLayout.yuv420(spec["depth_name"], spec["iwt_w"], spec["iwt_h"])
        self.coef = dev.PictureSlab(lay, B, zero=False)
        self.recon = dev.PictureSlab(lay, B, zero=False)
        for p in range(B):                                   # every picture of the batch is different
            planes = make_coeff_frame(spec, np.random.default_rng(2026 + p))
            for c in range(3):
                self.coef.upload(p, c, planes[c])
        self.ws = dev.Workspace()
        ncoef = sum(w * h for w, h in lay.comp_sizes)
        self.working_set = self.coef.nbytes + self.recon.nbytes
        self.stages.append(dict(
            name="iwt_inverse", frames=B, alg_bytes=2.0 * ncoef * lay.bpp * B,
            run=lambda: dev.iwt_inverse(self.coef, self.recon, spec["filter"],
                                        spec["transform_depth"], self.ws)))
        if spec.get("also_forward"):
            self.stages.append(dict(
                name="iwt_forward", frames=B, alg_bytes=2.0 * ncoef * lay.bpp * B,
                run=lambda: dev.iwt_forward(self.recon, self.coef, spec["filter"],
                                            spec["transform_depth"], self.ws)))
        self.me_stream = None
        if not spec.get("full_core", True):
            self.me_stream = torch.cuda.Stream()
            self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
            return
        # ---- stage 2: OBMC render, residual = the wavelet output, 2 upsampled references ----
        npix = W * H * 3 // 2
        ref_lay = dev.FrameLayout.yuv420("u8", W, H, 32, True)
        self.refs = [dev.PictureSlab(ref_lay, B) for _ in range(2)]
        self.newref = dev.PictureSlab(ref_lay, B)          # the decoded picture = next reference
        for r, slab in enumerate(self.refs):
            for p in range(B):
                fr = textured_frame(W, H, np.random.default_rng(3000 + 100 * r + p), pan=(5 * r, 3 * r))
                for c in range(3):
                    slab.upload(p, c, fr[c])
            dev.mc_edgeextend(slab)
            dev.upsample(slab)
        nbx, nby = block_counts(W, H)
        self.nblocks = nbx * nby
        self.mvs = torch.from_numpy(np.concatenate(
            [make_mv_field(nbx, nby, np.random.default_rng(4000 + p)).view(np.uint8) for p in range(B)])).cuda()
        self.obmc_params = dev.ObmcParams(BLOCK["xbsep"], BLOCK["ybsep"], BLOCK["xblen"], BLOCK["yblen"],
                                          nbx, nby, BLOCK["prec"], 1, 1, 1, 1, 1)
        self.resid = dev.SlabView(self.recon, [(W, H), (W // 2, H // 2), (W // 2, H // 2)])
        bpp = lay.bpp
        self.stages.append(dict(
            name="obmc_render", frames=B,
            alg_bytes=(npix * (2 * 4 + bpp + 1) + self.nblocks * 20.0) * B,
            run=lambda: dev.obmc_render(self.obmc_params, self.mvs, self.refs[0], self.refs[1],
                                        self.resid, 1, out=self.newref)))
        # ---- stage 3: the decoded picture becomes a reference: edge-extend + half-pel upsample
        self.stages.append(dict(
            name="upsample", frames=B, alg_bytes=4.0 * npix * B,
            run=lambda: dev.edgeextend_upsample(self.newref)))
        # ---- stage 4: motion estimation: pyramids + hierarchical block matching vs ref 0 ----
        self.src_pyr = dev.Pyramid(W, H, B, HBM_LEVELS, 8)
        self.ref_pyr = dev.Pyramid(W, H, B, HBM_LEVELS, 8)
        for p in range(B):                                   # own noise and own pan for every pair
            srcf = textured_frame(W, H, np.random.default_rng(5000 + p), pan=picture_pan(p))
            base = textured_frame(W, H, np.random.default_rng(6000 + p))
            for pyr, fr in ((self.src_pyr, srcf), (self.ref_pyr, base)):
                for c in range(3):
                    pyr.slabs[0].upload(p, c, fr[c])
        self.hbm_params = dev.HbmParams(BLOCK["xbsep"], BLOCK["ybsep"], nbx, nby, 0, 0, 1, 1)
        self.fields = [torch.empty(B * self.nblocks * 20, dtype=torch.uint8, device="cuda")
                       for _ in range(HBM_LEVELS + 1)]
        pyr_bytes = npix * (1 + 0.25 + 1 / 16 + 1 / 64) + npix * (0.25 + 1 / 16 + 1 / 64 + 1 / 256)
        self.stages.append(dict(
            name="pyramid", frames=B, alg_bytes=2 * pyr_bytes * B,
            run=lambda: (self.src_pyr.build(), self.ref_pyr.build())))
        self.stages.append(dict(
            name="hier_block_match", frames=B,
            alg_bytes=(2 * npix * 1.332 + self.nblocks * 20 * 1.34) * B,
            run=lambda: dev.hbm_scan(self.hbm_params, self.src_pyr, self.ref_pyr, 3, self.fields, self.ws)))
        self.working_set += sum(s.nbytes for s in self.refs) + self.newref.nbytes

        self.me_stream = torch.cuda.Stream()
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()

    def step(self):
        """Decode-side stages on the current stream, motion-estimation stages (pyramids +
        hierarchical block matching: latency-bound wavefronts that leave most of the machine
        idle) concurrently on a second stream; the step ends when both have finished."""
        torch = self.torch
        cur = torch.cuda.current_stream()
        me = [s for s in self.stages if s["name"] in ("pyramid", "hier_block_match")]
        if not me or not self.spec.get("overlap", False):
            for s in self.stages:
                s["run"]()
            return
        self.ev_fork.record(cur)
        self.me_stream.wait_event(self.ev_fork)
        with torch.cuda.stream(self.me_stream):
            for s in me:
                s["run"]()
            self.ev_join.record(self.me_stream)
        for s in self.stages:
            if s not in me:
                s["run"]()
        cur.wait_event(self.ev_join)


def collect_profile(lib):
    n = lib.sb2_profile_count()
    recs = {}
    tag = ctypes.create_string_buffer(64)
    ms = ctypes.c_float()
    by = ctypes.c_double()
    for i in range(n):
        if lib.sb2_profile_get(i, tag, 64, ctypes.byref(ms), ctypes.byref(by)) != 0:
            continue
        r = recs.setdefault(tag.value.decode(), dict(ms=0.0, bytes=0.0, launches=0))
        r["ms"] += ms.value
        r["bytes"] += by.value
        r["launches"] += 1
    return recs


class HostFrames:
    """The e2e arm: the same step through the drop-in C API (schro_* symbols of
    libschro_b200.so) on pinned HOST frames, the way a decoder/encoder would drive it:
    per picture the coefficient frame and the source picture go H2D, the decoded picture and
    the motion fields come back D2H; reference pictures stay in the CUDA memory domain
    (as with the reference's own use_cuda path, schrodecoder.c:1731-1736)."""

    def __init__(self, spec, lib, nthreads, widen=True):
        """widen: the coefficients travel QUANTISED as s16 and are dequantised into the s32 coefficient
        frame on the device (schro_b200_frame_dequantise_widen) -- half the upload of the s32 frame."""
        from schroedinger_b200 import compat
        self.compat, self.lib, self.spec = compat, lib, spec
        self.nthreads = nthreads
        self.full = spec.get("full_core", True)
        self.widen = bool(widen and self.full and spec["depth_name"] == "s32")
        B = spec["batch"]
        W, H = spec["width"], spec["height"]
        import torch
        lib.schro_b200_set_device(torch.cuda.current_device())
        self.pinned = compat.pinned_domain()
        self.cuda = compat.cuda_domain()
        s32 = spec["depth_name"] == "s32"
        cfmt = compat.FORMAT_S32_420 if s32 else compat.FORMAT_S16_420
        rng = np.random.default_rng(7)
        A = compat.frame_new_and_alloc
        self.params = compat.make_params(W, H, spec["filter"], spec["transform_depth"], spec["iwt_w"],
                                         spec["iwt_h"], num_refs=2, **{k: BLOCK[k] for k in ("xbsep", "ybsep", "xblen", "yblen")},
                                         mv_precision=BLOCK["prec"])
        self.coef_host, self.coef16_host, self.pairs = [], [], None
        for i in range(B):
            planes = make_coeff_frame(spec, np.random.default_rng(7000 + i))       # every picture differs
            if self.widen:
                # quantised values; with quantiser index 12 (factor 32: 8 x the value) the dequantised
                # coefficients span the same +-512 as the s32 frames of the other leg
                f = A(self.pinned, compat.FORMAT_S16_420, spec["iwt_w"], spec["iwt_h"])
                for c in range(3):
                    compat.frame_plane(f, c)[...] = (planes[c] >> 3).astype(np.int16)
                self.coef16_host.append(f)
            else:
                f = A(self.pinned, cfmt, spec["iwt_w"], spec["iwt_h"])
                for c in range(3):
                    compat.frame_plane(f, c)[...] = planes[c]
                self.coef_host.append(f)
        if self.widen:
            gold = np.load(os.path.join(ROOT, "tests", "golden", "dequant.npz"))
            depth = spec["transform_depth"]
            for l in range(depth + 1):
                self.params.horiz_codeblocks[l] = self.params.vert_codeblocks[l] = 1 if l < 2 else min(8, 1 << (l - 1))
            from schroedinger_b200._lib import DequantParams
            dp = DequantParams()
            dp.transform_depth = depth
            for l in range(7):
                dp.horiz_codeblocks[l] = self.params.horiz_codeblocks[l] if l <= depth else 1
                dp.vert_codeblocks[l] = self.params.vert_codeblocks[l] if l <= depth else 1
            npairs = lib.sb2_dequant_table_pairs(ctypes.byref(dp), 3)
            self.pairs = np.empty((npairs, 2), np.int32)
            self.pairs[:, 0] = int(gold["table_quant"][12])
            self.pairs[:, 1] = int(gold["table_offset_1_2"][12]) + 2
        bpp = 4 if s32 else 2
        self.h2d = int(spec["iwt_w"] * spec["iwt_h"] * 1.5 * bpp)
        self.d2h = self.h2d
        if not self.full:
            return
        npix = W * H * 3 // 2
        nb = self.params.x_num_blocks * self.params.y_num_blocks
        coef_bytes = int(spec["iwt_w"] * spec["iwt_h"] * 1.5 * (2 if self.widen else bpp))
        self.h2d = coef_bytes + (self.pairs.nbytes if self.widen else 0) + npix + nb * 20
        self.d2h = npix + nb * 20                       # decoded picture + the level-0 motion field
        # reference pictures and the reference pyramid live on the device
        self.refs = []
        for r in range(2):
            fr = textured_frame(W, H, rng, pan=(5 * r, 3 * r))
            hf = A(None, compat.FORMAT_U8_420, W, H, 32, 1)
            for c in range(3):
                compat.frame_plane(hf, c)[...] = fr[c]
            lib.schro_frame_mc_edgeextend(hf)
            lib.schro_upsampled_frame_upsample(hf)
            df = A(self.cuda, compat.FORMAT_U8_420, W, H, 32, 1)
            lib.schro_frame_to_gpu(df, hf)
            lib.schro_frame_unref(hf)
            self.refs.append(df)
        self.ref_pyr = self._pyramid_frames()
        basef = textured_frame(W, H, rng)
        hf = A(None, compat.FORMAT_U8_420, W, H, 32, 0)
        for c in range(3):
            compat.frame_plane(hf, c)[...] = basef[c]
        lib.schro_frame_to_gpu(self.ref_pyr[0], hf)
        lib.schro_frame_unref(hf)
        self._build_pyramid(self.ref_pyr)
        self.src_host, self.out_host = [], []
        for i in range(B):
            srcf = textured_frame(W, H, np.random.default_rng(7100 + i), pan=picture_pan(i))
            f = A(self.pinned, compat.FORMAT_U8_420, W, H, 0, 0)
            for c in range(3):
                compat.frame_plane(f, c)[...] = srcf[c]
            self.src_host.append(f)
            self.out_host.append(A(self.pinned, compat.FORMAT_U8_420, W, H, 0, 0))
        mv = make_mv_field(self.params.x_num_blocks, self.params.y_num_blocks, rng)
        # per-thread device frames
        self.th = []
        for _ in range(nthreads):
            t = {}
            t["coef"] = A(self.cuda, cfmt, spec["iwt_w"], spec["iwt_h"])
            if self.widen:
                t["coef16"] = A(self.cuda, compat.FORMAT_S16_420, spec["iwt_w"], spec["iwt_h"])
            t["acc"] = A(self.cuda, compat.FORMAT_S16_420, W, H)
            t["out"] = A(self.cuda, compat.FORMAT_U8_420, W, H, 32, 1)
            t["motion"] = lib.schro_motion_new(ctypes.byref(self.params), self.refs[0], self.refs[1])
            ctypes.memmove(t["motion"].contents.motion_vectors, mv.ctypes.data, mv.nbytes)
            t["src_pyr"] = self._pyramid_frames()
            self.th.append(t)

    def _pyramid_frames(self):
        A, compat, W, H = self.compat.frame_new_and_alloc, self.compat, self.spec["width"], self.spec["height"]
        frames = [A(self.cuda, compat.FORMAT_U8_420, W, H, 32, 0)]
        w, h = W, H
        for _ in range(HBM_LEVELS):
            w, h = (w + 1) // 2, (h + 1) // 2
            frames.append(A(self.cuda, compat.FORMAT_U8_420, w, h, 8, 0))
        return frames

    def _build_pyramid(self, frames):
        lib = self.lib
        lib.schro_frame_mc_edgeextend(frames[0])
        for l in range(HBM_LEVELS):
            lib.schro_frame_downsample(frames[l + 1], frames[l])
            lib.schro_frame_mc_edgeextend(frames[l + 1])

    def picture(self, t, i):
        lib, th = self.lib, self.th[t] if self.full else None
        if not self.full:
            lib.schro_frame_inverse_iwt_transform(self.coef_host[i], ctypes.byref(self.params))
            return
        # decode side: coefficients in, decoded picture out
        if self.widen:
            lib.schro_frame_to_gpu(th["coef16"], self.coef16_host[i])
            lib.schro_b200_frame_dequantise_widen(th["coef"], th["coef16"], ctypes.byref(self.params),
                                                  self.pairs.ctypes.data_as(ctypes.c_void_p))
        else:
            lib.schro_frame_to_gpu(th["coef"], self.coef_host[i])
        lib.schro_frame_inverse_iwt_transform(th["coef"], ctypes.byref(self.params))
        # dest (picture size) gives the rendered area, the iwt-padded coefficient frame is the addframe
        lib.schro_motion_render(th["motion"], th["acc"], th["coef"], 1, th["out"])
        lib.schro_frame_mc_edgeextend(th["out"])
        th["out"].contents.upsample_done = 0
        lib.schro_upsampled_frame_upsample(th["out"])
        lib.schro_gpuframe_to_cpu(self.out_host[i], th["out"])
        # encode side: source picture in, motion fields out
        lib.schro_frame_to_gpu(th["src_pyr"][0], self.src_host[i])
        self._build_pyramid(th["src_pyr"])
        arr = self.compat.FrameP * (HBM_LEVELS + 1)
        hbm = lib.schro_hbm_new_from_frames(ctypes.byref(self.params), 0, HBM_LEVELS, 0,
                                            arr(*th["src_pyr"]), arr(*self.ref_pyr))
        lib.schro_hbm_scan(hbm)
        lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
        lib.schro_hbm_motion_field(hbm, 0)               # the host reads the finest field
        lib.schro_hbm_unref(hbm)

    def native_start(self):
        """Hand the frames to the pthread driver (bench_native/e2e_driver.c): from here on the
        host side of a step is plain C calling the drop-in API, as a C application would."""
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_native", "libsb2_e2e_driver.so")
        if not os.path.exists(path):
            raise RuntimeError(path + " is missing: run `make` (or __graft_entry__.build())")
        drv = ctypes.CDLL(path)
        compat = self.compat
        FP, n, T = compat.FrameP, self.spec["batch"], self.nthreads

        class Job(ctypes.Structure):
            _fields_ = [("nthreads", ctypes.c_int), ("npictures", ctypes.c_int), ("levels", ctypes.c_int),
                        ("pic_height", ctypes.c_int), ("full_core", ctypes.c_int),
                        ("params", ctypes.POINTER(compat.SchroParams)),
                        ("coef_host", ctypes.POINTER(FP)), ("src_host", ctypes.POINTER(FP)),
                        ("out_host", ctypes.POINTER(FP)), ("ref_pyr", ctypes.POINTER(FP)),
                        ("coef_dev", ctypes.POINTER(FP)), ("acc_dev", ctypes.POINTER(FP)),
                        ("out_dev", ctypes.POINTER(FP)),
                        ("motion", ctypes.POINTER(ctypes.POINTER(compat.SchroMotion))),
                        ("src_pyr", ctypes.POINTER(FP)), ("widen", ctypes.c_int),
                        ("coef16_host", ctypes.POINTER(FP)), ("coef16_dev", ctypes.POINTER(FP)),
                        ("pairs", ctypes.c_void_p)]

        def arr(frames):
            a = (FP * max(1, len(frames)))(*frames)
            self._keep.append(a)
            return a

        self._keep = []
        job = Job()
        job.nthreads, job.npictures, job.levels = T, n, HBM_LEVELS
        job.pic_height, job.full_core = self.spec["height"], int(self.full)
        job.params = ctypes.pointer(self.params)
        job.coef_host = arr(self.coef_host)
        job.widen = int(self.widen)
        if self.widen:
            job.coef16_host = arr(self.coef16_host)
            job.coef16_dev = arr([t["coef16"] for t in self.th])
            job.pairs = self.pairs.ctypes.data
        if self.full:
            job.src_host, job.out_host, job.ref_pyr = arr(self.src_host), arr(self.out_host), arr(self.ref_pyr)
            job.coef_dev = arr([t["coef"] for t in self.th])
            job.acc_dev = arr([t["acc"] for t in self.th])
            job.out_dev = arr([t["out"] for t in self.th])
            mo = (ctypes.POINTER(compat.SchroMotion) * T)(*[t["motion"] for t in self.th])
            self._keep.append(mo)
            job.motion = mo
            job.src_pyr = arr([f for t in self.th for f in t["src_pyr"]])
        drv.sb2_e2e_step.restype = ctypes.c_double
        drv.sb2_e2e_run.restype = ctypes.c_double
        drv.sb2_e2e_run.argtypes = [ctypes.c_int]
        if drv.sb2_e2e_start(ctypes.byref(job)) != 0:
            raise RuntimeError("sb2_e2e_start failed")
        self._keep.append(job)
        self.drv = drv

    def step(self):
        """One e2e step: every picture of the batch, pictures spread over a pool of persistent
        host threads (the reference's own threading model, schroasync-pthread.c); each thread
        owns a stream and its staging buffers."""
        if hasattr(self, "drv"):
            self.drv.sb2_e2e_step()
            return
        n = self.spec["batch"]
        if not hasattr(self, "pool"):
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(max_workers=self.nthreads)

        def work(tid):
            for i in range(tid, n, self.nthreads):
                self.picture(tid, i)

        for f in [self.pool.submit(work, t) for t in range(self.nthreads)]:
            f.result()

    def close(self):
        """Stop the worker threads while CUDA is still alive (their thread-exit hooks release
        per-thread streams and buffers)."""
        if hasattr(self, "drv"):
            self.drv.sb2_e2e_stop()
            del self.drv
        if hasattr(self, "pool"):
            import threading as _t
            gate = _t.Barrier(self.nthreads)

            def release():
                gate.wait()                      # one task per worker thread
                self.lib.schro_b200_thread_release()

            for f in [self.pool.submit(release) for _ in range(self.nthreads)]:
                f.result()
            self.pool.shutdown(wait=True)
            del self.pool


METRIC = "frames/s at 2160p 4:2:0 (wavelet+OBMC+SAD); HBM GB/s as % of B200 peak"


def stage_names(spec):
    if spec.get("full_core", True):
        return ["iwt_inverse", "obmc_render", "upsample", "pyramid", "hier_block_match"]
    return ["iwt_inverse"]


def common_config(args, spec):
    """The `config` object: identical in both arms (ours / --impl reference), functions of the workload only."""
    bpp = 4 if spec["depth_name"] == "s32" else 2
    B = spec["batch"]
    ws = 2 * spec["iwt_w"] * spec["iwt_h"] * 1.5 * bpp * B
    if spec.get("full_core", True):
        ref_bytes = 4 * (spec["width"] + 64) * (spec["height"] + 64) * 1.5
        ws += 3 * ref_bytes * B
    return {"workload": args.workload, "what": spec["label"], "content": args.content, "batch_per_gpu": B,
            "stages": stage_names(spec),
            "pictures": "every picture of a batch differs (own coefficients, noise, motion field, pan)",
            "l2": f"inputs larger than L2: about {ws / 1e6:.0f} MB touched per step on the GPU arm (L2 is 126 MB)",
            "parallelism": "picture-parallel, one process per GPU, no collective"}


def time_device_resident(torch, st, steps, warmup, barrier):
    """W warm-up steps, then K steps between CUDA events on the launching stream; returns milliseconds."""
    for _ in range(warmup):
        st.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st.step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


# Measured on the pool's B200 by tools/vabsdiff_probe.cu (profiles/r02g_vabsdiff_probe.txt): the byte-SAD
# instruction VABSDIFF4.U8.ACC issues at 64 lanes / clock / SM = 18.58 T lane-operations/s at 1965 MHz --
# the roofline of the dependency-free full search (its DRAM traffic is negligible: 1 GB/s).
VABSDIFF4_PEAK_TOPS = 18.58


def next_rows(torch, dev, world, all_ranks, barrier):
    """Device-resident timings of SURVEY.md 8f's rows 3 and 4 (sub-pel refinement, rough / bigblock
    search), which sit beside the picture core rather than in its step: the encoder runs either the
    hierarchical block matcher of the main step or the rough search (schromotionest.c:66-90)."""
    out = {}

    class OneStage:
        def __init__(self, fn):
            self.step = fn

    def timed(fn, steps=10):
        return all_ranks(time_device_resident(torch, OneStage(fn), steps, 3, barrier), "max") / steps

    # ---- rough search: the full search on its own (every 8x8 block of 1080p pictures, +-12) and the chain
    w, h, count = 1920, 1080, 32
    rng = np.random.default_rng(99)
    nbx, nby = block_counts(w, h)
    ps, pr = dev.Pyramid(w, h, count, HBM_LEVELS), dev.Pyramid(w, h, count, HBM_LEVELS)
    orig = ps.slabs[0]
    up = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, h, 32, True), count)
    for p in range(count):
        ref = textured_frame(w, h, rng)
        src = textured_frame(w, h, rng, picture_pan(p))
        for c in range(3):
            ps.slabs[0].upload(p, c, src[c])
            pr.slabs[0].upload(p, c, ref[c])
            up.upload(p, c, ref[c])
    ps.build()
    pr.build()
    dev.edgeextend_upsample(up)
    prm = dev.HbmParams(BLOCK["xbsep"], BLOCK["ybsep"], nbx, nby, 0, 0, 1, 1)
    fld = torch.empty(count * nbx * nby * 20, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: dev.rough_scan_nohint(prm, ps.slabs[0], pr.slabs[0], 0, 12, fld))
    lane_ops = (w // 8) * (h // 8) * count * 625 * 16          # 8x8 block x 625 positions = 16 four-byte SADs each
    tops = lane_ops / (ms * 1e-3) / 1e12
    out["rough_full_search_1080p"] = {
        "what": "schro_rough_me_heirarchical_scan_nohint on level 0: every 8x8 block of 1080p pictures, +-12 (625 positions)",
        "batch_per_gpu": count, "value": round(count * world / (ms * 1e-3), 1), "unit": "frames/s",
        "launch_ms": round(ms, 4),
        "roofline": {"bound": "alu (VABSDIFF4.U8.ACC)", "achieved": round(tops, 2), "peak": VABSDIFF4_PEAK_TOPS,
                     "unit": "T lane-ops/s", "frac": round(tops / VABSDIFF4_PEAK_TOPS, 4),
                     "peak_source": "measured (tools/vabsdiff_probe.cu, profiles/r02g_vabsdiff_probe.txt)"}}
    fields = dev.rough_scan(prm, ps, pr)
    ms = timed(lambda: dev.rough_scan(prm, ps, pr, fields=fields))
    out["rough_me_1080p"] = {
        "what": "schro_rough_me_heirarchical_scan: full search +-12 at level 4, hint levels 3..1 (+-4)",
        "batch_per_gpu": count, "value": round(count * world / (ms * 1e-3), 1), "unit": "frames/s", "ms_per_step": round(ms, 4)}
    # ---- sub-pel refinement of the level-0 field of the block matcher, one reference, quarter-pel
    hf = dev.hbm_scan(prm, ps, pr, 3)
    f0 = hf[0].clone()
    work = f0.clone()

    def subpel():
        work.copy_(f0)
        dev.subpel_refine(orig, up, work, BLOCK["xbsep"], BLOCK["ybsep"], nbx, nby, 2, 0, 0.1)
    ms = timed(subpel)
    out["subpel_refine_1080p"] = {
        "what": "schro_encoder_motion_predict_subpel_deep, one reference, mv_precision 2 (two passes of 8 probes per block)",
        "batch_per_gpu": count, "value": round(count * world / (ms * 1e-3), 1), "unit": "frames/s", "ms_per_step": round(ms, 4)}
    # ---- split-2 pass of the mode decision on the refined field (one reference here: the row's pictures have one)
    ms = timed(lambda: dev.split2_decide(orig, [up], [work], BLOCK["xbsep"], BLOCK["ybsep"], nbx, nby, 2, 0.1))
    out["mode_decision_split2_1080p"] = {
        "what": "schro_do_split2 for every superblock (chroma SADs, DC candidates, entropy + lambda * error decisions), one reference, mv_precision 2",
        "batch_per_gpu": count, "value": round(count * world / (ms * 1e-3), 1), "unit": "frames/s", "ms_per_step": round(ms, 4)}
    return out


def _lowdelay_slice(rng, nbytes, n_luma, n_chroma2, base):
    """One well-formed low-delay slice (schroedinger/schrolowdelay.c:101-178): 7-bit base index, luma length,
    luma then interleaved chroma coefficients as interleaved exp-Golomb codes, zero-run tail trimmed by the
    1-bit guard; values drawn like a quantised residual (mostly 0 / +-1)."""
    def code(v):
        a = abs(int(v)) + 1
        n = a.bit_length()
        bits = []
        for i in range(n - 1):
            bits += [0, (a >> (n - 2 - i)) & 1]
        bits.append(1)
        if v:
            bits.append(1 if v < 0 else 0)
        return bits
    lb = (8 * nbytes).bit_length()
    room = 8 * nbytes - 7 - lb
    vals = (rng.geometric(0.55, size=n_luma + n_chroma2) - 1) * rng.choice([-1, 1], size=n_luma + n_chroma2)
    ybits, uvbits = [], []
    for v in vals[:n_luma]:
        ybits += code(v)
    for v in vals[n_luma:]:
        uvbits += code(v)
    ylen = min(len(ybits), (2 * room) // 3)
    bits = [(base >> (6 - i)) & 1 for i in range(7)] + [(ylen >> (lb - 1 - i)) & 1 for i in range(lb)]
    bits += ybits[:ylen] + uvbits[:room - ylen]
    bits += [1] * (8 * nbytes - len(bits))
    return np.packbits(np.array(bits[:8 * nbytes], np.uint8))


def lowdelay_rows(torch, dev, lib, world, all_ranks, barrier, e2e_threads=12):
    """BASELINE configs[1] as a decoder sees it: VC-2 low-delay intra 1080p pictures arrive as compressed
    slices (60 x 34 slices of 190 bytes = 388 KB per picture), are decoded + dequantised + DC-predicted on
    the device, inverse-transformed (DD 9/7, 4 levels, s16) and converted to 8 bits.  Device-resident for a
    batch, and end to end through the drop-in C API (pinned host slices up, pinned u8 pictures down)."""
    from schroedinger_b200 import compat
    w, h, depth, nh, nv, nbytes, count = 1920, 1088, 4, 60, 34, 190, 64
    rng = np.random.default_rng(4242)
    templates = [_lowdelay_slice(rng, nbytes, 32 * 32, 2 * 16 * 16, int(rng.integers(8, 28))) for _ in range(16)]
    pic_bytes = nh * nv * nbytes
    pitch = (pic_bytes + 255) // 256 * 256
    host = np.zeros((count, pitch), np.uint8)
    for p in range(count):
        pick = rng.integers(0, len(templates), size=nh * nv)
        host[p, :pic_bytes] = np.concatenate([templates[k] for k in pick])
    slices = torch.from_numpy(host.reshape(-1)).cuda()
    qm = [0, 2, 2, 4, 2, 2, 4, 4, 4, 6, 6, 6, 8]
    tq, to = [], []
    for q in range(61):                                       # the Dirac specification's quantiser tables
        base = 1 << (q // 4)
        f = [4 * base, (503829 * base + 52958) // 105917, (665857 * base + 58854) // 117708, (440253 * base + 32722) // 65444][q & 3]
        tq.append(f)
        to.append(1 if q == 0 else 2 if q == 1 else (f + 1) // 2)
    coeffs = dev.PictureSlab(dev.FrameLayout.yuv420("s16", w, h), count)
    out8 = dev.PictureSlab(dev.FrameLayout.yuv420("u8", w, 1080), count)

    class OneStage:
        def __init__(self, fn):
            self.step = fn

    def step():
        dev.lowdelay_decode(slices, pic_bytes, coeffs, depth, nh, nv, nbytes, 1, qm, tq, to, picture_pitch=pitch)
        dev.iwt_inverse_convert(coeffs, out8, 0, depth)
    lib.sb2_profile_reset()
    ms = all_ranks(time_device_resident(torch, OneStage(step), 10, 3, barrier), "max") / 10
    lib.sb2_profile_enable(1)
    step()
    torch.cuda.synchronize()
    lib.sb2_profile_enable(0)
    kern = {k: round(v["ms"], 4) for k, v in collect_profile(lib).items()}
    lib.sb2_profile_reset()
    res = {"what": "VC-2 low-delay intra 1080p decode (BASELINE configs[1]): 2040 slices of 190 bytes per picture -> "
                   "slice decode + dequantise + DC prediction -> inverse DD 9/7 4-level s16 with the conversion to the 8-bit "
                   "picture fused into its last level",
           "batch_per_gpu": count, "value": round(count * world / (ms * 1e-3), 1), "unit": "frames/s",
           "ms_per_step": round(ms, 4), "kernel_ms": kern, "compressed_bytes_per_picture": pic_bytes}
    del coeffs, out8
    # ---- end to end through the drop-in API: one picture per call chain, a pool of host threads
    params = compat.make_params(w, 1080, wavelet_filter_index=0, transform_depth=depth, iwt_luma_width=w, iwt_luma_height=h)
    params.is_lowdelay = 1
    params.n_horiz_slices, params.n_vert_slices = nh, nv
    params.slice_bytes_num, params.slice_bytes_denom = nbytes, 1
    for i, v in enumerate(qm):
        params.quant_matrix[i] = v
    A = compat.frame_new_and_alloc
    pinned, cuda = compat.pinned_domain(), compat.cuda_domain()
    # the slices of each picture in page-locked memory (a u8 frame's luma plane serves as the buffer)
    bufs = []
    for p in range(count):
        f = A(pinned, compat.FORMAT_U8_444, pitch // 16, 16, 0, 0)
        ctypes.memmove(f.contents.components[0].data, host[p].ctypes.data, pic_bytes)
        bufs.append(f)
    outs = [A(pinned, compat.FORMAT_U8_420, w, 1080, 0, 0) for _ in range(count)]
    th = [dict(coef=A(cuda, compat.FORMAT_S16_420, w, h, 0, 0), u8=A(cuda, compat.FORMAT_U8_420, w, 1080, 0, 0))
          for _ in range(e2e_threads)]

    path = os.path.join(ROOT, "bench_native", "libsb2_e2e_driver.so")
    drv = ctypes.CDLL(path)

    class Job(ctypes.Structure):
        _fields_ = [("nthreads", ctypes.c_int), ("npictures", ctypes.c_int), ("slice_bytes", ctypes.c_int),
                    ("params", ctypes.POINTER(compat.SchroParams)), ("slices", ctypes.POINTER(ctypes.c_void_p)),
                    ("out_host", ctypes.POINTER(compat.FrameP)), ("coef_dev", ctypes.POINTER(compat.FrameP)),
                    ("u8_dev", ctypes.POINTER(compat.FrameP))]
    job = Job(e2e_threads, count, pic_bytes, ctypes.pointer(params),
              (ctypes.c_void_p * count)(*[f.contents.components[0].data for f in bufs]),
              (compat.FrameP * count)(*outs), (compat.FrameP * e2e_threads)(*[t["coef"] for t in th]),
              (compat.FrameP * e2e_threads)(*[t["u8"] for t in th]))
    drv.sb2_e2e_lowdelay_run.restype = ctypes.c_double
    drv.sb2_e2e_lowdelay_run(ctypes.byref(job), 1)
    barrier()
    steps = 40                                # 2 560 pictures: the workers' first-call buffer allocations amortise
    tm = (ctypes.c_double * 4)()
    # three timed runs each, the median reported: the per-picture leg occasionally runs several times slower (seen:
    # 900 against 8 900 frames/s with the same threads; blocking waits of a dozen workers on a 16-core host), all
    # three values are kept in the line
    runs = []
    for _ in range(3):
        drv.sb2_e2e_lowdelay_times(tm)
        runs.append(all_ranks(drv.sb2_e2e_lowdelay_run(ctypes.byref(job), steps), "max"))
    wall = sorted(runs)[1]
    drv.sb2_e2e_lowdelay_times(tm)
    res["e2e_call_ms"] = {k: round(tm[i] / (steps * count) * 1e3, 3) for i, k in enumerate(
        ("decode_lowdelay", "(unused)", "inverse_iwt_combine", "gpuframe_to_cpu"))}
    res["e2e"] = {"value": round(count * world * steps / wall, 1), "unit": "frames/s",
                  "runs": [round(count * world * steps / r, 1) for r in runs],
                  "h2d_bytes_per_step": pic_bytes * count, "d2h_bytes_per_step": 1920 * 1080 * 3 // 2 * count,
                  "api": f"schro_b200_decode_lowdelay_transform_data, schro_b200_frame_inverse_iwt_combine, "
                         f"schro_gpuframe_to_cpu; {e2e_threads} host threads (pthreads, bench_native/e2e_driver.c), pinned slices and pictures"}
    # ... and through the batched drop-in: one launch per stage and one wait per batch of pictures
    drv.sb2_e2e_lowdelay_run_batched.restype = ctypes.c_double
    bt, bb = int(os.environ.get("SB2_LD_BATCH_THREADS", "4")), int(os.environ.get("SB2_LD_BATCH", "8"))
    bjob = Job(bt, count, pic_bytes, ctypes.pointer(params), job.slices, job.out_host, job.coef_dev, job.u8_dev)
    drv.sb2_e2e_lowdelay_run_batched(ctypes.byref(bjob), 1, bb)
    barrier()
    runs = [all_ranks(drv.sb2_e2e_lowdelay_run_batched(ctypes.byref(bjob), steps, bb), "max") for _ in range(3)]
    res["e2e_batched"] = {"value": round(count * world * steps / sorted(runs)[1], 1), "unit": "frames/s",
                          "runs": [round(count * world * steps / r, 1) for r in runs],
                          "api": f"schro_b200_decode_lowdelay_pictures, {bb} pictures per call, {bt} host threads (pthreads)"}
    for f in bufs + outs:
        lib.schro_frame_unref(f)
    for t in th:
        lib.schro_frame_unref(t["coef"])
        lib.schro_frame_unref(t["u8"])
    return {"lowdelay_1080p": res}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from schroedinger_b200 import device as dev
    from schroedinger_b200 import lib
    spec = workload_spec(args.workload)
    if args.batch:
        spec["batch"] = args.batch
    spec["overlap"] = args.overlap
    st = Stages(spec, torch, dev)
    B = spec["batch"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from schroedinger_b200 import sharding

    def all_ranks(value, op):
        """max / min of a float over the ranks (device-side all-reduce; identity for one rank)"""
        if op == "max":
            return sharding.max_over_ranks(value, device="cuda")
        return -sharding.max_over_ranks(-value, device="cuda")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # per-launch CUDA events (pooled: the warm-up creates them, the timed loop only records them)
    lib.sb2_profile_enable(1)
    for _ in range(max(3, args.warmup)):
        st.step()
    barrier()
    lib.sb2_profile_reset()
    launches0 = lib.sb2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        st.step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    lib.sb2_profile_enable(0)
    # stop polling NVML before the API-heavy e2e arm: nvidia-smi at 10 Hz stalls the host-side
    # CUDA calls it contends with (measured: e2e 29 fps with the sampler, 157 without)
    clocks = sampler.stop((t_begin, t_end)) if rank == 0 else None
    launches = lib.sb2_launch_count() - launches0
    prof = collect_profile(lib)
    lib.sb2_profile_reset()
    elapsed_ms = all_ranks(e0.elapsed_time(e1), "max")

    # ---- e2e through the drop-in C API with pinned host frames ----
    def e2e_leg(widen):
        """One end-to-end measurement.  Every collective sits outside the per-rank try blocks: the ranks
        first agree that all of them are set up, then run chunks of K steps until 3 s have been timed,
        deciding together when to stop (max time over ranks / common step count)."""
        hf, err = None, None
        try:
            nthreads = min(args.e2e_threads, B)
            hf = HostFrames(spec, lib, nthreads, widen=widen)
            if args.e2e_driver == "native":
                hf.native_start()
            for _ in range(3):
                hf.step()
        except Exception as ex:                      # noqa: BLE001 - reported in the JSON line
            err = repr(ex)
        if all_ranks(0.0 if err else 1.0, "min") < 0.5:
            if hf is not None:
                hf.close()
            return {"value": None, "unit": "frames/s", "error": err or "another rank failed to set up"}
        barrier()
        e2e_steps, t0 = 0, time.perf_counter()
        while True:
            if args.e2e_driver == "native":
                hf.drv.sb2_e2e_run(args.steps)       # the worker threads run the steps back to back
            else:
                for _ in range(args.steps):
                    hf.step()
            e2e_steps += args.steps
            go_on = time.perf_counter() - t0 < 3.0 and e2e_steps < 50 * args.steps
            if all_ranks(1.0 if go_on else 0.0, "min") < 0.5:
                break
        torch.cuda.synchronize()
        dt = all_ranks(time.perf_counter() - t0, "max")
        res = {"value": B * e2e_steps * world / dt, "unit": "frames/s", "steps": e2e_steps,
               "h2d_bytes_per_step": hf.h2d * B, "d2h_bytes_per_step": hf.d2h * B,
               "coefficients": ("quantised s16 up, dequantised to s32 on the device" if hf.widen
                                else ("dequantised s32 up" if spec["depth_name"] == "s32" else "s16 up")),
               "api": ("drop-in schro_* C API (schro_frame_to_gpu, "
                       + ("schro_b200_frame_dequantise_widen, " if hf.widen else "")
                       + "schro_frame_inverse_iwt_transform, schro_motion_render, schro_frame_mc_edgeextend, "
                       "schro_upsampled_frame_upsample, schro_frame_downsample, schro_hbm_scan, "
                       "schro_hierarchical_bm_scan_hint, schro_hbm_motion_field, schro_gpuframe_to_cpu)" if hf.full else
                       "drop-in schro_frame_inverse_iwt_transform on the host frame (staged H2D, transform, D2H)")
                      + f" on pinned host SchroFrames, {min(args.e2e_threads, B)} host threads/GPU"
                        f" ({args.e2e_driver} driver), one stream each"}
        hf.close()
        return res

    if args.no_e2e:
        e2e = {"value": None, "unit": "frames/s", "error": "skipped (--no-e2e: profiler runs)"}
        e2e_s32 = None
    else:
        e2e = e2e_leg(True)
        # the round-1 data path for comparison: dequantised s32 coefficients travel (twice the bytes)
        e2e_s32 = e2e_leg(False) if (spec.get("full_core", True) and spec["depth_name"] == "s32") else None

    # ---- the other BASELINE.json configurations that fit one GPU, device-resident, same script ----
    other = {}
    if args.workload == "picture_core_2160p" and not args.no_other:
        del st                                        # its slabs go back to the allocator
        torch.cuda.empty_cache()
        for name in ("wavelet_1080p_dd97", "wavelet_1080p_legall", "picture_core_1080p"):
            ospec = workload_spec(name)
            ospec["overlap"] = False
            ost = Stages(ospec, torch, dev)
            nst = 20 if not ospec.get("full_core", True) else 10
            ms = all_ranks(time_device_resident(torch, ost, nst, 3, barrier), "max")
            alg = sum(s["alg_bytes"] for s in ost.stages)
            other[name] = {"what": ospec["label"], "batch_per_gpu": ospec["batch"],
                           "value": round(ospec["batch"] * world * nst / (ms * 1e-3), 1), "unit": "frames/s",
                           "alg_GBps": round(alg * nst / (ms * 1e-3) / 1e9, 1)}
            if ospec.get("full_core", True):
                # per-kernel times of one more step (CUDA events per launch), as for the headline workload
                lib.sb2_profile_reset()
                lib.sb2_profile_enable(1)
                ost.step()
                torch.cuda.synchronize()
                lib.sb2_profile_enable(0)
                other[name]["kernel_ms"] = {k: round(v["ms"], 4) for k, v in sorted(collect_profile(lib).items())}
                lib.sb2_profile_reset()
            del ost
            torch.cuda.empty_cache()
        torch.cuda.empty_cache()
        other.update(next_rows(torch, dev, world, all_ranks, barrier))
        torch.cuda.empty_cache()
        other.update(lowdelay_rows(torch, dev, lib, world, all_ranks, barrier))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ms_per_step = elapsed_ms / args.steps
    value = sharding.aggregate_throughput(B * args.steps, elapsed_ms * 1e-3, world)
    # dominant kernel = largest share of device time in the timed region
    dom_tag, dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    # DRAM bytes per launch come from the committed ncu capture of the same command (valid for the
    # workload / batch it was taken at); the capture is named beside the number
    traffic, traffic_src = {}, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj["workload"] == args.workload and tj["batch_per_gpu"] == B and not args.overlap:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    total_prof_ms = max(1e-9, sum(r["ms"] for r in prof.values()))
    roofline = None
    if dom:
        ach = dom["bytes"] / dom["launches"] / (dom["ms"] / dom["launches"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom_tag, "achieved": round(ach, 1), "peak": peak,
                    "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic.get(dom_tag),
                    "traffic_source": traffic_src if traffic.get(dom_tag) is not None else None,
                    "alg_bytes_per_launch": round(dom["bytes"] / dom["launches"]),
                    "peak_source": peak_src, "launch_ms": round(dom["ms"] / dom["launches"], 4),
                    "share_of_step": round(dom["ms"] / total_prof_ms, 3)}
    roofline_all = {}
    for k, v in sorted(prof.items()):
        a_ = v["bytes"] / max(v["ms"], 1e-9) / 1e6
        roofline_all[k] = {"achieved_GBps": round(a_, 1), "frac": round(a_ / peak, 4),
                           "launch_ms": round(v["ms"] / v["launches"], 4), "traffic": traffic.get(k),
                           "share_of_step": round(v["ms"] / total_prof_ms, 3)}
    kern = {k: {"ms_per_step": round(v["ms"] / args.steps, 4), "launches_per_step": v["launches"] // args.steps,
                "alg_GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in sorted(prof.items())}
    alg_step = sum(v["bytes"] for v in prof.values()) / args.steps
    out = {
        "metric": METRIC,
        "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "s32" if spec["depth_name"] == "s32" else "s16",
        "data": "synthetic",
        "config": common_config(args, spec),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_all": roofline_all,
        "step_alg_GBps": round(alg_step / (ms_per_step * 1e-3) / 1e9, 1),
        "step_frac_of_peak": round(alg_step / (ms_per_step * 1e-3) / 1e9 / peak, 4),
        "kernels": kern,
        "clocks": clocks,
    }
    if e2e_s32 is not None:
        out["e2e_s32"] = e2e_s32
    if other:
        out["other_configs"] = other
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(spec, seconds=args.cpu_seconds)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on host cores
# --------------------------------------------------------------------------------------
def load_cpu_lib():
    ref = os.path.join(ROOT, "oracle", "_ref", "libschro_ref.so")
    if os.path.exists(ref):
        return ctypes.CDLL(ref, mode=ctypes.RTLD_LOCAL), "reference", "ref"
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(port):
        subprocess.check_call(["make", "-C", ROOT, "oracle/liboracle.so"])
    return ctypes.CDLL(port, mode=ctypes.RTLD_LOCAL), "port", "oracle"


class CpuWorkload:
    """The same step on the host: the reference's own C (oracle/_ref, kind "reference") or,
    where that is not available, the oracle port.  One picture = inverse wavelet + OBMC render
    + edge-extend/upsample of the decoded picture + pyramids + hierarchical block matching."""

    def __init__(self, spec, nthreads):
        from tests import helpers as H
        self.H, self.spec, self.nthreads = H, spec, nthreads
        self.lib, self.kind, self.prefix = load_cpu_lib()
        self.full = spec.get("full_core", True)
        rng = np.random.default_rng(11)
        self.base = make_coeff_frame(spec, rng)
        self.coef = [[p.copy() for p in self.base] for _ in range(nthreads)]
        if not self.full:
            return
        W, Hh = spec["width"], spec["height"]
        self.W, self.Hh = W, Hh
        self.nbx, self.nby = block_counts(W, Hh)
        self.mvs = make_mv_field(self.nbx, self.nby, rng)
        self.refs = []
        for r in range(2):
            fr = textured_frame(W, Hh, rng, pan=(5 * r, 3 * r))
            planes = []
            for c in range(3):
                pl = H.HostPlane(fr[c].shape[1], fr[c].shape[0], ext=32, upsampled=True)
                pl.set_image(fr[c])
                H.cpu_edgeextend(self.lib, self.prefix, pl)
                H.cpu_upsample(self.lib, self.prefix, pl)
                planes.append(pl)
            self.refs.append(planes)
        self.src = textured_frame(W, Hh, rng, pan=(5, 3))
        self.refpic = textured_frame(W, Hh, rng)
        self.newref = [[H.HostPlane(p.w, p.h, ext=32, upsampled=True) for p in self.refs[0]]
                       for _ in range(nthreads)]
        self.acc = [[np.zeros((p.h, p.w), np.int16) for p in self.refs[0]] for _ in range(nthreads)]
        self.hbm_scratch = [{} for _ in range(nthreads)]      # pyramid / field buffers, allocated once
        # the reference initialises static tables lazily and not thread-safely
        # (schromotion8.c:13-18): touch them once before the threads start
        self.frame(0)

    def frame(self, t):
        """All stages for one picture on thread slot t."""
        lib, prefix, spec, H = self.lib, self.prefix, self.spec, self.H
        is32 = 1 if spec["depth_name"] == "s32" else 0
        fn = getattr(lib, f"{prefix}_iwt_inv")
        fn.restype = None
        for p in self.coef[t]:
            fn(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(p.strides[0]), p.shape[1], p.shape[0],
               is32, spec["filter"], spec["transform_depth"])
        if not self.full:
            return
        P, I = ctypes.c_void_p * 3, ctypes.c_int * 3
        resid, outp, sizes = self.coef[t], self.newref[t], [(p.w, p.h) for p in self.refs[0]]
        if self.kind == "reference":
            mp = H.RefMotionParams(self.W, self.Hh, 2, BLOCK["xbsep"], BLOCK["ybsep"], BLOCK["xblen"],
                                   BLOCK["yblen"], self.nbx, self.nby, BLOCK["prec"], 1, 1, 1, 2)
            f = lib.ref_motion_render
            f.restype = None
            f(ctypes.byref(mp), self.mvs.ctypes.data_as(ctypes.c_void_p),
              P(*[p.ptr for p in self.refs[0]]), I(*[p.stride for p in self.refs[0]]),
              P(*[p.ptr for p in self.refs[1]]), I(*[p.stride for p in self.refs[1]]),
              P(*[a.ctypes.data for a in self.acc[t]]), I(*[a.strides[0] for a in self.acc[t]]),
              P(*[a.ctypes.data for a in resid]), I(*[a.strides[0] for a in resid]), is32, 1,
              P(*[p.ptr for p in outp]), I(*[p.stride for p in outp]), 0)
        else:
            f = lib.oracle_obmc_render
            f.restype = None
            for k, (w, h) in enumerate(sizes):
                hs = 1 if k else 0
                prm = H.OracleObmcParams(BLOCK["xbsep"] >> hs, BLOCK["ybsep"] >> hs, BLOCK["xblen"] >> hs,
                                         BLOCK["yblen"] >> hs, self.nbx, self.nby, BLOCK["prec"], 1, 1, 1,
                                         hs, hs, k)
                f(ctypes.byref(prm), self.mvs.ctypes.data_as(ctypes.c_void_p), self.refs[0][k].ptr,
                  self.refs[1][k].ptr, self.refs[0][k].stride, w, h,
                  self.acc[t][k].ctypes.data_as(ctypes.c_void_p), self.acc[t][k].strides[0],
                  resid[k].ctypes.data_as(ctypes.c_void_p), resid[k].strides[0], is32, 1, outp[k].ptr,
                  outp[k].stride)
        for pl in outp:
            H.cpu_edgeextend(lib, prefix, pl)
            H.cpu_upsample(lib, prefix, pl)
        if self.kind == "reference":
            H.ref_hbm(lib, self.src, self.refpic, self.W, self.Hh, BLOCK["xbsep"], BLOCK["ybsep"],
                      HBM_LEVELS, 0, 0, 3, scratch=self.hbm_scratch[t])
        else:
            H.oracle_hbm(lib, self.src, self.refpic, self.W, self.Hh, BLOCK["xbsep"], BLOCK["ybsep"],
                         HBM_LEVELS, 0, 0, 3)

    def one_pass(self):
        """nthreads pictures, one per host thread (the reference's picture-parallel model)."""
        ts = [threading.Thread(target=self.frame, args=(t,)) for t in range(self.nthreads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0


def cpu_baseline(spec, seconds=10.0, steps=None, warmup=1):
    cores = os.cpu_count() or 1
    nthreads = cores
    wl = CpuWorkload(spec, nthreads)
    for _ in range(warmup):
        wl.one_pass()
    total, n = 0.0, 0
    while (steps is None and total < seconds and n < 50) or (steps is not None and n < steps):
        total += wl.one_pass()
        n += 1
    fps = nthreads * n / total
    what = ("unmodified reference C (gcc -O3 -DDISABLE_ORC: the plain-C Orc kernels, not liborc's SIMD JIT; "
            "the five runtime-JIT OBMC kernels run through a C interpreter shim)"
            if wl.kind == "reference" else "oracle port")
    return {"value": round(fps, 3), "unit": "frames/s", "cores": nthreads, "kind": wl.kind,
            "sample": f"{n} passes x {nthreads} pictures of the same workload "
                      f"({'all stages' if wl.full else 'wavelet only'}), picture-parallel on {nthreads} host threads; {what}",
            "seconds": round(total, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = workload_spec(args.workload)
    if args.batch:
        spec["batch"] = args.batch
    t0 = time.perf_counter()
    cb = cpu_baseline(spec, steps=max(1, args.steps), warmup=1 if args.warmup > 0 else 0)
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(cb["seconds"] / max(1, args.steps) * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "s32" if spec["depth_name"] == "s32" else "s16",
        "data": "synthetic",
        "config": common_config(args, spec),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="picture_core_2160p")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--content", default="natural", choices=["natural", "periodic"],
                    help="texture of the synthetic pictures (see textured_frame)")
    ap.add_argument("--e2e-threads", type=int, default=32, help="host threads driving the drop-in API in the e2e leg")
    ap.add_argument("--e2e-driver", default="native", choices=["native", "python"],
                    help="host threads of the e2e leg: pthreads in bench_native/e2e_driver.c, or Python threads + ctypes")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (for ncu runs of the timed region)")
    ap.add_argument("--no-other", action="store_true", help="skip the device-resident 1080p configurations")
    ap.add_argument("--overlap", action="store_true",
                    help="run the motion-estimation stages on a second stream (measured: no gain in the "
                         "device-resident loop, whose batched wavefronts already keep every SM occupied)")
    args = ap.parse_args()
    global CONTENT
    CONTENT = args.content
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
