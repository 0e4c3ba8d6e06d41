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
                    transform_depth=5, batch=8,
                    label="2160p 4:2:0 10-bit: inverse Daubechies 9/7 5-level s32 wavelet"
                          " + half-pel upsample + 1/4-pel OBMC render (2 refs, 12x12/8x8)"
                          " + 4-level hierarchical SAD block matching")
    if name == "wavelet_1080p_dd97":
        return dict(width=1920, height=1080, iwt_w=1920, iwt_h=1088, depth_name="s16", filter=0,
                    transform_depth=4, batch=64,
                    label="1080p 4:2:0 8-bit: inverse Deslauriers-Dubuc 9/7 4-level s16 (config 2)")
    if name == "wavelet_1080p_legall":
        return dict(width=1920, height=1080, iwt_w=1920, iwt_h=1088, depth_name="s16", filter=1,
                    transform_depth=4, batch=64,
                    label="1080p 4:2:0 8-bit: forward+inverse LeGall 5/3 4-level s16 (config 1)")
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
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
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


class Stages:
    """Device-resident stages of one step.  Each stage launches on the current stream."""

    def __init__(self, spec, torch, dev):
        self.spec, self.torch, self.dev = spec, torch, dev
        self.stages = []
        B = spec["batch"]
        rng = np.random.default_rng(2026)
        # ---- stage 1: inverse wavelet (BASELINE config 2/3) ----
        lay = dev.FrameLayout.yuv420(spec["depth_name"], spec["iwt_w"], spec["iwt_h"])
        self.coef = dev.PictureSlab(lay, B, zero=False)
        self.recon = dev.PictureSlab(lay, B, zero=False)
        planes = make_coeff_frame(spec, rng)
        for c in range(3):
            self.coef.upload(0, c, planes[c])
        one = self.coef.buf[:lay.pitch]
        for p in range(1, B):
            self.coef.buf[p * lay.pitch:(p + 1) * lay.pitch].copy_(one)
        self.ws = dev.Workspace()
        ncoef = sum(w * h for w, h in lay.comp_sizes)
        self.stages.append(dict(
            name="iwt_inverse", frames=B, alg_bytes=2.0 * ncoef * lay.bpp * B,
            run=lambda: dev.iwt_inverse(self.coef, self.recon, spec["filter"],
                                        spec["transform_depth"], self.ws)))
        if spec.get("also_forward"):
            self.stages.append(dict(
                name="iwt_forward", frames=B, alg_bytes=2.0 * ncoef * lay.bpp * B,
                run=lambda: dev.iwt_forward(self.recon, self.coef, spec["filter"],
                                            spec["transform_depth"], self.ws)))
        extra = getattr(dev, "bench_stages", None)
        if extra and spec.get("full_core", True):
            self.stages.extend(extra(spec, rng))
        self.working_set = self.coef.nbytes + self.recon.nbytes

    def step(self):
        for s in self.stages:
            s["run"]()


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
    """Pinned host frames + the drop-in C API (schro_* symbols) for the e2e measurement."""

    def __init__(self, spec, lib, nthreads):
        from schroedinger_b200 import compat
        self.compat, self.lib, self.spec = compat, lib, spec
        self.nthreads = nthreads
        B = spec["batch"]
        self.domain = compat.pinned_domain()
        fmt = compat.FORMAT_S32_420 if spec["depth_name"] == "s32" else compat.FORMAT_S16_420
        rng = np.random.default_rng(7)
        planes = make_coeff_frame(spec, rng)
        self.frames = []
        for _ in range(B):
            f = compat.frame_new_and_alloc(self.domain, fmt, spec["iwt_w"], spec["iwt_h"])
            for c in range(3):
                compat.frame_plane(f, c)[...] = planes[c]
            self.frames.append(f)
        self.params = compat.make_params(spec["width"], spec["height"], spec["filter"],
                                         spec["transform_depth"], spec["iwt_w"], spec["iwt_h"])
        bpp = 4 if spec["depth_name"] == "s32" else 2
        self.bytes_per_frame = int(spec["iwt_w"] * spec["iwt_h"] * 1.5 * bpp)
        self.extra = getattr(compat, "bench_host_stages", None)

    def step(self):
        """One e2e step: every picture of the batch through the drop-in API, pictures spread
        over host threads (the reference's own threading model, one stream per thread)."""
        lib, frames, params = self.lib, self.frames, self.params

        def work(tid):
            for i in range(tid, len(frames), self.nthreads):
                lib.schro_frame_inverse_iwt_transform(frames[i], ctypes.byref(params))

        ts = [threading.Thread(target=work, args=(t,)) for t in range(self.nthreads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()


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
    st = Stages(spec, torch, dev)
    B = spec["batch"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        st.step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.sb2_profile_reset()
    lib.sb2_profile_enable(1)
    launches0 = lib.sb2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        st.step()
    e1.record()
    barrier()
    lib.sb2_profile_enable(0)
    elapsed_ms = e0.elapsed_time(e1)
    launches = lib.sb2_launch_count() - launches0
    prof = collect_profile(lib)
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    # ---- e2e through the drop-in C API with pinned host frames ----
    e2e = None
    try:
        nthreads = min(4, B)
        hf = HostFrames(spec, lib, nthreads)
        for _ in range(2):
            hf.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            hf.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": B * args.steps * world / dt, "unit": "frames/s",
               "h2d_bytes_per_step": hf.bytes_per_frame * B, "d2h_bytes_per_step": hf.bytes_per_frame * B,
               "api": "schro_frame_inverse_iwt_transform on pinned host SchroFrames, "
                      f"{nthreads} host threads/GPU, one stream each"}
    except Exception as ex:  # keep the device-resident number even if the host arm breaks
        e2e = {"value": None, "unit": "frames/s", "error": repr(ex)}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ms_per_step = elapsed_ms / args.steps
    value = B * world * args.steps / (elapsed_ms * 1e-3)
    # dominant kernel = largest share of device time in the timed region
    dom_tag, dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roofline = None
    if dom:
        ach = dom["bytes"] / dom["launches"] / (dom["ms"] / dom["launches"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom_tag, "achieved": round(ach, 1), "peak": peak,
                    "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                    "peak_source": peak_src, "launch_ms": round(dom["ms"] / dom["launches"], 4),
                    "share_of_step": round(dom["ms"] / max(1e-9, sum(r["ms"] for r in prof.values())), 3)}
    stage_report = {}
    for s in st.stages:
        stage_report[s["name"]] = {"frames": s["frames"], "alg_bytes": s["alg_bytes"]}
    kern = {k: {"ms_per_step": round(v["ms"] / args.steps, 4), "launches_per_step": v["launches"] // args.steps,
                "alg_GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in sorted(prof.items())}
    out = {
        "metric": "frames/s at 2160p 4:2:0 (wavelet+OBMC+SAD); HBM GB/s as % of B200 peak",
        "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "s32" if spec["depth_name"] == "s32" else "s16",
        "data": "synthetic",
        "config": {"workload": args.workload, "what": spec["label"], "batch_per_gpu": B,
                   "stages": [s["name"] for s in st.stages],
                   "l2": f"inputs larger than L2: {st.working_set / 1e6:.0f} MB working set per step",
                   "parallelism": f"picture-parallel x{world}, no collective"},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kern,
        "clocks": clocks,
    }
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


def cpu_pass(lib, prefix, spec, frames, nthreads):
    """All stages of the step on `len(frames)` pictures, picture-parallel over host threads."""
    is32 = 1 if spec["depth_name"] == "s32" else 0
    fn = getattr(lib, f"{prefix}_iwt_inv")
    fn.restype = None

    def work(tid):
        for i in range(tid, len(frames), nthreads):
            for p in frames[i]:
                fn(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(p.strides[0]), p.shape[1],
                   p.shape[0], is32, spec["filter"], spec["transform_depth"])
        extra = globals().get("cpu_extra_stages")
        if extra:
            extra(lib, prefix, spec, tid, nthreads, len(frames))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def cpu_baseline(spec, seconds=10.0, steps=None, warmup=1):
    lib, kind, prefix = load_cpu_lib()
    cores = os.cpu_count() or 1
    nthreads = cores
    rng = np.random.default_rng(11)
    nframes = max(nthreads, 1)
    base = make_coeff_frame(spec, rng)
    frames = [[p.copy() for p in base] for _ in range(nframes)]
    for _ in range(warmup):
        cpu_pass(lib, prefix, spec, frames, nthreads)
    total, n = 0.0, 0
    while (steps is None and total < seconds and n < 50) or (steps is not None and n < steps):
        total += cpu_pass(lib, prefix, spec, frames, nthreads)
        n += 1
    fps = nframes * n / total
    return {"value": round(fps, 3), "unit": "frames/s", "cores": nthreads, "kind": kind,
            "sample": f"{n} passes x {nframes} pictures of the same workload, picture-parallel on "
                      f"{nthreads} host threads ({'unmodified reference C, -O3 -DDISABLE_ORC' if kind == 'reference' else 'oracle port'})",
            "seconds": round(total, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = workload_spec(args.workload)
    t0 = time.perf_counter()
    cb = cpu_baseline(spec, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    out = {
        "impl": "reference",
        "metric": "frames/s at 2160p 4:2:0 (wavelet+OBMC+SAD); HBM GB/s as % of B200 peak",
        "value": cb["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(cb["seconds"] / max(1, args.steps) * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "s32" if spec["depth_name"] == "s32" else "s16",
        "data": "synthetic",
        "config": {"workload": args.workload, "what": spec["label"], "batch_per_gpu": spec["batch"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="picture_core_2160p")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
