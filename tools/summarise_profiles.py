#!/usr/bin/env python
"""Turns the files tools/collect_profiles.sh leaves in gpurun_out/ into the summaries committed under
profiles/: <R>_summary.txt (kernel shares under ncu next to the live shares, per-kernel counters of the
--set full capture), profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py) and the NS-1 table.
usage: summarise_profiles.py <R>"""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def read_ncu_csv(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def short(name):
    m = re.match(r"(?:void )?(?:[a-z0-9_]+::)*([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


out = []
plain = json.load(open(os.path.join(G, f"{R}_plain.json")))
out.append(f"round 2, capture {R[-1]} (end of round): one B200, bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other")
out.append("the same command exited 0 without ncu first; ncu serialises kernels with cold caches: compare SHARES, not times")
out.append(f"plain run: {plain['value']} frames/s device-resident, {plain['ms_per_step']} ms per step of {plain['config']['batch_per_gpu']} pictures")
out.append("")
# ---- launch list
rows = read_ncu_csv(os.path.join(G, f"{R}_launches.csv"))
agg = {}
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
live = plain.get("kernels", {})
live_tot = sum(v["ms_per_step"] for v in live.values()) or 1.0
out.append(f"launch list ({R}_launches.csv, warm-up + timed steps): share of kernel time under ncu")
out.append(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:70]:70s} {a[0]:8d} {a[1]:9.3f} {100 * a[1] / tot:6.1f}%")
out.append("")
out.append("live shares (CUDA events inside the plain run's timed region, by launch tag):")
for k, v in sorted(live.items(), key=lambda kv: -kv[1]["ms_per_step"]):
    out.append(f"  {k:32s} {v['ms_per_step']:8.3f} ms/step {100 * v['ms_per_step'] / live_tot:6.1f}%  ({v['launches_per_step']} launches, {v['alg_GBps']} GB/s algorithmic)")
out.append("")
# ---- full capture: per-launch counters
full = os.path.join(G, f"{R}_full_raw.csv")
traffic = {}
if os.path.exists(full):
    rows = list(csv.reader(open(full)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    units = rows[1]
    out.append(f"per-launch counters of one timed step (ncu --set full --clock-control none, {R}_full_raw.csv not committed: 30 MB)")
    out.append("kernel | grid | " + " | ".join(w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")
                                                .replace(".avg.pct_of_peak_sustained_active", "%").replace(".sum", "") for w in want))

    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    seq = []
    for r in data:
        name = short(r[col["Kernel Name"]])
        vals = []
        for w in want:
            i = col.get(w)
            vals.append((r[i] + " " + units[i]).strip() if i is not None else "-")
        out.append(f"{name[:60]} | {r[col['Grid Size']]} | " + " | ".join(vals))
        i_r, i_w = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
        if i_r is not None:
            seq.append((name, tobytes(r[i_r], units[i_r]) + tobytes(r[i_w], units[i_w])))
    # DRAM bytes per launch by bench.py's launch tags, in launch order of a step
    wav = {}
    grids = {}
    for r in data:
        grids.setdefault(short(r[col["Kernel Name"]]), []).append(int(r[col["Grid Size"]].strip("()").split(",")[0]))
    for (name, b), r in zip(seq, data):
        if name.startswith("hbm_wave_kernel"):
            g = int(r[col["Grid Size"]].strip("()").split(",")[0])
            tag = {"hbm_wave_kernel<32, 8>": "hbm_level_s4_r20", "hbm_wave_kernel<32, 2>": "hbm_level_s3_r10",
                   "hbm_wave_kernel<16, 2>": "hbm_level_s2_r5"}.get(name)
            if tag is None:         # <8, 1> serves levels 1 and 0: the larger grid is level 0
                tag = "hbm_level_s0_r3" if g == max(grids[name]) else "hbm_level_s1_r3"
            traffic[tag] = b
        elif name.startswith("obmc_kernel_v4"):
            traffic["obmc_render_add"] = b
        elif name.startswith("upsample_kernel_words"):
            traffic["upsample"] = b
        elif name.startswith("wavelet_inv_fast_kernel"):
            wav.setdefault("list", []).append(b)
    lst = sorted(wav.get("list", []), reverse=True)
    if len(lst) >= 2:
        traffic["wavelet_inv_s32_f6_w3840"], traffic["wavelet_inv_s32_f6_w1920"] = lst[0], lst[1]
    json.dump({"source": f"profiles/{R}_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
               "workload": "picture_core_2160p", "batch_per_gpu": plain["config"]["batch_per_gpu"],
               "dram_bytes_per_launch": traffic}, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    out.append("")
# ---- NS-1
dram = os.path.join(G, f"{R}_wavelet_fused_dram.csv")
if os.path.exists(dram):
    rows = read_ncu_csv(dram)
    per = {}
    for r in rows:
        k = (r["ID"], short(r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if "bytes" in r["Metric Name"]:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        else:
            v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        per.setdefault(k, {})[r["Metric Name"]] = v
    agg = {}
    for (_, name), m in per.items():
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0)
        a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    out.append("NS-1, levels 1 + 0 of the inverse Daubechies 9/7 s32 transform, 32 pictures 2160p 4:2:0, depth 2 (ncu, per launch):")
    for name, a in sorted(agg.items()):
        out.append(f"  {name[:70]:70s} launches {a[0]:3d}  {a[1] / a[0]:7.3f} ms  DRAM {a[2] / a[0] / 1e9:6.3f} GB per launch")
    t = os.path.join(G, f"{R}_wavelet_fused_times.txt")
    if os.path.exists(t):
        out.append("  timed without a profiler (tools/time_wavelet_fused.py):")
        out += ["    " + l.rstrip() for l in open(t)]
open(os.path.join(P, f"{R}_summary.txt"), "w").write("\n".join(out) + "\n")
for f in (f"{R}_launches.csv", f"{R}_plain.json"):
    if os.path.exists(os.path.join(G, f)):
        open(os.path.join(P, f), "w").write(open(os.path.join(G, f)).read())
print("\n".join(out[:60]))
