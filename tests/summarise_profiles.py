"""Turn the raw evidence of tests/gpu_profile_round.sh (gpurun_out/*) into the committed summaries under profiles/:
    python tests/summarise_profiles.py r01d
  * profiles/<tag>_ncu_full_<kernel>.csv      selected metrics of the captured launches (raw pages exported on the box)
  * profiles/<tag>_stalls.txt                 one line per kernel: utilisations and warp-stall mix
  * profiles/<tag>_launches_8x8x8.csv         the ncu launch list of bench.py's timed region, verbatim
  * profiles/<tag>_launches_8x8x8_summary.txt per-kernel totals and shares of that list
  * profiles/<tag>_bench_8x8x8.json, profiles/<tag>_bench_reference_arm.json
"""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01x"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active"] + [
        f"smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio" for r in
        ("long_scoreboard", "wait", "short_scoreboard", "math_pipe_throttle", "not_selected", "selected", "branch_resolving",
         "lg_throttle", "barrier")]

def num(r, ix, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (ValueError, KeyError):
        return 0.0


def gbs(r, ix, units, k):
    scale = {"Tbyte/s": 1e3, "Gbyte/s": 1.0, "Mbyte/s": 1e-3, "Kbyte/s": 1e-6, "byte/s": 1e-9}
    return num(r, ix, k) * scale.get(units[ix[k]], 0.0) if k in ix else 0.0


stall_lines, seen = [], set()
for f in sorted(os.listdir(OUT)):
    if not (f.startswith(f"ncu_{TAG}_") and f.endswith(".raw.csv")):
        continue
    rows = list(csv.reader(open(os.path.join(OUT, f))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    by_kernel = collections.OrderedDict()
    for r in data:
        by_kernel.setdefault(r[ix["Kernel Name"]].split("(")[0].split("::")[-1].replace("<", "_").replace(">", ""), []).append(r)
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for kernel, launches in by_kernel.items():
        launches = launches[-2:]       # the last two captured launches of each kernel
        with open(os.path.join(PROF, f"{TAG}_ncu_full_{kernel}.csv"), "w", newline="") as fo:
            w = csv.writer(fo)
            w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
            for name in WANT:
                if name in ix:
                    w.writerow([name, units[ix[name]]] + [d[ix[name]] for d in launches])
        r = launches[-1]
        tot = sum(num(r, ix, h) for h in stall) or 1.0
        top = sorted(((num(r, ix, h), h) for h in stall), reverse=True)[:5]
        stall_lines.append(
            f"{kernel:18s} {r[ix['gpu__time_duration.sum']]:>9s} {units[ix['gpu__time_duration.sum']]}  regs {r[ix['launch__registers_per_thread']]:>3s}  "
            f"grid {r[ix['launch__grid_size']]:>6s}  warps {num(r, ix, 'sm__warps_active.avg.pct_of_peak_sustained_active'):4.1f}%  "
            f"issue {num(r, ix, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):4.1f}%  "
            f"fp64 {num(r, ix, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):4.1f}%  "
            f"L1pipe {num(r, ix, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):4.1f}%  "
            f"dram {gbs(r, ix, units, 'dram__bytes_read.sum.per_second') + gbs(r, ix, units, 'dram__bytes_write.sum.per_second'):6.0f} GB/s | "
            + "  ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {100 * v / tot:.0f}%" for v, h in top))
        print("wrote", kernel)
# DRAM bytes per launch of every captured kernel -> profiles/<tag>_traffic.json (what bench.py's roofline.traffic reads)
def to_bytes(r, ix, units, k):
    scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    return num(r, ix, k) * scale.get(units[ix[k]], 0.0) if k in ix else 0.0


traffic = []
for f in sorted(os.listdir(OUT)):
    if not (f.startswith(f"ncu_{TAG}_") and f.endswith(".raw.csv")):
        continue
    rows = list(csv.reader(open(os.path.join(OUT, f))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    per, l1 = collections.OrderedDict(), {}
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
        name = "k_spmv2" if name.startswith("k_spmv2") else ("k_far_H" if name.startswith("k_far_H") else name)
        per.setdefault(name, []).append(to_bytes(r, ix, units, "dram__bytes_read.sum") + to_bytes(r, ix, units, "dram__bytes_write.sum"))
        l1.setdefault(name, []).append(num(r, ix, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"))
    for name, vals in per.items():
        traffic.append({"kernel": name, "cells": [8, 8, 8], "dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                        "l1_data_pipe_pct": sum(l1[name]) / len(l1[name]),
                        "source": f"ncu --set full --clock-control none, {f} (dram__bytes_read.sum + dram__bytes_write.sum)"})
if traffic:
    import json
    json.dump(traffic, open(os.path.join(PROF, f"{TAG}_traffic.json"), "w"), indent=1)
    print("wrote", f"{TAG}_traffic.json", [(t["kernel"], round(t["dram_bytes_per_launch"] / 1e6, 1)) for t in traffic])

if stall_lines:
    head = ["# ncu --set full (TATB 8x8x8, steady-state steps): per kernel, duration, registers, grid, resident warps, issue / fp64 /",
            "# L1 data-pipe utilisation (% of peak), DRAM read+write rate, and the share of each warp-stall reason"]
    open(os.path.join(PROF, f"{TAG}_stalls.txt"), "w").write("\n".join(head + stall_lines) + "\n")
    print("\n".join(stall_lines))

src = os.path.join(OUT, f"launches_{TAG}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PROF, f"{TAG}_launches_8x8x8.csv"))
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    t, c = collections.Counter(), collections.Counter()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].split("::")[-1]
        t[name] += float(r["Metric Value"]); c[name] += 1
    tot = sum(t.values())
    out = ["# ncu launch list of bench.py's timed region (10 steps, TATB 8x8x8), per-kernel totals; cold-cache serialised times",
           f"# launches {len(rows)}  total {tot / 1e6:.3f} ms  ({tot / 1e7:.3f} ms/step serialised)"]
    for k, v in t.most_common():
        out.append(f"{k:28s} {v / 1e6:8.3f} ms {100 * v / tot:5.1f}%  n={c[k]:4d} avg {v / c[k] / 1e3:8.1f} us")
    open(os.path.join(PROF, f"{TAG}_launches_8x8x8_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))
for a, b in ((f"bench_{TAG}.json", f"{TAG}_bench_8x8x8.json"), (f"bench_{TAG}_reference.json", f"{TAG}_bench_reference_arm.json")):
    if os.path.exists(os.path.join(OUT, a)):
        shutil.copy(os.path.join(OUT, a), os.path.join(PROF, b))
