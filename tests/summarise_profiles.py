"""Turn the raw evidence of tests/gpu_profile_round.sh (gpurun_out/*) into the committed summaries under profiles/:
    python tests/summarise_profiles.py r01d
  * profiles/<tag>_ncu_full_<kernel>.csv      selected metrics of every captured launch (from `ncu -i ... --page raw --csv`)
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
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]

for f in sorted(os.listdir(OUT)):
    if not (f.startswith(f"ncu_{TAG}_") and f.endswith(".ncu-rep")):
        continue
    kernel = f[len(f"ncu_{TAG}_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", os.path.join(OUT, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(os.path.join(PROF, f"{TAG}_ncu_full_{kernel}.csv"), "w", newline="") as fo:
        w = csv.writer(fo)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for name in WANT:
            if name in hdr:
                i = hdr.index(name)
                w.writerow([hdr[i], units[i]] + [d[i] for d in data])
    print("wrote", kernel)

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
