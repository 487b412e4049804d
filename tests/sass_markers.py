"""cuobjdump -sass <lib> | python tests/sass_markers.py > profiles/<tag>_sass_markers.txt : per-kernel counts of the SASS
mnemonics that matter on this path (PDL, bulk-async copies, mbarrier, fp64 pipe, SFU, memory)."""
import collections
import re
import subprocess
import sys

cur = None
cnt = collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        cnt[cur][m.group(1).split(".")[0]] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
short = {n: re.sub(r"rxb::(\(anonymous namespace\)::)?", "", d).split("(")[0] for n, d in zip(names, dem)}
keys = ["ACQBULK", "PREEXIT", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "LDS", "STG", "ATOMG", "RED", "BAR"]
print("# cuobjdump -sass sw_reaxff_b200/librxb200.so (sm_100a): instruction counts per kernel")
print("# ACQBULK/PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch); UBLKCP = cp.async.bulk (TMA engine);")
print("# SYNCS = mbarrier; DFMA/DMUL/DADD = fp64 pipe; MUFU = SFU; LDG/LDS/STG/ATOMG/RED = memory")
print(f"{'kernel':44s} " + " ".join(f"{k:>7s}" for k in keys) + "   total")
for f in sorted(cnt, key=lambda f: -sum(cnt[f].values())):
    d = short[f]
    if "k_" not in d:
        continue
    print(f"{d[:44]:44s} " + " ".join(f"{cnt[f][k]:7d}" for k in keys) + f"  {sum(cnt[f].values()):6d}")
