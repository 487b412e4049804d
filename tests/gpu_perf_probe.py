"""Quick per-phase timing of the resident MD step (development aid; bench.py is the contract benchmark)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import *  # noqa
from sw_reaxff_b200 import Rxb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
T = float(sys.argv[4]) if len(sys.argv) > 4 else 300.0
box, x, t, tag = tatb_cell(nx, nx, nx, scale=scale)
v = maxwell_velocities(t, T, 12345)
tab = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if tab > 0:
    CONTROL = control_variant("/tmp/control.tab%d" % tab, tab)
r = Rxb(0)
r.pair_settings(CONTROL)
r.pair_coeff(FFIELD, ELEMENTS)
r.fix_qeq(0.0, 10.0, 1e-6)
t0 = time.time()
r.md_setup(box, x, v, t, tag, MASS, dt=0.0625, every=5, thermo=5)
print("setup s", time.time() - t0, "counts", r.counts())
r.md_run(10)  # warm the QEq history
t0 = time.time()
r.md_run(steps)
dt = time.time() - t0
n = len(x)
print(f"atoms {n} steps {steps} wall {dt:.3f}s  ms/step {1e3*dt/steps:.3f}  atom-steps/s {n*steps/dt:.3e}")
c0 = r.counts()
r.profile(1)
r.md_run(steps)
prof = r.profile(0)
c1 = r.counts()
print(f"device ms/step (events) {r.md_last_run_ms()/steps:.3f}")
for nm, (ms, calls) in prof.items():
    print(f"  {nm:20s} {ms/steps:8.3f} ms/step  calls/step {calls/steps:6.1f}  avg {1e3*ms/max(calls,1):9.1f} us")
print("  qeq iters/step", (c1[7] - c0[7]) / steps, "launches/step", (c1[6] - c0[6]) / steps)
print("thermo", r.md_thermo())
