"""ncu target: `ncu --profile-from-start off ... python tests/gpu_ncu_target.py NX STEPS [DT]` profiles STEPS steady-state steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import *  # noqa
from sw_reaxff_b200 import Rxb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.625
box, x, t, tag = tatb_cell(nx, nx, nx)
v = maxwell_velocities(t, 300.0, 12345)
r = Rxb(0)
r.pair_settings(CONTROL)
r.pair_coeff(FFIELD, ELEMENTS)
r.fix_qeq(0.0, 10.0, 1e-6)
r.md_setup(box, x, v, t, tag, MASS, dt=dt, every=5, thermo=5)
r.md_run(11)
r.profiler_range(1)
r.md_run(steps)
r.profiler_range(0)
print("done", r.counts())
