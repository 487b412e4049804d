"""Quick A/B probe on the GPU box: `python tests/gpu_perf_probe.py [nx] [steps] [dt] [profile]` prints ms/step of the
resident run (overlapped two-stream step) and, with profile=1, the per-phase device times of the serialised step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import *  # noqa
from sw_reaxff_b200 import Rxb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.625
prof = int(sys.argv[4]) if len(sys.argv) > 4 else 0
box, x, t, tag = tatb_cell(nx, nx, nx)
v = maxwell_velocities(t, 300.0, 12345)
r = Rxb(0)
tabn = int(os.environ.get("RXB_PROBE_TAB", "0"))      # > 0: table mode with that many knots (control file variant)
r.pair_settings(control_variant("/tmp/control.probe_tab", tabn) if tabn > 0 else CONTROL)
r.pair_coeff(FFIELD, ELEMENTS)
r.fix_qeq(0.0, 10.0, 1e-6)
r.md_setup(box, x, v, t, tag, MASS, dt=dt, every=5, thermo=5)
r.md_run(10)
c0 = r.counts()
r.md_run(steps)
ms = r.md_last_run_ms()
c1 = r.counts()
line = f"probe nx={nx} dt={dt} env={ {k: v for k, v in os.environ.items() if k.startswith('RXB_')} }: {ms / steps:.3f} ms/step, {(c1[7] - c0[7]) / steps:.1f} CG it/step"
if prof:
    r.profile(1)
    r.md_run(steps)
    p = r.profile(0)
    line += "  | " + " ".join(f"{k}={p[k][0] / steps:.3f}" for k in p)
print(line)
