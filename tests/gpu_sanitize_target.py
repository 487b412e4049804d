"""compute-sanitizer target: a 2x2x2 TATB cell (3072 atoms) through both entry paths - 6 resident steps (reneighbouring at
step 5, bond table + species) and one plugin-path evaluation with host buffers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as H
from sw_reaxff_b200 import Rxb

box, x, t, tag = H.tatb_cell(2, 2, 2)
v = H.maxwell_velocities(t, 1500.0, 3)
r = Rxb(0)
r.pair_settings(H.CONTROL); r.pair_coeff(H.FFIELD, H.ELEMENTS); r.fix_qeq(0.0, 10.0, 1e-6)
r.md_setup(box, x, v, t, tag, H.MASS, dt=0.25, every=5, thermo=1)
r.species_config(1, 2, 2, natoms=len(x))
r.md_run(6)
bt = r.bond_table()
print("resident: pe", r.md_thermo()["pe"], "bond table entries", len(bt["nbr"]), "species outputs", len(r.species_log()))
cfg = H.static_config(1, 1, 1, perturb=0.05, seed=2, qeq=False)
p = Rxb(0)
p.pair_settings(H.CONTROL); p.pair_coeff(H.FFIELD, H.ELEMENTS); p.fix_qeq(0.0, 10.0, 1e-6)
p.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], np.zeros(len(cfg["x"])), cfg["owner"])
p.neigh_build()
p.qeq_pre_force_async()
res = p.pair_compute(True, True)
print("plugin: e", res["eng"].sum(), "iterations", p.qeq_matvecs())
# host-planned halo mode (one-rank communicator: plan, permutation, peer windows / NCCL self paths, q forward), 5 MD steps
import lammps_comm as LC
box1, x1, t1, tag1 = H.tatb_cell(1, 1, 1)
comm = LC.HostComm(box1, (1, 1, 1), 12.5)
a = LC.host_md(Rxb, H, comm, 0, 0, box1, x1, H.maxwell_velocities(t1, 1500.0, 5), t1, tag1, 5, every=3, uid=Rxb.dist_unique_id(),
               use_comm=True, tol=1e-8)
print("comm mode: pe", a["pe"][-1], "ghost q err", a["ghost_q_err"])
# table mode with the coefficient tables in shared memory (force-only step) and in L2 (energy step)
ctl = H.control_variant("/tmp/control.sanitize_tab", 300)
tb = Rxb(0)
tb.pair_settings(ctl); tb.pair_coeff(H.FFIELD, H.ELEMENTS)
tb.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], np.zeros(len(cfg["x"])), cfg["owner"])
tb.neigh_build()
f0 = tb.pair_compute(False, False)["f"]
f1 = tb.pair_compute(True, True)["f"]
print("table mode: max |f_smem - f_l2|", float(np.abs(f0 - f1).max()))
