"""Generates tests/golden/*.npz from the CPU oracle (oracle/liboracle.so).

The reference ships NO golden vectors, logs or tests and cannot run here (Sunway toolchain + patched
LAMMPS), so these fixtures pin the oracle against drift and give the GPU path committed answers to hit on the GPU box,
where /root/reference does not exist.  What anchors the oracle itself: (1) oracle/_ref — the reference's own parsers and
live serial routines compiled from /root/reference and compared in tests/test_oracle_vs_ref.py, (2) finite-difference
force checks and invariants in tests/test_oracle_physics.py, (3) the step-0 energies of this very input agree with the
stock LAMMPS examples/reax/tatb log to all printed digits (PotEng -44760.998; recorded from memory, not gated on).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402


def one(name, nx, ny, nz, perturb, seed, scale, tol):
    cfg = H.static_config(nx, ny, nz, perturb=perturb, seed=seed, scale=scale, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    o.set_atoms(n, x, ty, tg, np.zeros(len(x)))
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, tol)
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
    mv = o.qeq_pre_force(owner)
    o.compute()
    e, vir = o.energies()
    f = o.forces()
    fl = f[:n].copy()
    np.add.at(fl, owner, f[n:])  # reverse_comm
    bs, be, nbr, sym, fld = o.bonds()
    w = o.workspace()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        nx=nx, ny=ny, nz=nz, perturb=perturb, seed=seed, scale=scale, tol=tol,
        n=n, nall=len(x), energies=e, virial=vir, f_local=fl, q_local=o.q()[:n], matvecs=np.array(mv),
        nbonds_per_atom=(be - bs)[:n].astype(np.int16), total_bo=w[:n, 0], nlp=w[:n, 8],
        verlet_count=np.diff(o.get_neighbors()[0])[:n].astype(np.int32))
    print(name, "n", n, "N", len(x), "pe", e.sum(), "matvecs", mv)


if __name__ == "__main__":
    one("tatb_1x1x1", 1, 1, 1, 0.0, 0, 1.0, 1e-6)
    one("tatb_1x1x1_perturbed", 1, 1, 1, 0.1, 1, 1.0, 1e-10)
    one("tatb_2x1x1_compressed", 2, 1, 1, 0.05, 2, 0.92, 1e-10)
