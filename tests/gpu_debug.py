"""Ad-hoc GPU-vs-oracle comparison with verbose diagnostics (run under gpurun while developing; not a pytest file)."""
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import *  # noqa
from sw_reaxff_b200 import Rxb


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / den


def make_rxb(tol=1e-6):
    r = Rxb(0)
    r.pair_settings(CONTROL)
    r.pair_coeff(FFIELD, ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


def bond_map(bs, bc_or_be, nbr, fld, N, is_end):
    d = {}
    for i in range(N):
        s = bs[i]
        e = bc_or_be[i] if is_end else s + bc_or_be[i]
        for p in range(s, e):
            d[(i, int(nbr[p]))] = fld[p]
    return d


def static_case(nx, ny, nz, perturb=0.0, seed=0, scale=1.0):
    print(f"=== static {nx}x{ny}x{nz} perturb={perturb} scale={scale}")
    cfg = static_config(nx, ny, nz, perturb=perturb, seed=seed, scale=scale, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    N = len(x)
    print("n", n, "N", N)
    # oracle: fresh static evaluation, zero charges then QEq from zero history
    q0 = np.zeros(N)
    o.set_atoms(n, x, ty, tg, q0)
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, 1e-6)
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
    t0 = time.time()
    mvo = o.qeq_pre_force(owner)
    o.compute()
    print("oracle time", time.time() - t0)
    r = make_rxb()
    r.set_atoms(n, x, ty, tg, q0, owner)
    t0 = time.time()
    r.neigh_build()
    # neighbour list parity
    off_o, nb_o = o.get_neighbors()
    off_g, nb_g = r.neighbors(0)
    ok = True
    for i in range(n):
        a = np.sort(nb_g[off_g[i]:off_g[i + 1]]); b = nb_o[off_o[i]:off_o[i + 1]]
        if len(a) != len(b) or np.any(a != b):
            ok = False
            print("verlet mismatch row", i, len(a), len(b)); break
    print("verlet list exact:", ok, "nnz", off_g[-1])
    mvg = r.qeq_pre_force()
    print("matvecs oracle", mvo, "gpu", mvg)
    qg = r.get_charges(); qo = o.q()
    print("q maxabs diff", np.abs(qg - qo).max(), "max|q|", np.abs(qo).max())
    # H parity
    offH, numH, colH, valH = o.qeq_H()
    num_g, idx_g, val_g = r.far()
    bad = 0
    for i in range(0, n, max(1, n // 50)):
        a = dict(zip(idx_g[off_g[i]:off_g[i] + num_g[i]].tolist(), val_g[off_g[i]:off_g[i] + num_g[i]].tolist()))
        b = dict(zip(colH[offH[i]:offH[i] + numH[i]].tolist(), valH[offH[i]:offH[i] + numH[i]].tolist()))
        if set(a) != set(b):
            bad += 1
        else:
            m = max(abs(a[k] - b[k]) for k in b) / max(abs(v) for v in b.values()) if b else 0
            if m > 1e-12: bad += 1; print("H row", i, "err/max", m)
    print("H rows bad:", bad)
    # use oracle charges on the GPU for the force comparison so QEq tolerance does not blur it
    r.set_charges(qo)
    res = r.pair_compute(True, True)
    print("gpu time", time.time() - t0)
    eo, vo = o.energies()
    pv = res["pvector"]
    eg = np.array([pv[0], 0, 0, pv[2], pv[4], pv[5], pv[6], pv[7], pv[8], pv[9], pv[10], pv[11], pv[13]])
    for k, nm in enumerate(E_NAMES):
        if nm in ("e_ov", "e_un"):
            continue
        print(f"  {nm:7s} oracle {eo[k]: .10e} gpu {eg[k]: .10e} rel {abs(eg[k]-eo[k])/max(abs(eo[k]),1e-300):.2e}")
    print(f"  e_ov+un oracle {eo[1]+eo[2]: .10e} gpu {pv[1]: .10e} rel {abs(pv[1]-eo[1]-eo[2])/abs(eo[1]+eo[2]):.2e}")
    fo = o.forces(); fg = res["f"]
    print("forces: max|f|", np.abs(fo).max(), "maxabs diff", np.abs(fg - fo).max(), "rel", rel(fg, fo))
    worst = np.argmax(np.abs(fg - fo).max(1))
    print("  worst atom", worst, "local" if worst < n else "ghost", fg[worst], fo[worst])
    print("virial oracle", vo, "\n       gpu   ", res["virial"], "rel", rel(res["virial"], vo))
    # bonds
    bs, be, nbr, sym, fld = o.bonds()
    gbs, gbc, gnbr, gsym, gfld = r.bonds()
    mo = bond_map(bs, be, nbr, fld, N, True)
    mg = bond_map(gbs, gbc, gnbr, gfld, N, False)
    print("bonds oracle", len(mo), "gpu", len(mg), "same keys", set(mo) == set(mg))
    if set(mo) == set(mg):
        A = np.array([mg[k] for k in mo]); B = np.array([mo[k] for k in mo])
        names = ["d", "dx", "dy", "dz", "BO", "BO_s", "BO_pi", "BO_pi2"] + [f"dBOp{t}" for t in range(3)] + [f"dlnpi{t}" for t in range(3)] + [f"dlnpi2{t}" for t in range(3)] + ["C1dbo", "C2dbo", "C3dbo", "C1dbopi", "C2dbopi", "C3dbopi", "C4dbopi", "C1dbopi2", "C2dbopi2", "C3dbopi2", "C4dbopi2", "Cdbo", "Cdbopi", "Cdbopi2"]
        for c, nm in enumerate(names):
            den = max(np.abs(B[:, c]).max(), 1e-300)
            e = np.abs(A[:, c] - B[:, c]).max() / den
            if e > 1e-10:
                print(f"   bond field {nm}: rel {e:.3e}")
    wo = o.workspace(); wg = r.workspace()
    for c, nm in [(0, "total_bo"), (1, "Delta_boc"), (2, "Deltap"), (3, "Deltap_boc"), (4, "Delta"), (6, "Delta_val"), (7, "vlpex"), (8, "nlp"), (9, "Delta_lp"), (11, "dDelta_lp"), (13, "Delta_lp_temp")]:
        e = np.abs(wg[:, c] - wo[:, c]).max()
        if e > 1e-10:
            print(f"   workspace {nm}: maxabs {e:.3e}")
    cd_o = o.cddelta()
    print("CdDelta rel", rel(wg[:, 15], cd_o))
    print("counts", r.counts())
    return r, o


def md_case(nx, steps, T=300.0, dt=0.0625):
    print(f"=== md {nx}^3 steps={steps}")
    box, x, t, tag = tatb_cell(nx, nx, nx)
    v = maxwell_velocities(t, T, 12345)
    o = Oracle()
    t0 = time.time()
    o.md_init(box, x, v, t, tag, dt=dt)
    o.md_run(steps)
    print("oracle md time", time.time() - t0)
    ro = o.md_get()
    r = make_rxb()
    t0 = time.time()
    r.md_setup(box, x, v, t, tag, MASS, dt=dt, every=5, thermo=1)
    r.md_run(steps)
    rg = r.md_get()
    print("gpu md time", time.time() - t0)
    th = r.md_thermo()
    print("x diff", np.abs(rg["x"] - ro["x"]).max(), "v rel", rel(rg["v"], ro["v"]), "f rel", rel(rg["f"], ro["f"]), "q diff", np.abs(rg["q"] - ro["q"]).max())
    print("pe oracle", ro["pe"], "gpu", th["pe"], "ke oracle", ro["ke"], "gpu", th["ke"])
    print("matvecs oracle", o.md_matvecs(), "counts", r.counts())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "static"):
        static_case(1, 1, 1)
        static_case(1, 1, 1, perturb=0.1, seed=1)
        static_case(2, 2, 2, perturb=0.05, seed=2, scale=0.93)
    if which in ("all", "md"):
        md_case(1, 10)
        md_case(2, 10, T=1000.0)
