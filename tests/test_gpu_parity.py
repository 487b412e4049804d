"""GPU parity tests (-m gpu): every stage of the CUDA path against the CPU oracle on identical inputs, through the C ABI.

Tolerances (north_star): forces and per-term energies 1e-8 relative (fp64, atomic-order nondeterminism only);
charges within the CG tolerance; neighbour/bond/hbond index work bit-exact.
"""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-8


def make_rxb(tol=1e-6):
    from sw_reaxff_b200 import Rxb
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def pvector_from_oracle(e):
    return np.array([e[0], e[1] + e[2], e[3], 0.0, e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11], 0.0, e[12]])


def bond_dict(bs, cnt_or_end, nbr, fld, N, is_end):
    out = {}
    for i in range(N):
        s = bs[i]
        e = cnt_or_end[i] if is_end else s + cnt_or_end[i]
        for p in range(s, e):
            out[(i, int(nbr[p]))] = fld[p]
    return out


CASES = [
    dict(id="cell", nx=1, ny=1, nz=1, perturb=0.0, seed=0, scale=1.0),
    dict(id="perturbed1", nx=1, ny=1, nz=1, perturb=0.1, seed=1, scale=1.0),
    dict(id="perturbed2", nx=1, ny=1, nz=1, perturb=0.1, seed=2, scale=1.0),
    dict(id="compressed2x2x2", nx=2, ny=2, nz=2, perturb=0.05, seed=3, scale=0.93),
    dict(id="ragged2x1x1", nx=2, ny=1, nz=1, perturb=0.2, seed=4, scale=1.05),
]


@pytest.fixture(scope="module", params=CASES, ids=[c["id"] for c in CASES])
def case(request):
    c = request.param
    cfg = H.static_config(c["nx"], c["ny"], c["nz"], perturb=c["perturb"], seed=c["seed"], scale=c["scale"], qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    q0 = np.zeros(len(x))
    o.set_atoms(n, x, ty, tg, q0)
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, 1e-6)
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
    mvo = o.qeq_pre_force(owner)
    o.compute()
    r = make_rxb(1e-6)
    r.set_atoms(n, x, ty, tg, q0, owner)
    r.neigh_build()
    mvg = r.qeq_pre_force()
    qg = r.get_charges()
    far = r.far()
    r.set_charges(o.q())  # identical charges for the force comparison (QEq has its own tests)
    res = r.pair_compute(True, True)
    return dict(cfg=cfg, o=o, r=r, mvo=mvo, mvg=mvg, qg=qg, res=res, far=far)


def test_neighbor_list_exact(case):
    o, r, n = case["o"], case["r"], case["cfg"]["n"]
    off_o, nb_o = o.get_neighbors()
    off_g, nb_g = r.neighbors(0)
    assert np.array_equal(np.diff(off_g), np.diff(off_o)[:n])
    for i in range(n):
        assert np.array_equal(np.sort(nb_g[off_g[i]:off_g[i + 1]]), nb_o[off_o[i]:off_o[i + 1]])
    # bond-candidate rows (all atoms, ghosts too) hold exactly the pairs within (bond reach + skin), where the reach is the
    # largest distance at which any element pair can have BO' >= bo_cut (3.354 A for this force field, N-N)
    cut = r.cutoffs()
    assert abs(cut["verlet"] - 12.5) < 1e-12 and 3.35 < cut["bond_reach"] < 3.36 and abs(cut["bond_candidates"] - cut["bond_reach"] - 2.5) < 1e-12
    off_b, nb_b = r.neighbors(1)
    x = case["cfg"]["x"]
    for i in list(range(0, len(x), 97)):
        row = set(nb_b[off_b[i]:off_b[i + 1]].tolist())
        close = [j for j in nb_o[off_o[i]:off_o[i + 1]] if np.linalg.norm(x[j] - x[i]) <= cut["bond_candidates"]]
        assert set(close) == row


def test_far_list_and_H(case):
    o, r, n = case["o"], case["r"], case["cfg"]["n"]
    offH, numH, colH, valH = o.qeq_H()
    num_g, idx_g, val_g = case["far"]
    off_g, _ = r.neighbors(0)
    assert np.array_equal(num_g, numH)
    hmax = np.abs(valH).max()
    for i in range(0, n, max(1, n // 64)):
        a = dict(zip(idx_g[off_g[i]:off_g[i] + num_g[i]].tolist(), val_g[off_g[i]:off_g[i] + num_g[i]].tolist()))
        b = dict(zip(colH[offH[i]:offH[i] + numH[i]].tolist(), valH[offH[i]:offH[i] + numH[i]].tolist()))
        assert a.keys() == b.keys()
        assert max(abs(a[k] - b[k]) for k in b) < 1e-12 * hmax


def test_qeq_charges_and_iterations(case):
    o, n = case["o"], case["cfg"]["n"]
    # same pipelined-CG iteration counts as the reference's two serial solves; the stopping test is a threshold on a
    # rounded quantity, so allow the last iteration to fall on either side of it
    assert abs(case["mvg"][0] - case["mvo"][0]) <= 1 and abs(case["mvg"][1] - case["mvo"][1]) <= 1
    qo = o.q()
    # both solves stop at a relative preconditioned residual of 1e-6: each answer is within ~1e-5 e of the exact one
    assert np.abs(case["qg"] - qo).max() < 2e-5
    assert abs(case["qg"][:n].sum()) < 1e-9
    assert np.array_equal(case["qg"][n:], case["qg"][case["cfg"]["owner"]])   # ghost charges forwarded


@pytest.mark.parametrize("exact", [True, False], ids=["exact_H", "packed_H"])
def test_qeq_tight_tolerance_same_solution(exact):
    """With the tolerance driven to 1e-10 the dual-RHS device solve and the oracle's two serial solves meet.  With the
    exact 12-byte H entries the iteration counts are the reference's (+-2 at the threshold); with the default packed
    8-byte entries (H quantised to 2^-38) the solution is the same to 1e-8 and the counts may differ by a few iterations,
    because a 1e-10 relative residual sits only two orders above the quantisation."""
    cfg = H.static_config(1, 1, 1, perturb=0.1, seed=11, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    rng = np.random.default_rng(5)
    sh = rng.normal(scale=0.05, size=(n, 5)); th = rng.normal(scale=0.05, size=(n, 5))   # non-trivial history -> extrapolated guess
    o.set_atoms(n, x, ty, tg, np.zeros(len(x)))
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, 1e-10)
    o.qeq_set_hist(sh, th)
    mvo = o.qeq_pre_force(owner)
    r = make_rxb(1e-10)
    r.set_h_exact(exact)
    r.set_atoms(n, x, ty, tg, None, owner)
    r.neigh_build()
    r.qeq_set_history(sh, th)
    mvg = r.qeq_pre_force()
    slack = 8    # at a 1e-10 residual the stopping iteration is sensitive to the summation order of the SpMV rows
    assert abs(mvg[0] - mvo[0]) <= slack and abs(mvg[1] - mvo[1]) <= slack, (mvg, mvo)
    assert max(mvo) < 200
    assert np.abs(r.get_charges() - o.q()).max() < 1e-8, np.abs(r.get_charges() - o.q()).max()
    so, to = o.qeq_get_hist()
    sg, tg_ = r.qeq_get_history()
    assert np.abs(sg - so).max() < 1e-7 and np.abs(tg_ - to).max() < 1e-7, (np.abs(sg - so).max(), np.abs(tg_ - to).max())
    assert np.array_equal(sg[:, 1:], sh[:, :4])


def test_bond_list_and_bond_orders(case):
    o, r = case["o"], case["r"]
    N = len(case["cfg"]["x"])
    bs, be, nbr, sym, fld = o.bonds()
    gbs, gbc, gnbr, gsym, gfld = r.bonds()
    assert np.array_equal(gbc, be - bs)                    # bonds per atom: exact
    for i in range(N):                                     # rows in ascending neighbour order, same neighbours
        assert np.array_equal(gnbr[gbs[i]:gbs[i] + gbc[i]], nbr[bs[i]:be[i]])
    mo = bond_dict(bs, be, nbr, fld, N, True)
    mg = bond_dict(gbs, gbc, gnbr, gfld, N, False)
    A = np.array([mg[k] for k in mo]); B = np.array([mo[k] for k in mo])
    for c in range(31):
        den = max(np.abs(B[:, c]).max(), 1e-300)
        assert np.abs(A[:, c] - B[:, c]).max() / den < 1e-9, f"bond field {c}"
    # sym_index points back
    rows = np.repeat(np.arange(N), gbc)
    order = np.concatenate([np.arange(gbs[i], gbs[i] + gbc[i]) for i in range(N)])
    owner_of_slot = np.empty(len(gnbr), dtype=np.int64); owner_of_slot[order] = rows
    assert np.array_equal(gnbr[gsym[order]], owner_of_slot[order])


def test_workspace(case):
    wo, wg = case["o"].workspace(), case["r"].workspace()
    for c in (0, 1, 2, 3, 4, 6, 7, 8, 9, 11, 13):
        assert np.abs(wg[:, c] - wo[:, c]).max() < 1e-10, c
    assert rel(wg[:, 15], case["o"].cddelta()) < 1e-9


def test_energies_forces_virial(case):
    o, res = case["o"], case["res"]
    eo, vo = o.energies()
    pvo = pvector_from_oracle(eo)
    for k in range(14):
        if pvo[k] != 0.0:
            assert abs(res["pvector"][k] - pvo[k]) <= RTOL * abs(pvo[k]), (k, res["pvector"][k], pvo[k])
    assert abs(res["eng"].sum() - eo.sum()) <= RTOL * abs(eo.sum())
    fo = o.forces()
    assert rel(res["f"], fo) < RTOL
    assert rel(res["virial"], vo) < RTOL


@pytest.mark.parametrize("name", ["tatb_1x1x1", "tatb_1x1x1_perturbed", "tatb_2x1x1_compressed"])
def test_gpu_matches_committed_golden(name):
    """End to end (neighbours + QEq + forces + reverse) against the committed fixtures; no oracle call involved."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    box, x, t, tag = H.tatb_cell(int(g["nx"]), int(g["ny"]), int(g["nz"]), float(g["perturb"]), int(g["seed"]), float(g["scale"]))
    r = make_rxb(float(g["tol"]))
    r.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS, thermo=1)
    out = r.md_get()
    th = r.md_thermo()
    assert abs(th["pe"] - g["energies"].sum()) < 1e-7 * abs(g["energies"].sum())
    np.testing.assert_allclose(th["pvector"], pvector_from_oracle(g["energies"]), rtol=2e-7, atol=1e-6)
    assert np.abs(out["q"] - g["q_local"]).max() < 10 * float(g["tol"])
    assert rel(out["f"], g["f_local"]) < 1e-5 if float(g["tol"]) > 1e-8 else rel(out["f"], g["f_local"]) < 1e-7


def test_md_trajectory_vs_oracle():
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    v = H.maxwell_velocities(t, 300.0, 12345)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-10)
    o.md_run(12)     # crosses two reneighbouring steps
    ro = o.md_get()
    r = make_rxb(1e-10)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=1)
    r.md_run(12)
    rg = r.md_get()
    th = r.md_thermo()
    assert np.abs(rg["x"] - ro["x"]).max() < 1e-9
    assert rel(rg["v"], ro["v"]) < 1e-7
    assert rel(rg["f"], ro["f"]) < 1e-6
    assert np.abs(rg["q"] - ro["q"]).max() < 1e-7
    assert abs(th["pe"] - ro["pe"]) < 1e-8 * abs(ro["pe"])
    assert abs(th["ke"] - ro["ke"]) < 1e-7 * abs(ro["ke"])


def test_md_trajectory_hot_compressed_vs_oracle():
    """BASELINE.json's hot-compressed case (0.90 linear scale, 3000 K) at a size the oracle runs in seconds: more
    bonds/angles/torsions per atom, atoms crossing the box faces, QEq far from its initial guess."""
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=0.90)
    v = H.maxwell_velocities(t, 3000.0, 4242)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-10)
    o.md_run(12)
    ro = o.md_get()
    r = make_rxb(1e-10)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=1)
    r.md_run(12)
    rg = r.md_get()
    th = r.md_thermo()
    assert np.abs(rg["x"] - ro["x"]).max() < 1e-8
    assert rel(rg["f"], ro["f"]) < 1e-6
    assert np.abs(rg["q"] - ro["q"]).max() < 1e-7
    assert abs(th["pe"] - ro["pe"]) < 1e-8 * abs(ro["pe"])
    assert abs(th["ke"] - ro["ke"]) < 1e-7 * abs(ro["ke"])


def test_md_energy_conservation_gpu():
    box, x, t, tag = H.tatb_cell(2, 2, 2)
    v = H.maxwell_velocities(t, 300.0, 777)
    r = make_rxb(1e-8)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=1)
    t0 = r.md_thermo()
    r.md_run(100)
    t1 = r.md_thermo()
    drift = abs((t1["pe"] + t1["ke"]) - (t0["pe"] + t0["ke"]))
    # The reference's constants make E(x, q(x)) slightly non-conservative by construction: the pair style uses
    # C_ele = 332.06371 while fix qeq/reax minimises with 14.4 eV*A x 23.02 = 331.488 (reaxc_defs_sunway.h:62,69;
    # fix_qeq_reax_sunway.cpp:978), a 0.17 % mismatch that breaks the Hellmann-Feynman condition (measured: 0.1
    # kcal/mol/A on individual forces; with consistent constants the QEq-resolved finite difference matches to 1e-5),
    # plus the hbond_cut / BO-threshold discontinuities of the force field.  Observed drift: +0.46 % of KE0 over 100 steps,
    # identical for tol 1e-8 and 1e-10 and identical in the CPU oracle.  The bound below catches integrator/halo bugs
    # (which show up as percent-level jumps at reneighbouring steps), not that inherited drift.
    assert drift < 1e-2 * t0["ke"], (t0, t1)


def test_plugin_path_host_buffers_matches_resident_path():
    """The LAMMPS-facing calls (host x in, host f out each step) and the resident run give the same forces."""
    cfg = H.static_config(1, 1, 1, perturb=0.05, seed=9, qeq=False)
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    r = make_rxb(1e-10)
    r.set_atoms(n, x, ty, tg, None, None)          # owner map derived from tags (atom->map)
    r.neigh_build()
    r.qeq_pre_force()
    a = r.pair_compute(True, True)
    x2 = x + 0.0
    r.set_positions(x2)
    r.qeq_pre_force()
    b = r.pair_compute(True, True)
    assert rel(b["f"], a["f"]) < 1e-6              # second solve starts from history: same answer within tolerance
    r2 = make_rxb(1e-10)
    r2.set_atoms(n, x, ty, tg, None, owner)
    r2.neigh_build()
    r2.qeq_pre_force()
    c = r2.pair_compute(True, True)
    assert rel(c["f"], a["f"]) < 1e-12


def test_empty_and_tiny_inputs():
    r = make_rxb()
    x = np.array([[0.0, 0.0, 0.0], [1.2, 0.0, 0.0], [40.0, 40.0, 40.0]])
    r.set_atoms(3, x, np.array([1, 3, 2], dtype=np.int32), np.array([1, 2, 3], dtype=np.int32))
    r.neigh_build()
    r.qeq_pre_force()
    out = r.pair_compute(True, True)
    assert np.isfinite(out["f"]).all() and np.abs(out["f"][2]).max() == 0.0   # isolated atom feels nothing
    assert np.abs(out["f"][0] + out["f"][1]).max() < 1e-9 * np.abs(out["f"]).max()
    o = H.Oracle()
    o.set_atoms(3, x, np.array([1, 3, 2], dtype=np.int32), np.array([1, 2, 3], dtype=np.int32), r.get_charges())
    o.build_neighbors(12.5)
    o.compute()
    assert rel(out["f"], o.forces()) < 1e-9


def _forces_and_energies(ffield, elements, perturb, seed, qeq=True):
    """One (QEq +) force evaluation on the oracle and on the GPU with the same force field / element map.
    qeq=False: the pair style alone with fixed charges (a run without fix qeq/reax)."""
    from sw_reaxff_b200 import Rxb
    orc = H.Oracle(ffield=ffield, elements=elements)
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, qeq=False, oracle=orc)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    q0 = np.zeros(len(x))
    if not qeq:
        ql = np.random.default_rng(seed).uniform(-0.4, 0.4, n)
        q0 = np.concatenate([ql, ql[owner]])
    o.set_atoms(n, x, ty, tg, q0)
    o.build_neighbors(12.5)
    if qeq:
        o.qeq_init(0.0, 10.0, 1e-10)
        o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
        o.qeq_pre_force(owner)
    o.compute()
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(ffield, elements)
    r.fix_qeq(0.0, 10.0, 1e-10)
    r.set_atoms(n, x, ty, tg, q0, owner)
    r.neigh_build()
    if qeq:
        r.qeq_pre_force()
    qg = r.get_charges()
    res = r.pair_compute(True, True)
    return cfg, o, qg, res


@pytest.mark.parametrize("vdw_type", [3, 2])
def test_inner_wall_vdw_forms(vdw_type, tmp_path):
    """Force-field variants that select vdw_type 3 (shielding + inner wall) and 2 (inner wall only): branches of
    reaxc_nonbonded_sw64.c:118-165 the shipped TATB force field never takes."""
    ff = H.ffield_variant(tmp_path / "ffield.v", vdw_type)
    cfg, o, qg, res = _forces_and_energies(ff, H.ELEMENTS, 0.1, 11)
    assert int(o.params_dump()[1]) == vdw_type
    eo, vo = o.energies()
    assert np.abs(qg[:cfg["n"]] - o.q()[:cfg["n"]]).max() < 1e-8
    assert rel(res["pvector"], pvector_from_oracle(eo)) < 1e-8 and abs(eo[10]) > 10.0
    assert abs(res["eng"].sum() - eo.sum()) < 1e-8 * abs(eo.sum())
    assert rel(res["f"], o.forces()) < 1e-7
    assert rel(res["virial"], vo) < 1e-7


def test_null_mapped_element_is_ignored_like_the_reference():
    """pair_coeff * * ffield C H O NULL: atoms of the unmapped LAMMPS type take part in no interaction
    (type < 0 skips in every reference loop, e.g. reaxc_nonbonded_sw64.c:78,87).  Run without fix qeq/reax: the
    reference's QEq divides by eta = 0 for such atoms (Pair::extract gives chi = eta = 0, Hdia_inv = 1/eta), so the
    combination is only defined with fixed charges."""
    cfg, o, qg, res = _forces_and_energies(H.FFIELD, ["C", "H", "O", "NULL"], 0.05, 12, qeq=False)
    eo, vo = o.energies()
    n = cfg["n"]
    null_atoms = np.nonzero(cfg["type"][:n] == 4)[0]
    assert len(null_atoms) == 96
    assert np.abs(res["f"][null_atoms]).max() == 0.0 and np.abs(o.forces()[null_atoms]).max() == 0.0
    assert rel(res["pvector"], pvector_from_oracle(eo)) < 1e-8
    assert rel(res["f"], o.forces()) < 1e-7
    assert np.array_equal(qg[:n], o.q()[:n])


def test_tabulated_long_range_mode(tmp_path):
    """control: tabulate_long_range 10000 -> the nonbonded kernel evaluates the cubic-spline tables (a9' / f4).  GPU vs
    the oracle's table mode to 1e-8, and — because the spline error at N = 10000 is ~1e-12 — also vs the analytic form."""
    from sw_reaxff_b200 import Rxb
    ctl = H.control_variant(tmp_path / "control.tab", 10000)
    out = {}
    for name, control in (("table", ctl), ("analytic", H.CONTROL)):
        cfg = H.static_config(1, 1, 1, perturb=0.1, seed=21, qeq=True, oracle=H.Oracle(control=control))
        o = cfg["oracle"]
        n, x, ty, tg, owner, q = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"], cfg["q"]
        o.set_atoms(n, x, ty, tg, q)
        o.build_neighbors(12.5)
        o.compute()
        out[name] = (o.energies()[0], o.forces(), o.energies()[1])
    r = Rxb(0)
    r.pair_settings(ctl)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.set_atoms(n, x, ty, tg, q, owner)
    r.neigh_build()
    res = r.pair_compute(True, True)
    for name, tol in (("table", 1e-8), ("analytic", 1e-8)):
        eo, fo, vo = out[name]
        assert rel(res["pvector"], pvector_from_oracle(eo)) < tol, name
        assert rel(res["f"], fo) < 10 * tol, name
        assert rel(res["virial"], vo) < 10 * tol, name


def test_tabulated_shared_memory_mode(tmp_path, monkeypatch):
    """Coarse tables (tabulate_long_range 300: the CEvd / CEclmb coefficient sets of the 10 TATB type pairs fit in shared
    memory) switch force-only steps to k_nonbonded_tab_smem.  It must give the forces of the L2-resident table kernel
    (same records, same operations) and the oracle's table-mode forces; the price of the coarse table against the analytic
    form is measured and bounded here."""
    from sw_reaxff_b200 import Rxb
    ctl = H.control_variant(tmp_path / "control.tab300", 300)
    out = {}
    for name, control in (("table", ctl), ("analytic", H.CONTROL)):
        cfg = H.static_config(2, 1, 1, perturb=0.1, seed=22, qeq=True, oracle=H.Oracle(control=control))
        o = cfg["oracle"]
        n, x, ty, tg, owner, q = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"], cfg["q"]
        o.set_atoms(n, x, ty, tg, q)
        o.build_neighbors(12.5)
        o.compute()
        out[name] = o.forces()
    r = Rxb(0)
    r.pair_settings(ctl)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.set_atoms(n, x, ty, tg, q, owner)
    r.neigh_build()
    f_ev = r.pair_compute(True, True)["f"].copy()          # energy step: L2-resident table kernel
    monkeypatch.setenv("RXB_TAB_SMEM", "1")
    f_smem = r.pair_compute(False, False)["f"].copy()      # force-only step: shared-memory tables
    monkeypatch.setenv("RXB_TAB_SMEM", "0")
    f_l2 = r.pair_compute(False, False)["f"].copy()        # force-only step, L2 tables
    assert rel(f_smem, f_l2) < 1e-12 and rel(f_smem, f_ev) < 1e-12
    assert rel(f_smem, out["table"]) < 1e-7
    dev = rel(f_smem, out["analytic"])
    print(f"coarse-table deviation from the analytic forces (N = 300): {dev:.3e}")
    assert dev < 2e-3


def test_neighbor_rows_outgrowing_the_stride_are_rebuilt():
    """Neighbour rows live at a fixed stride sized by the previous build.  Re-using one handle for a configuration whose
    rows are ~60 % longer (1.05 -> 0.90 linear scale) must trigger the re-run with a larger stride and still give the
    oracle's list and energies."""
    from sw_reaxff_b200 import Rxb
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, 1e-8)
    for scale in (1.05, 0.90):
        cfg = H.static_config(1, 1, 1, perturb=0.05, seed=31, scale=scale, qeq=False)
        o = cfg["oracle"]
        n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
        q0 = np.zeros(len(x))
        o.set_atoms(n, x, ty, tg, q0)
        o.build_neighbors(12.5)
        o.qeq_init(0.0, 10.0, 1e-8)
        o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
        o.qeq_pre_force(owner)
        o.compute()
        r.set_atoms(n, x, ty, tg, q0, owner)
        r.qeq_set_history(np.zeros((n, 5)), np.zeros((n, 5)))
        r.neigh_build()
        off_o, nb_o = o.get_neighbors()
        off_g, nb_g = r.neighbors(0)
        assert np.array_equal(np.diff(off_g), np.diff(off_o)[:n])
        for i in range(0, n, 7):
            assert np.array_equal(np.sort(nb_g[off_g[i]:off_g[i + 1]]), nb_o[off_o[i]:off_o[i + 1]])
        r.qeq_pre_force()
        res = r.pair_compute(True, True)
        eo, _ = o.energies()
        assert rel(res["pvector"], pvector_from_oracle(eo)) < 1e-7
        assert rel(res["f"], o.forces()) < 1e-5          # charges converged to 1e-8 on both sides
    assert np.diff(off_g).max() > 1200                    # the compressed rows are far longer than the first build's stride


def test_nall_mismatch_is_an_error_not_an_overrun():
    """rxb_set_positions / rxb_pair_compute refuse a caller whose nall differs from the device's atom count."""
    from sw_reaxff_b200.api import RxbError
    cfg = H.static_config(1, 1, 1, qeq=False)
    r = make_rxb()
    r.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], np.zeros(len(cfg["x"])), cfg["owner"])
    r.neigh_build()
    with pytest.raises(RxbError, match="nall"):
        r.set_positions(cfg["x"][:-3])
    r.qeq_pre_force()
    with pytest.raises(RxbError, match="nall"):
        r.pair_compute(True, True, f_out=np.zeros((len(cfg["x"]) + 5, 3)))


def test_inner_skin_fallback_far_list_exact_beyond_the_margin():
    """Adaptive inner skin (rxb_nonbonded.cu): the far-list sweep reads only the inner block of each Verlet row while no
    atom has moved more than (cut_in - far)/2 = 0.25 A since the build, and must fall back to the full rows beyond that.
    Atoms are displaced WITHOUT rebuilding the lists, just below and well above the threshold; far list and H must equal
    the oracle's (which filters a fresh 12.5 A list) in both regimes."""
    cfg = H.static_config(1, 1, 1, perturb=0.05, seed=11, qeq=False)
    n, x0, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    r = make_rxb(1e-6)
    r.set_atoms(n, x0, ty, tg, np.zeros(len(x0)), owner)
    r.neigh_build()
    rng = np.random.default_rng(5)
    for amp in (0.2, 0.9):     # max displacement: 0.2*sqrt(3)/... below 0.25 A per atom norm is enforced explicitly
        d = rng.uniform(-1, 1, size=(n, 3))
        d *= (amp / np.linalg.norm(d, axis=1).max())
        x = x0.copy()
        x[:n] += d
        x[n:] = x[owner] + (x0[n:] - x0[owner])            # ghosts follow their owners
        r.set_positions(x)
        r.qeq_pre_force()
        num_g, idx_g, val_g = r.far()
        off_g, _ = r.neighbors(0)
        o = H.Oracle()
        o.set_atoms(n, x, ty, tg, np.zeros(len(x)))
        o.build_neighbors(12.5 + 2 * amp)                  # every pair within 10 A now is in this list
        o.qeq_init(0.0, 10.0, 1e-6)
        o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
        o.qeq_pre_force(owner)
        offH, numH, colH, valH = o.qeq_H()
        assert np.array_equal(num_g, numH), amp
        for i in range(0, n, 7):
            a = dict(zip(idx_g[off_g[i]:off_g[i] + num_g[i]].tolist(), val_g[off_g[i]:off_g[i] + num_g[i]].tolist()))
            b = dict(zip(colH[offH[i]:offH[i] + numH[i]].tolist(), valH[offH[i]:offH[i] + numH[i]].tolist()))
            assert a.keys() == b.keys(), (amp, i)
            assert max(abs(a[k] - b[k]) for k in b) < 1e-11 * np.abs(valH).max()


def test_packed_and_exact_h_formats_agree():
    """The SpMV streams H as 8-byte packed entries (22-bit column + 42-bit fixed point) by default; the exact 12-byte
    format (fp64 value + int32 column) must give the same far list, H to the quantisation step, the same iteration
    counts and charges far inside the CG tolerance."""
    cfg = H.static_config(2, 1, 1, perturb=0.1, seed=21, qeq=False)
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    out = {}
    for exact in (False, True):
        r = make_rxb(1e-10)
        r.set_h_exact(exact)
        r.set_atoms(n, x, ty, tg, np.zeros(len(x)), owner)
        r.neigh_build()
        mv = r.qeq_pre_force()
        assert r.h_format()["bytes_per_entry"] == (12 if exact else 8)
        out[exact] = dict(mv=mv, q=r.get_charges(), far=r.far(), f=r.pair_compute(True, True)["f"])
    a, b = out[False], out[True]
    assert np.array_equal(a["far"][0], b["far"][0]) and np.array_equal(a["far"][1], b["far"][1])
    assert np.abs(a["far"][2] - b["far"][2]).max() < 2e-12          # half a quantisation step of 2^-38
    assert abs(a["mv"][0] - b["mv"][0]) <= 8 and abs(a["mv"][1] - b["mv"][1]) <= 8   # 1e-10 residual: round-off sensitive
    # both solves stop at a 1e-10 relative residual (different iterates of the same system up to the 2^-39 quantisation):
    # the charges agree to the solver's own accuracy, ~1e-9
    assert np.abs(a["q"] - b["q"]).max() < 5e-9
    assert np.abs(a["f"] - b["f"]).max() < 1e-8 * np.abs(b["f"]).max()


def test_qeq_async_enqueue_settles_in_pair_compute():
    """rxb_qeq_pre_force(NULL): the solve is enqueued without polling and settled inside rxb_pair_compute; first with no
    prediction (polls), then with a prediction from a much EASIER solve (under-prediction -> continued + force replay)."""
    cfg = H.static_config(1, 1, 1, perturb=0.1, seed=3, qeq=False)
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    ref = make_rxb(1e-10)
    ref.set_atoms(n, x, ty, tg, np.zeros(len(x)), owner); ref.neigh_build()
    mv_ref = ref.qeq_pre_force(); q_ref = ref.get_charges(); f_ref = ref.pair_compute(True, True)["f"]
    r = make_rxb(1e-2)                                      # a loose solve first: its iteration count is the prediction
    r.set_atoms(n, x, ty, tg, np.zeros(len(x)), owner); r.neigh_build()
    mv_easy = r.qeq_pre_force()
    assert max(mv_easy) + 4 < max(mv_ref)
    r.fix_qeq(0.0, 10.0, 1e-10)
    r.qeq_set_history(np.zeros((n, 5)), np.zeros((n, 5)))   # same starting guess as the reference handle
    r.qeq_pre_force_async()
    f = r.pair_compute(True, True)["f"]
    assert r.qeq_matvecs() == mv_ref
    assert np.abs(r.get_charges() - q_ref).max() < 1e-12
    assert np.abs(f - f_ref).max() < 1e-9 * np.abs(f_ref).max()


def test_hbond_candidate_set_exact(case):
    """a5: the hydrogen-bond list is never stored on the device; its (H atom, acceptor-type partner within hbond_cut)
    pairs are emitted by the far-list sweep.  As a SET they must equal the oracle's hbond list
    (Init_Forces_noQEq_HB_Full_C semantics, reaxc_forces_sw64.c:787-863)."""
    o, r, n = case["o"], case["r"], case["cfg"]["n"]
    Hindex, hs, he, nbr = o.hbonds()
    want = set()
    for i in range(n):
        h = Hindex[i]
        if h >= 0:
            for p in range(hs[h], he[h]):
                want.add((i, int(nbr[p])))
    got = r.hbond_pairs()
    got_set = set(map(tuple, got.tolist()))
    assert len(got_set) == len(got)              # no duplicates
    assert got_set == want and len(want) > 0


def test_workspace_columns_not_materialised_are_covered_by_their_consumers(case):
    """rxb_get_workspace leaves columns 5 (Delta_e), 10 (Clp), 12 (nlp_temp) and 14 (dDelta_lp_temp) at zero: the fused
    multi-body kernel keeps them in registers (they are one-line functions of the stored columns, reaxc_multi_body_sw64.c:
    60-120).  What consumes them - e_lp, e_ov, e_un and the CdDelta they feed - is compared with the oracle here at 1e-9."""
    o, res = case["o"], case["res"]
    eo, _ = o.energies()
    assert abs(res["pvector"][1] - (eo[1] + eo[2])) <= 1e-9 * max(abs(eo[1] + eo[2]), 1.0)     # e_ov + e_un
    assert abs(res["pvector"][2] - eo[3]) <= 1e-9 * max(abs(eo[3]), 1.0)                       # e_lp
    wg = case["r"].workspace()
    assert rel(wg[:, 15], o.cddelta()) < 1e-9


def test_fix_qeq_param_file_mode(tmp_path):
    """fix qeq/reax ... <param file> (FixQEqReaxSunway::pertype_parameters, fix_qeq_reax_sunway.cpp:198-245): chi, eta and
    gamma per atom type come from a file instead of the pair style; the QEq (H shielding, diagonal, right-hand side) must
    use them while the pair style keeps its own.  Charges vs the oracle with the same overrides; also through the C++
    host style reading the file, and its error texts."""
    cfg = H.static_config(1, 1, 1, perturb=0.1, seed=13, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    r = make_rxb(1e-10)
    chi0, eta0, gam0 = r.pair_extract("chi"), r.pair_extract("eta"), r.pair_extract("gamma")
    chi = chi0 * np.array([0, 1.1, 0.9, 1.05, 0.95]); eta = eta0 * np.array([0, 0.97, 1.04, 1.0, 1.02])
    gam = gam0 * np.array([0, 1.03, 0.98, 0.99, 1.01])
    o.set_atoms(n, x, ty, tg, np.zeros(len(x)))
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, 1e-10)
    o.L.orc_qeq_override(o.h, H._c(chi[1:]), H._c(eta[1:]), H._c(gam[1:]))      # LAMMPS types 1..4 = elements 0..3
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
    o.qeq_pre_force(owner)
    r.set_atoms(n, x, ty, tg, None, owner)
    r.neigh_build()
    q_plain = None
    r.qeq_pre_force(); q_plain = r.get_charges()
    r.fix_qeq_params(chi, eta, gam)
    r.qeq_set_history(np.zeros((n, 5)), np.zeros((n, 5)))
    r.qeq_pre_force()
    q = r.get_charges()
    assert np.abs(q - o.q()).max() < 1e-8
    assert np.abs(q - q_plain).max() > 1e-3          # the override really changes the answer
    r.fix_qeq_params(None)                           # back to reax/c
    r.qeq_set_history(np.zeros((n, 5)), np.zeros((n, 5)))
    r.qeq_pre_force()
    assert np.abs(r.get_charges() - q_plain).max() < 1e-9
