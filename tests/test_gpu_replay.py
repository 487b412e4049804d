"""GPU tests (-m gpu) of SURVEY.md §8 row a12: overflow -> grow -> replay of the force phase instead of the reference's
MPI_Abort(INSUFFICIENT_MEMORY) (reaxc_reset_tools_sunway.cpp:122-212, reaxc_forces_sw64.c:866-930), and the absence of fixed
per-atom limits (the reference: 35 bonds per atom while building, MAX_BOND 20 afterwards).

Every growable capacity is shrunk through rxb_debug_set_caps so that each replay branch of System::compute runs on an
ordinary cell; the answer must equal the one of an untouched handle bit for bit in the lists and to round-off in forces
(only the order of the atomic additions differs).  Then cells dense enough to exceed the default per-atom staging
(64 bonds per atom, 32 strong bonds per centre) are checked against the oracle, which has no such limits.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def make_rxb(tol=1e-6):
    from sw_reaxff_b200 import Rxb
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


def evaluate(cfg, q, caps=None, plugin_overlap=True):
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    r = make_rxb()
    r.set_atoms(n, x, ty, tg, q, owner)
    r.neigh_build()
    if caps:
        r.debug_set_caps(**caps)
    if plugin_overlap:
        r.qeq_pre_force()          # starts the bonded chain on the second stream: the replay is entered from overlapped_back
        r.set_charges(q)
    res = r.pair_compute(True, True)
    res["caps"] = r.debug_get_caps()
    res["bonds"] = r.bonds()
    res["counts"] = r.counts()
    return res


@pytest.fixture(scope="module")
def cell():
    cfg = H.static_config(1, 1, 1, perturb=0.1, seed=5, scale=0.95, qeq=True)
    return cfg, cfg["q"].copy()


@pytest.mark.parametrize("caps", [
    dict(cap_bonds=500),                                   # directed-bond arrays (overflow bit 2)
    dict(cap_ang=100), dict(cap_tor=100), dict(cap_hb=100),   # work lists of the angle / torsion / hydrogen-bond items
    dict(row_cap=8), dict(strong_cap=2),                   # per-atom shared-memory staging (overflow bits 1 / 8)
    dict(cap_bonds=300, cap_ang=50, cap_tor=50, cap_hb=50, row_cap=8, strong_cap=2),   # everything at once
], ids=["bonds", "angles", "torsions", "hbonds", "bond_row", "strong_list", "all"])
@pytest.mark.parametrize("overlap", [True, False], ids=["two_stream", "sequential"])
def test_every_replay_branch_gives_the_untouched_answer(cell, caps, overlap):
    cfg, q = cell
    ref = evaluate(cfg, q, None, overlap)
    got = evaluate(cfg, q, caps, overlap)
    for k, v in caps.items():
        assert got["caps"][k] > v, (k, got["caps"])        # the capacity really was outgrown and grown
    bs0, bc0, nbr0, sym0, fld0 = ref["bonds"]
    bs1, bc1, nbr1, sym1, fld1 = got["bonds"]
    assert np.array_equal(bc0, bc1)
    for i in range(0, len(bc0), 37):                        # rows are carved from an atomic cursor: compare row contents
        assert np.array_equal(nbr0[bs0[i]:bs0[i] + bc0[i]], nbr1[bs1[i]:bs1[i] + bc1[i]])
    np.testing.assert_allclose(got["pvector"], ref["pvector"], rtol=1e-12, atol=1e-9)
    assert np.abs(got["f"] - ref["f"]).max() < 1e-11 * np.abs(ref["f"]).max()
    np.testing.assert_allclose(got["virial"], ref["virial"], rtol=1e-10, atol=1e-7)


def oracle_forces(cfg, q):
    o = cfg["oracle"]
    o.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], q)
    o.build_neighbors(12.5)
    o.compute()
    e, _ = o.energies()
    return o.forces(), e, o.bonds()


@pytest.mark.parametrize("scale", [0.62, 0.55], ids=["scale0.62", "scale0.55"])
def test_cells_beyond_the_default_per_atom_staging_match_the_oracle(scale):
    """Cells compressed until atoms carry more bonds than the default staging holds (64 per atom / 32 strong per centre):
    the kernels flag it, the host grows the staging and replays, and the result is the oracle's.  (Physically absurd
    densities; the point is that no fixed limit of the reference survives.)"""
    cfg = H.static_config(1, 1, 1, perturb=0.05, seed=9, scale=scale, qeq=False)
    q = np.zeros(len(cfg["x"]))
    fo, eo, (obs, obe, onbr, _, _) = oracle_forces(cfg, q)
    longest = int((obe - obs).max())
    r = make_rxb()
    r.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], q, cfg["owner"])
    r.neigh_build()
    res = r.pair_compute(True, True)
    caps = r.debug_get_caps()
    bs, bc, nbr, _, _ = r.bonds()
    assert int(bc.max()) == longest
    if longest > 64:
        assert caps["row_cap"] >= longest                  # the default staging was outgrown
    assert np.array_equal(bc, obe - obs)
    for i in range(0, len(bc), 53):
        assert np.array_equal(nbr[bs[i]:bs[i] + bc[i]], onbr[obs[i]:obe[i]])
    if np.isfinite(fo).all():
        # (at 0.55 of the lattice constant - 67 bonds on one atom - the reference formulas themselves overflow to NaN in
        # the oracle as well; there only the lists are compared)
        assert np.isfinite(res["f"]).all()
        assert np.abs(res["f"] - fo).max() < 1e-8 * np.abs(fo).max()
        assert abs(res["eng"].sum() - eo.sum()) < 1e-8 * abs(eo.sum())
    else:
        assert longest > 64


def test_hot_dense_run_never_hits_a_limit():
    """0.80-scale cell at 4000 K, 50 resident steps: bonds break and form, lists are rebuilt every 5
    steps, nothing reports `capacity exceeded`; the trajectory equals the oracle's."""
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=0.80)
    v = H.maxwell_velocities(t, 4000.0, 77)
    r = make_rxb(1e-8)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.1, every=5, thermo=5)
    r.md_run(50)
    g = r.md_get()
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.1, qeq_tol=1e-8)
    o.md_run(50)
    ref = o.md_get()
    dx = np.abs(g["x"] - ref["x"]).max()
    assert np.isfinite(g["x"]).all() and dx < 1e-5, dx      # 50 chaotic steps at 4000 K: round-off grows, stays << 1e-5 A
