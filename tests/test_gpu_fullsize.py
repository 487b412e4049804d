"""Full-size (BASELINE.json configs[1]: TATB 8x8x8, 196,608 atoms) checks through size-independent properties.

The oracle cannot finish 196,608 atoms in seconds, so the large system is checked against the SMALL system's oracle
answer through exact physical invariances of the periodic crystal:
  * replication invariance: every per-term energy of the 8x8x8 replica = 512 x the unit cell's; forces, charges, bond
    counts and bond orders of atom i equal those of its unit-cell image;
  * Newton's third law (zero net force after reverse_comm), charge neutrality, symmetric bond orders.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def make_rxb(tol):
    from sw_reaxff_b200 import Rxb
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


@pytest.fixture(scope="module")
def unit_cell_oracle():
    cfg = H.static_config(1, 1, 1, qeq=False)
    o = cfg["oracle"]
    n, x, ty, tg, owner = cfg["n"], cfg["x"], cfg["type"], cfg["tag"], cfg["owner"]
    o.set_atoms(n, x, ty, tg, np.zeros(len(x)))
    o.build_neighbors(12.5)
    o.qeq_init(0.0, 10.0, 1e-12)
    o.qeq_set_hist(np.zeros((n, 5)), np.zeros((n, 5)))
    o.qeq_pre_force(owner)
    o.compute()
    f = o.forces()
    fl = f[:n].copy()
    np.add.at(fl, owner, f[n:])
    bs, be, *_ = o.bonds()
    return dict(e=o.energies()[0], f=fl, q=o.q()[:n], nb=(be - bs)[:n], tbo=o.workspace()[:n, 0])


def test_8x8x8_replication_invariance(unit_cell_oracle):
    u = unit_cell_oracle
    nrep = 8
    box, x, t, tag = H.tatb_cell(nrep, nrep, nrep)
    assert len(x) == 196608
    r = make_rxb(1e-12)
    r.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS, thermo=1)
    out = r.md_get()
    th = r.md_thermo()
    ncell = nrep ** 3
    e = u["e"]
    pv = np.array([e[0], e[1] + e[2], e[3], 0.0, e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11], 0.0, e[12]])
    np.testing.assert_allclose(th["pvector"], ncell * pv, rtol=1e-8, atol=1e-6)
    f = out["f"].reshape(ncell, 384, 3)
    fmax = np.abs(u["f"]).max()
    assert np.abs(f - u["f"][None]).max() < 1e-8 * fmax
    q = out["q"].reshape(ncell, 384)
    assert np.abs(q - u["q"][None]).max() < 1e-9
    assert abs(out["q"].sum()) < 1e-8
    assert np.abs(out["f"].sum(0)).max() < 1e-7 * fmax
    bs, bc, nbr, sym, fld = r.bonds()
    assert np.array_equal(bc[:196608].reshape(ncell, 384), np.broadcast_to(u["nb"], (ncell, 384)))
    w = r.workspace()
    assert np.abs(w[:196608, 0].reshape(ncell, 384) - u["tbo"][None]).max() < 1e-10
    assert np.abs(fld[:, 4] - fld[sym, 4]).max() < 1e-12     # BO_ij == BO_ji over 2.6 M directed bonds


def test_8x8x8_translation_and_reneighbour_idempotence():
    """Shifting every atom by a lattice-incommensurate vector (atoms re-wrap, ghosts and lists change) changes nothing."""
    box, x, t, tag = H.tatb_cell(4, 4, 4)
    r1 = make_rxb(1e-12)
    r1.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS, thermo=1)
    a = r1.md_get(); ta = r1.md_thermo()
    r2 = make_rxb(1e-12)
    r2.md_setup(box, x + np.array([3.217, -7.31, 11.09]), np.zeros_like(x), t, tag, H.MASS, thermo=1)
    b = r2.md_get(); tb = r2.md_thermo()
    assert abs(ta["pe"] - tb["pe"]) < 1e-9 * abs(ta["pe"])
    assert np.abs(a["f"] - b["f"]).max() < 1e-8 * np.abs(a["f"]).max()
    assert np.abs(a["q"] - b["q"]).max() < 1e-9
