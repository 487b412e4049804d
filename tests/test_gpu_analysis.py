"""fix reax/c/bonds and fix reax/c/species on the GPU vs the CPU oracle: byte-identical output text."""
import numpy as np
import pytest

import helpers as H
from sw_reaxff_b200 import Rxb, analysis

pytestmark = pytest.mark.gpu


def make_rxb(tol):
    r = Rxb(0)
    r.pair_settings(H.CONTROL)
    r.pair_coeff(H.FFIELD, H.ELEMENTS)
    r.fix_qeq(0.0, 10.0, tol)
    return r


def canonical(text):
    """Rows of the connection table with their (id, bo) pairs sorted by neighbour ID.  The reference orders a row by the
    neighbour's LOCAL index, i.e. by where LAMMPS' comm happened to place ghost images; the resident run builds its own
    ghosts, so the order of two images inside a row is the one thing that may legitimately differ.
    The plugin path, where the host supplies the ghosts, is compared byte for byte in test_host_styles.py."""
    out = []
    for ln in text.splitlines():
        if ln.startswith("#"):
            out.append(ln); continue
        w = ln.split()
        nb = int(w[2])
        pairs = sorted(zip((int(a) for a in w[3:3 + nb]), w[4 + nb:4 + 2 * nb]))
        out.append(" ".join(w[:3] + ["%d:%s" % p for p in pairs] + w[4 + 2 * nb:]))
    return out


@pytest.mark.parametrize("scale,T,steps", [(1.0, 300.0, 7), (0.90, 3000.0, 12)], ids=["cold", "hot_compressed"])
def test_bond_table_text_identical(scale, T, steps):
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=scale)
    v = H.maxwell_velocities(t, T, 2024)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-10)
    r = make_rxb(1e-10)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=1)
    for step in (0, steps):
        if step:
            o.md_run(step); r.md_run(step)
        ref = o.md_bonds_text(step)
        tb = r.bond_table()
        got = analysis.bonds_text(tb, step, len(x), 0.3)
        assert tb["off"][-1] == len(tb["nbr"]) and tb["max_nb"] == np.diff(tb["off"]).max()
        assert canonical(got) == canonical(ref)
        if scale == 1.0:
            assert got == ref


def test_species_single_sample_matches_oracle():
    box, x, t, tag = H.tatb_cell(2, 1, 1)
    v = H.maxwell_velocities(t, 300.0, 5)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-8)
    o.md_species_init(1, 1, 1)
    r = make_rxb(1e-8)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=5, thermo=1)
    assert r.species_config(1, 1, 1, natoms=len(x)) is False
    assert not o.md_species_step(0) and not r.species_step(0)
    assert o.md_species_step(1); o.md_run(1)        # post_integrate of step 1 reads the bond list of step 0
    r.md_run(1)                                     # md_run calls the post_integrate hook itself
    log = r.species_log()
    assert [rec["step"] for rec in log] == [1]
    so = o.md_species_get()
    assert log[0]["nmole"] == so["nmole"] == 32
    assert np.array_equal(log[0]["composition"], so["composition"])
    assert np.array_equal(r.species_cluster(), so["cluster"])
    assert analysis.species_text(1, log[0]["composition"]) == o.md_species_text(1)


def test_species_averaged_hot_compressed_matches_oracle():
    """nevery 1, nrepeat 5, nfreq 5 on the reacting system: bond orders averaged slot by slot over 5 steps with the lists
    frozen (the fix resets reneighbouring to every 5), molecules of many different compositions."""
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=0.80)
    v = H.maxwell_velocities(t, 4000.0, 99)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-10, every=5)
    bc = np.full((5, 5), 0.30); bc[1, 4] = bc[4, 1] = 0.9       # `cutoff 1 4 0.9`: C-N bonds below 0.9 do not count
    o.md_species_init(1, 5, 5, bocut=bc)
    r = make_rxb(1e-10)
    r.md_setup(box, x, v, t, tag, H.MASS, dt=0.0625, every=10, thermo=1)
    assert r.species_config(1, 5, 5, natoms=len(x), bocut=bc) is True      # every 10 -> 5: "Resetting reneighboring criteria"
    o.md_species_step(0); r.species_step(0)
    outs = []
    for step in range(1, 11):
        if o.md_species_step(step):      # post_integrate: before this step's force evaluation
            outs.append((step, o.md_species_get(), o.md_species_text(step)))
        o.md_run(1)
    r.md_run(10)
    log = r.species_log()
    assert [rec["step"] for rec in log] == [s for s, _, _ in outs] == [5, 10]
    for rec, (step, so, txt) in zip(log, outs):
        assert rec["nmole"] == so["nmole"]
        assert np.array_equal(rec["composition"], so["composition"])
        assert analysis.species_text(step, rec["composition"]) == txt
    assert len(analysis.find_species(log[-1]["composition"])[0]) >= 4    # several fragment species
