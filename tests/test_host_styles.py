"""Host-side C++ styles (sw_reaxff_b200/host): the LAMMPS-facing classes + minimal core stand-in, driven by input
scripts exactly like `lmp -in ... -var S n` in the reference's run.sh.  CPU part: parser/error behaviour.  GPU part:
the host-buffer plugin path reproduces the resident path and the golden step-0 energies."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H

HOSTLIB = os.path.join(H.ROOT, "sw_reaxff_b200", "librxb200_host.so")
SCRIPT = os.path.join(H.DATA, "in.tatb.b200")


def build_host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "sw_reaxff_b200", "host"), "all"])


def run_script(script, **variables):
    build_host()
    L = C.CDLL(HOSTLIB)
    L.rxh_run_script.restype = C.c_long
    names = (C.c_char_p * len(variables))(*[k.encode() for k in variables])
    vals = (C.c_char_p * len(variables))(*[str(v).encode() for v in variables.values()])
    out = np.zeros((4096, 19))
    err = C.create_string_buffer(1024)
    n = L.rxh_run_script(script.encode(), len(variables), names, vals, 0, out.ctypes.data_as(C.c_void_p), C.c_long(4096), err, 1024)
    if n < 0:
        raise RuntimeError(err.value.decode())
    return out[:n]


def lattice_script(tmp_path, t):
    """Same system expressed the way in.reaxc.lattice does it: lattice custom + region prism + create_atoms basis."""
    box6, x, ty, _ = H.read_data_tatb()
    a1 = np.array([box6[0], 0, 0]); a2 = np.array([box6[3], box6[1], 0]); a3 = np.array([box6[4], box6[5], box6[2]])
    frac = np.linalg.solve(np.stack([a1, a2, a3], axis=1), x.T).T
    frac -= np.floor(frac)
    L = ["variable S index 1", "variable t index %d" % t, "units real", "atom_style charge",
         "variable xhi equal $S*%.10f" % box6[0], "variable yhi equal $S*%.10f" % box6[1], "variable zhi equal $S*%.10f" % box6[2],
         "variable xy equal $S*%.11f" % box6[3], "variable xz equal $S*%.11f" % box6[4], "variable yz equal $S*%.11f" % box6[5],
         "lattice custom 1 &", "a1 %.10f 0.0 0.0 &" % a1[0], "a2 %.11f %.10f 0.0 &" % (a2[0], a2[1]), "a3 %.11f %.11f %.10f &" % tuple(a3)]
    L += ["basis %.15f %.15f %.15f &" % tuple(f) for f in frac]
    L += ["", "region 1 prism 0.0 ${xhi} 0.0 ${yhi} 0.0 ${zhi} ${xy} ${xz} ${yz} units box", "create_box 4 1", "create_atoms 4 box &"]
    L += ["basis %d %d &" % (i + 1, t_) for i, t_ in enumerate(ty)]
    L += ["", "mass 1 12.0000", "mass 2 1.0080", "mass 3 15.9990", "mass 4 14.0000",
          "pair_style reax/c %s maxfar 512" % H.CONTROL, "pair_coeff * * %s C H O N" % H.FFIELD,
          "neighbor 2.5 bin", "neigh_modify delay 0 every 5 check no one 1024", "fix 1 all nve",
          "fix 2 all qeq/reax 1 0.0 10.0 1.0e-6 reax/c", "thermo 5", "timestep 0.0625", "run $t"]
    p = tmp_path / "in.lattice"
    p.write_text("\n".join(L) + "\n")
    return str(p)


def test_script_errors_reported(tmp_path):
    p = tmp_path / "in.bad"
    p.write_text("units real\natom_style charge\nfrobnicate 1 2 3\n")
    with pytest.raises(RuntimeError, match="Unknown command: frobnicate"):
        run_script(str(p))
    p.write_text("units lj\n")
    with pytest.raises(RuntimeError, match="units real"):
        run_script(str(p))
    p.write_text("units real\natom_style charge\nread_data %s\nfix 2 all qeq/reax 0 0.0 10.0 1e-6 reax/c\n" % H.DATAFILE)
    with pytest.raises(RuntimeError, match="Illegal fix qeq/reax command"):
        run_script(str(p))
    p.write_text("units real\natom_style charge\nread_data %s\nfix 1 all nve\nrun 1\n" % H.DATAFILE)
    with pytest.raises(RuntimeError, match="No pair style defined"):
        run_script(str(p))


def test_analysis_fix_argument_errors(tmp_path):
    """Argument validation of fix reax/c/bonds / fix reax/c/species (fix_reaxc_bonds_sunway.cpp:48-58,
    fix_reaxc_species_sunway.cpp:54-79, 193-227) happens before any GPU work."""
    head = "units real\natom_style charge\nread_data %s\n" % H.DATAFILE
    p = tmp_path / "in.bad"
    out = tmp_path / "o.txt"
    for line, msg in [("fix 3 all reax/c/bonds 0 %s" % out, "Illegal fix reax/c/bonds command"),
                      ("fix 3 all reax/c/bonds 5", "Illegal fix reax/c/bonds command"),
                      ("fix 3 all reax/c/bonds 5 %s.gz" % out, "Cannot open gzipped file"),
                      ("fix 3 all reax/c/bonds 5 /nonexistent_dir/x", "Cannot open fix reax/c/bonds file"),
                      ("fix 4 all reax/c/species 1 25 20 %s" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 2 5 25 %s" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s cutoff 1 9 0.5" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s cutoff 1 2 1.5" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s bogus" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s position 5" % out, "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 10 %s position 5 %s.pos" % (out, out), "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s position 7 %s.pos" % (out, out), "Illegal fix reax/c/species command"),
                      ("fix 4 all reax/c/species 1 5 5 %s position 5 /nonexistent_dir/p" % out,
                       "Cannot open fix reax/c/species position file")]:
        p.write_text(head + line + "\n")
        with pytest.raises(RuntimeError, match=msg):
            run_script(str(p))
    p.write_text(head + "fix 1 all nve\nfix 3 all reax/c/bonds 5 %s\nrun 1\n" % out)
    with pytest.raises(RuntimeError, match="No pair style defined"):
        run_script(str(p))


@pytest.mark.gpu
def test_script_with_bonds_and_species_fixes_matches_oracle(tmp_path):
    """The reference's C5-style input (fix reax/c/bonds N file + fix reax/c/species nevery nrepeat nfreq file) through the
    plugin path: both output files are byte-identical to the oracle's restatement of the reference writers."""
    fb, fs, fpos = tmp_path / "bonds.out", tmp_path / "species.out", tmp_path / "species.pos"
    lines = ["units real", "atom_style charge", "read_data %s" % H.DATAFILE,
             "pair_style reax/c %s" % H.CONTROL, "pair_coeff * * %s C H O N" % H.FFIELD,
             "neighbor 2.5 bin", "neigh_modify delay 0 every 5 check no", "fix 1 all nve",
             "fix 2 all qeq/reax 1 0.0 10.0 1.0e-10 reax/c", "fix 3 all reax/c/bonds 4 %s" % fb,
             "fix 4 all reax/c/species 1 4 4 %s position 4 %s" % (fs, fpos), "velocity all create 1500.0 4928459", "thermo 4",
             "timestep 0.0625", "run 8"]
    p = tmp_path / "in.c5"
    p.write_text("\n".join(lines) + "\n")
    run_script(str(p))
    # the same run on the oracle (velocities from the driver's own `velocity all create` generator)
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    v = host_velocities(1500.0, 4928459)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-10, every=4)    # the species fix resets reneighbouring to every 4
    o.md_species_init(1, 4, 4)
    bonds, species, pos = o.md_bonds_text(0), "", ""                   # setup(): end_of_step() / post_integrate()
    box6 = np.array([0.0, 0.0, 0.0, box[0], box[1], box[2]])
    o.md_species_step(0)
    for step in range(1, 9):
        if o.md_species_step(step):          # post_integrate of this step: reads the previous step's bond list
            species += o.md_species_text(step)
            pos += o.md_species_pos(step, box6)[0]
        o.md_run(1)
        if step % 4 == 0:
            bonds += o.md_bonds_text(step)   # end_of_step
    assert fb.read_text() == bonds
    assert fs.read_text() == species and species.count("# Timestep") == 2
    # `position 4 file` (WritePos): molecule ids, atom counts and formulas exactly, the printed averages to the last digit
    # or one unit of it (%.8f of quantities the two trajectories agree on to ~1e-10)
    got, want = fpos.read_text().splitlines(), pos.splitlines()
    assert len(got) == len(want) and sum(l.startswith("Timestep") for l in got) == 2
    for a, b in zip(got, want):
        fa, fb_ = a.split("\t"), b.split("\t")
        if a.startswith(("Timestep", "ID", "#")):
            assert a == b
            continue
        assert fa[:3] == fb_[:3] and len(fa) == len(fb_) == 7
        assert np.allclose([float(v) for v in fa[3:]], [float(v) for v in fb_[3:]], rtol=0, atol=2.1e-8), (a, b)


def host_velocities(T, seed):
    L = C.CDLL(HOSTLIB)
    L.rxh_velocities.restype = C.c_long
    v = np.zeros((384, 3))
    n = L.rxh_velocities(H.DATAFILE.encode(), C.c_double(T), C.c_long(seed), v.ctypes.data_as(C.c_void_p), C.c_long(384))
    assert n == 384
    return v


@pytest.mark.gpu
def test_plugin_run_matches_resident_run_and_golden():
    th = run_script(SCRIPT, S=1, t=10, T=0, D=H.DATA)
    assert [int(r[0]) for r in th] == [0, 5, 10]
    g = np.load(os.path.join(H.ROOT, "tests", "golden", "tatb_1x1x1.npz"))
    assert abs(th[0, 2] - g["energies"].sum()) < 1e-6 * abs(g["energies"].sum())      # step-0 PotEng -44760.998
    from sw_reaxff_b200 import Rxb
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    r = Rxb(0)
    r.pair_settings(H.CONTROL); r.pair_coeff(H.FFIELD, H.ELEMENTS); r.fix_qeq(0.0, 10.0, 1e-6)
    r.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS, dt=0.0625, every=5, thermo=5)
    r.md_run(10)
    res = r.md_thermo()
    assert abs(th[2, 2] - res["pe"]) < 1e-7 * abs(res["pe"])
    assert abs(th[2, 3] - res["ke"]) < 1e-4 * max(res["ke"], 1e-3)
    np.testing.assert_allclose(th[2, 5:], res["pvector"], rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
def test_reference_style_lattice_script(tmp_path):
    a = run_script(lattice_script(tmp_path, 5), S=1)
    b = run_script(SCRIPT, S=1, t=5, T=0, D=H.DATA)
    assert abs(a[0, 2] - b[0, 2]) < 1e-7 * abs(b[0, 2])       # same crystal from lattice custom + create_atoms basis
    a2 = run_script(lattice_script(tmp_path, 0), S=2)
    assert abs(a2[0, 2] - 8 * b[0, 2]) < 1e-7 * abs(8 * b[0, 2])


@pytest.mark.gpu
def test_two_runs_in_one_script_reupload_after_setup(tmp_path):
    """`run 8` then `run 8` with `every 5`: LAMMPS::setup() of the second run re-does remap + borders WITHOUT advancing
    the timestep (new ghost set, new index space).  The pair style must re-upload atoms and rebuild the lists then; the
    thermo of the split run equals the thermo of one `run 16` wherever both print (styles_b200.cpp)."""
    base = open(SCRIPT).read()
    one = tmp_path / "in.one"; two = tmp_path / "in.two"
    one.write_text(base.replace("run             $t", "run             16"))
    two.write_text(base.replace("run             $t", "run             8\nrun             8"))
    a = run_script(str(one), S=1, T=1500.0, D=H.DATA, dt=0.25)
    b = run_script(str(two), S=1, T=1500.0, D=H.DATA, dt=0.25)
    sa = {int(r[0]): r for r in a}; sb = {int(r[0]): r for r in b}
    assert 16 in sa and 16 in sb and 8 in sb
    for step in (5, 10, 15, 16):
        if step in sa and step in sb:
            assert abs(sa[step][2] - sb[step][2]) < 1e-7 * abs(sa[step][2]), (step, sa[step][2], sb[step][2])
            # the second run's setup() re-solves QEq from the extrapolated history to the same 1e-6 tolerance: charges, hence
            # e_ele / e_pol, agree to the solver tolerance, not to round-off
            np.testing.assert_allclose(sa[step][5:], sb[step][5:], rtol=5e-5, atol=1e-5)


def run_script_compute(script, cid, **variables):
    build_host()
    L = C.CDLL(HOSTLIB)
    L.rxh_run_script_compute.restype = C.c_long
    names = (C.c_char_p * len(variables))(*[k.encode() for k in variables])
    vals = (C.c_char_p * len(variables))(*[str(v).encode() for v in variables.values()])
    out = np.zeros(400000)
    ncols = C.c_int()
    err = C.create_string_buffer(1024)
    n = L.rxh_run_script_compute(script.encode(), len(variables), names, vals, 0, cid.encode(), out.ctypes.data_as(C.c_void_p),
                                 C.c_long(out.size), C.byref(ncols), err, 1024)
    if n < 0:
        raise RuntimeError(err.value.decode())
    return out[:n].reshape(-1, ncols.value)


def test_compute_spec_atom_argument_errors(tmp_path):
    """compute SPEC/ATOM argument validation (compute_spec_atom_sunway.cpp:38, 121) happens before any GPU work."""
    head = "units real\natom_style charge\nread_data %s\n" % H.DATAFILE
    p = tmp_path / "in.bad"
    for line, msg in [("compute 1 all SPEC/ATOM", "Illegal compute reax/c/atom command"),
                      ("compute 1 all SPEC/ATOM q abo25", "Invalid keyword in compute reax/c/atom command"),
                      ("compute 1 all SPEC/ATOM q fx", "Invalid keyword in compute reax/c/atom command")]:
        p.write_text(head + line + "\n")
        with pytest.raises(RuntimeError, match=msg):
            run_script(str(p))


@pytest.mark.gpu
def test_compute_spec_atom_standalone(tmp_path):
    """compute ID all SPEC/ATOM q x y z vx abo01 ... as a style of its own (compute_spec_atom_sunway.cpp:35-170): the abo
    columns are FindBond's tmpbo (pair_reaxc_sunway.cpp:1170-1198: bonds to partners of higher index with BO >= 0.10, in
    bond-row order), q/x/v are the atom arrays."""
    base = open(SCRIPT).read()
    names = ["q", "x", "y", "z", "vx"] + ["abo%02d" % k for k in range(1, 13)]
    s = tmp_path / "in.compute"
    s.write_text(base.replace("thermo          5", "compute         sa all SPEC/ATOM %s\nthermo          5" % " ".join(names)))
    a = run_script_compute(str(s), "sa", S=1, t=0, T=0, D=H.DATA)
    assert a.shape == (384, len(names))
    from sw_reaxff_b200 import Rxb
    box, x, t, tag = H.tatb_cell(1, 1, 1)
    r = Rxb(0)
    r.pair_settings(H.CONTROL); r.pair_coeff(H.FFIELD, H.ELEMENTS); r.fix_qeq(0.0, 10.0, 1e-6)
    r.md_setup(box, x, np.zeros_like(x), t, tag, H.MASS, dt=0.0625, every=5, thermo=5)
    g = r.md_get()
    assert np.abs(a[:, 0] - g["q"]).max() < 1e-6, np.abs(a[:, 0] - g["q"]).max()     # two solves to the same 1e-6 tolerance
    Hm = np.array([[box[0], box[3], box[4]], [0, box[1], box[5]], [0, 0, box[2]]])
    lam = np.linalg.solve(Hm, (a[:, 1:4] - g["x"]).T).T                               # equal up to a box vector (remap)
    assert np.abs(lam - np.round(lam)).max() < 1e-12, np.abs(lam - np.round(lam)).max()
    assert np.abs(a[:, 4]).max() == 0.0
    abo = r.spec_atom_abo()
    # (the slot order of a row follows the neighbour INDEX order, and the LAMMPS stand-in numbers its ghosts differently from
    # the resident run: compare the rows as multisets)
    assert np.abs(np.sort(a[:, 5:], axis=1) - np.sort(abo, axis=1)).max() < 1e-9
    bs, bc, nbr, _, fld = r.bonds()
    for i in range(384):                                   # FindBond restated
        want = [fld[p, 4] for p in range(bs[i], bs[i] + bc[i]) if nbr[p] >= i and fld[p, 4] >= 0.10]
        assert np.allclose(abo[i, :len(want)], want, rtol=0, atol=1e-15) and np.all(abo[i, len(want):] == 0.0)
    assert (abo > 0).sum() > 300
