"""Pins the oracle (and the product's parser) against the REFERENCE ITSELF where the reference compiles here:
oracle/_ref/libref.so = the reference's unmodified reaxc_ffield / reaxc_control / reaxc_tool_box sources compiled from
/root/reference with stub headers (oracle/ref/Makefile).  The library is prebuilt in this container and travels to the
GPU box; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from test_host_cpu import parse_dump

LIBREF = os.path.join(H.ROOT, "oracle", "_ref", "libref.so")


def ref_dump(control, ffield, elements, lgflag=0):
    L = C.CDLL(LIBREF)
    L.ref_params_dump.restype = C.c_long
    arr = (C.c_char_p * len(elements))(*[e.encode() for e in elements])
    n = L.ref_params_dump(control.encode() if control else None, ffield.encode(), len(elements), arr, lgflag, None, C.c_long(0))
    assert n > 0
    out = np.zeros(n)
    L.ref_params_dump(control.encode() if control else None, ffield.encode(), len(elements), arr, lgflag,
                      out.ctypes.data_as(C.c_void_p), C.c_long(n))
    return out


def mask_taper(d):
    """Tap[8] is produced by Init_Taper (reaxc_init_md), not by the parsers: the _ref dump leaves those 8 slots zero."""
    d = d.copy()
    ngp = int(d[2])
    d[3 + ngp + 10:3 + ngp + 18] = 0.0
    return d


pytestmark = pytest.mark.skipif(not os.path.exists(LIBREF), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.mark.parametrize("control,elements", [(H.CONTROL, H.ELEMENTS), (None, ["N", "O", "H", "C"]), (H.CONTROL, ["C", "H", "NULL", "N"])])
def test_oracle_and_product_parsers_equal_reference(control, elements):
    ref = ref_dump(control, H.FFIELD, elements)
    orc = H.Oracle(control=control, elements=elements).params_dump()
    mine = parse_dump(control, H.FFIELD, elements)
    assert ref.shape == orc.shape == mine.shape
    assert np.array_equal(mask_taper(orc), ref), np.nonzero(mask_taper(orc) != ref)[0][:10]
    assert np.array_equal(mask_taper(mine), ref), np.nonzero(mask_taper(mine) != ref)[0][:10]


def test_reference_parser_on_modified_force_field(tmp_path):
    """A force field with reordered torsion entries (explicit before wildcard and vice versa) and a duplicated angle
    line exercises the tor_flag / cnt++ corner cases of reaxc_ffield_sunway.cpp:541-543,585-646."""
    src = open(H.FFIELD).read().splitlines()
    # locate the torsion block: line that starts the count "17    ! Nr of torsions"
    it = next(i for i, l in enumerate(src) if "torsion" in l.lower() and l.split()[0].isdigit())
    nt = int(src[it].split()[0])
    block = src[it + 1:it + 1 + nt]
    src[it + 1:it + 1 + nt] = block[::-1]
    ia = next(i for i, l in enumerate(src) if "angle" in l.lower() and l.split()[0].isdigit())
    na = int(src[ia].split()[0])
    src[ia] = src[ia].replace(str(na), str(na + 1), 1)
    src.insert(ia + 1, src[ia + 1])
    p = tmp_path / "ffield.mod"
    p.write_text("\n".join(src) + "\n")
    ref = ref_dump(H.CONTROL, str(p), H.ELEMENTS)
    orc = H.Oracle(ffield=str(p)).params_dump()
    mine = parse_dump(H.CONTROL, str(p), H.ELEMENTS)
    assert np.array_equal(mask_taper(orc), ref)
    assert np.array_equal(mask_taper(mine), ref)


# ---------------------------------------------------------------------------------------------------------------------
# The reference's LIVE serial energy routines (BOp_single, Torsion_Angles with control->virial = 1, Hydrogen_Bonds with
# virial = 1, Add_dBond_to_Forces) compiled unmodified into oracle/_ref and driven on the oracle's own intermediate
# state (oracle/ref/ref_bonded.cpp).  This pins the oracle's bond-order, valence-angle, torsion, hydrogen-bond and
# bond-order chain-rule restatements against the reference itself, term by term.
def _ref_lib(ffield=None):
    L = C.CDLL(LIBREF)
    L.ref_load.restype = C.c_void_p
    return L, C.c_void_p(L.ref_load(H.CONTROL.encode(), (ffield or H.FFIELD).encode()))


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _state(perturb, seed, scale=1.0):
    """Oracle after bond list + hbond list + bond orders (phases 0,1,2,4) on a perturbed TATB cell with ghosts."""
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, scale=scale, qeq=False)
    o = cfg["oracle"]
    o.set_atoms(cfg["n"], cfg["x"], cfg["type"], cfg["tag"], np.zeros(len(cfg["x"])))
    o.build_neighbors(12.5)
    for ph in (0, 1, 2, 4):
        o.phase(ph)
    return cfg, o


def _call_ref(L, P, which, cfg, o, Cd_in=None, CdDelta_in=None):
    n, x = cfg["n"], cfg["x"]
    N = len(x)
    etype = _ip([{1: 0, 2: 1, 3: 2, 4: 3}[t] for t in cfg["type"]])      # pair_coeff * * ffield C H O N: type k -> element k-1
    bs, be, nbr, sym, fld = o.bonds()
    nb = len(nbr)
    w = o.workspace(); dd = o.ddeltap_self()
    Hindex, hs, he, hnbr = o.hbonds()
    numH = int((Hindex[:n] >= 0).sum())
    hrow = np.full(len(hnbr), -1, dtype=np.int64)
    for j in range(n):
        if Hindex[j] >= 0:
            hrow[hs[Hindex[j]]:he[Hindex[j]]] = j
    hvec = x[hnbr] - x[hrow]
    hd = np.sqrt((hvec ** 2).sum(1))
    en = np.zeros(6); fcd = np.zeros((N, 4)); cd = np.zeros((3, nb))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(a) for a in (x, fld, w, dd, hd, hvec)]
    ints = [_ip(a) for a in (cfg["tag"], bs, be, nbr, sym, Hindex, hs[:numH], he[:numH], hnbr)]
    Cd = None if Cd_in is None else np.ascontiguousarray(Cd_in)
    CdD = None if CdDelta_in is None else np.ascontiguousarray(CdDelta_in)
    rc = L.ref_bonded(P, which, n, N, p(keep[0]), p(etype), p(ints[0]), p(ints[1]), p(ints[2]), nb, p(ints[3]), p(ints[4]),
                      p(keep[1]), p(keep[2]), p(keep[3]), p(ints[5]), numH, p(ints[6]), p(ints[7]), len(hnbr), p(ints[8]),
                      p(keep[4]), p(keep[5]), None if Cd is None else p(Cd), None if CdD is None else p(CdD), p(en), p(fcd), p(cd))
    assert rc == 0
    return en, fcd, cd


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("perturb,seed,scale", [(0.0, 0, 1.0), (0.1, 3, 1.0), (0.1, 4, 0.9)])
def test_uncorrected_bond_orders_equal_reference_BOp_single(perturb, seed, scale):
    """a4: every pair within bond_cut goes through the reference's BOp_single; the accepted set and BO', BO_s, BO_pi,
    BO_pi2, dBOp, dln_BOp_pi, dln_BOp_pi2 of the oracle's bond list must match."""
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, scale=scale, qeq=False)
    o = cfg["oracle"]
    x = cfg["x"]
    o.set_atoms(cfg["n"], x, cfg["type"], cfg["tag"], np.zeros(len(x)))
    o.build_neighbors(12.5)
    o.phase(0); o.phase(1)
    bs, be, nbr, sym, fld = o.bonds()
    off, idx = o.get_neighbors()
    N = len(x)
    rows = np.repeat(np.arange(N), np.diff(off))
    dv = x[idx] - x[rows]
    d = np.sqrt((dv ** 2).sum(1))
    sel = d <= 4.5                                    # control bond_cut
    rows, cols, dv, d = rows[sel], idx[sel], np.ascontiguousarray(dv[sel]), np.ascontiguousarray(d[sel])
    et = np.array([{1: 0, 2: 1, 3: 2, 4: 3}[t] for t in cfg["type"]], dtype=np.int32)
    ti, tj = _ip(et[rows]), _ip(et[cols])
    L, P = _ref_lib()
    out = np.zeros((len(d), 15))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.ref_bop_pairs(P, len(d), p(ti), p(tj), p(d), p(dv), p(out))
    acc = out[:, 0] > 0
    ref_pairs = set(zip(rows[acc].tolist(), cols[acc].tolist()))
    brow = np.repeat(np.arange(N), be - bs) if np.array_equal(bs[1:], be[:-1]) else None
    orc_pairs = set()
    by_pair = {}
    for i in range(N):
        for pp in range(bs[i], be[i]):
            orc_pairs.add((i, int(nbr[pp]))); by_pair[(i, int(nbr[pp]))] = pp
    assert orc_pairs == ref_pairs and len(orc_pairs) > 1000
    k = np.nonzero(acc)[0]
    pp = np.array([by_pair[(int(rows[a]), int(cols[a]))] for a in k])
    ref = out[k]
    assert relerr(fld[pp, 4], ref[:, 1]) < 1e-13      # BO' - bo_cut
    assert relerr(fld[pp, 5], ref[:, 2]) < 1e-13      # BO_s
    assert relerr(fld[pp, 6], ref[:, 3]) < 1e-13 and relerr(fld[pp, 7], ref[:, 4]) < 1e-13
    assert relerr(fld[pp, 8:11], ref[:, 5:8]) < 1e-12     # dBOp
    assert relerr(fld[pp, 11:14], ref[:, 8:11]) < 1e-12 and relerr(fld[pp, 14:17], ref[:, 11:14]) < 1e-12


@pytest.mark.parametrize("perturb,seed,scale", [(0.0, 0, 1.0), (0.1, 1, 1.0), (0.1, 2, 0.9)])
def test_valence_torsion_hbond_dbond_equal_reference_serial_routines(perturb, seed, scale):
    L, P = _ref_lib()
    cfg, o = _state(perturb, seed, scale)
    n = cfg["n"]
    # --- valence angles + torsions (a10): reference Torsion_Angles, serial virial path
    en, fcd, cd = _call_ref(L, P, 1, cfg, o)
    o.phase(0); o.phase(7)
    e, _ = o.energies()
    bs, be, nbr, sym, fld = o.bonds()
    for k_ref, k_orc in ((0, 4), (1, 5), (2, 6), (3, 8), (4, 9)):     # e_ang e_pen e_coa e_tor e_con
        assert abs(en[k_ref] - e[k_orc]) <= 1e-10 * max(1.0, abs(e[k_orc])), (k_ref, en[k_ref], e[k_orc])
    assert relerr(-fcd[:, :3], o.forces()) < 1e-10
    assert relerr(fcd[:, 3], o.cddelta()) < 1e-10
    assert relerr(cd.T, fld[:, 28:31]) < 1e-10
    Cd_vt, CdD_vt = fld[:, 28:31].T.copy(), o.cddelta().copy()
    # --- hydrogen bonds (a8), accumulating on top of the valence/torsion coefficients like Compute_Bonded_Forces does
    f_before = o.forces().copy()
    en2, fcd2, cd2 = _call_ref(L, P, 2, cfg, o, Cd_in=Cd_vt, CdDelta_in=CdD_vt)
    o.phase(6)
    e, _ = o.energies()
    assert abs(en2[5] - e[7]) <= 1e-10 * max(1.0, abs(e[7])) and abs(e[7]) > 1.0
    assert relerr(-fcd2[:, :3], o.forces() - f_before) < 1e-10
    bs, be, nbr, sym, fld = o.bonds()
    assert relerr(cd2.T, fld[:, 28:31]) < 1e-10
    # --- bond-order chain rule (a11): reference Add_dBond_to_Forces over every bond once
    Cd_all, CdD_all = fld[:, 28:31].T.copy(), o.cddelta().copy()
    f_before = o.forces().copy()
    en3, fcd3, cd3 = _call_ref(L, P, 4, cfg, o, Cd_in=Cd_all, CdDelta_in=CdD_all)
    o.phase(8)
    df = o.forces() - f_before
    assert np.abs(df).max() > 10.0
    assert relerr(-fcd3[:, :3], df) < 1e-10


def test_npt_branches_same_forces_and_zero_ext_press():
    """control->virial = 1 (PuReMD's NPT switch): Torsion_Angles and Hydrogen_Bonds take their serial branches and
    Add_dBond_to_Forces_NPT replaces Add_dBond_to_Forces; each adds rvec_iMultiply(ext_press, rel_box, force) to
    data->my_ext_press.  Behind the LAMMPS interface rel_box is zero for every neighbour (pair_reaxc_sunway.cpp:927), so the
    branches must give the virial = 0 forces and a zero my_ext_press - which is why the product path has no NPT variant
    (DESIGN.md section 8)."""
    L, P = _ref_lib()
    cfg, o = _state(0.08, 11, 0.95)
    press = np.zeros(3)
    get = lambda: (L.ref_last_ext_press(press.ctypes.data_as(C.c_void_p)), press.copy())[1]
    en, fcd, cd = _call_ref(L, P, 1, cfg, o)                   # valence + torsion, virial = 1 branch
    assert np.abs(fcd[:, :3]).max() > 1.0 and np.all(get() == 0.0)
    o.phase(0); o.phase(7)
    bs, be, nbr, sym, fld = o.bonds()
    Cd_vt, CdD_vt = fld[:, 28:31].T.copy(), o.cddelta().copy()
    en2, fcd2, cd2 = _call_ref(L, P, 2, cfg, o, Cd_in=Cd_vt, CdDelta_in=CdD_vt)     # hydrogen bonds, virial = 1 branch
    assert abs(en2[5]) > 1.0 and np.all(get() == 0.0)
    o.phase(6)
    bs, be, nbr, sym, fld = o.bonds()
    Cd_all, CdD_all = fld[:, 28:31].T.copy(), o.cddelta().copy()
    _, f_plain, _ = _call_ref(L, P, 4, cfg, o, Cd_in=Cd_all, CdDelta_in=CdD_all)    # Add_dBond_to_Forces
    _, f_npt, _ = _call_ref(L, P, 8, cfg, o, Cd_in=Cd_all, CdDelta_in=CdD_all)      # Add_dBond_to_Forces_NPT
    assert np.all(get() == 0.0)
    assert np.abs(f_plain[:, :3]).max() > 10.0
    assert relerr(f_npt[:, :3], f_plain[:, :3]) < 1e-13


def test_taper_equals_reference_Init_Taper():
    L, P = _ref_lib()
    tap = np.zeros(8)
    L.ref_taper(P, tap.ctypes.data_as(C.c_void_p))
    d = H.Oracle().params_dump()
    ngp = int(d[2])
    assert np.array_equal(d[3 + ngp + 10:3 + ngp + 18], tap) and tap[7] != 0.0
    mine = parse_dump(H.CONTROL, H.FFIELD, H.ELEMENTS)
    assert np.array_equal(mine[3 + ngp + 10:3 + ngp + 18], tap)


@pytest.mark.parametrize("perturb,seed,scale", [(0.0, 0, 1.0), (0.1, 5, 1.0), (0.1, 6, 0.9)])
def test_bond_and_atom_energies_equal_reference_MPE_serial_routine(perturb, seed, scale):
    """a7: Merge_Bonds_Atom_Energy_C_New (reaxc_multi_body_sw64.c:21-333), the routine the production run executes."""
    L, P = _ref_lib()
    cfg, o = _state(perturb, seed, scale)
    n, N = cfg["n"], len(cfg["x"])
    q = np.ascontiguousarray(np.random.default_rng(seed).uniform(-0.5, 0.5, N))
    etype = _ip([t - 1 for t in cfg["type"]])
    bs, be, nbr, sym, fld = o.bonds()
    nb = len(nbr)
    w = o.workspace()
    en = np.zeros(5); ev = np.zeros(2); fcd = np.zeros((N, 4)); cd = np.zeros((3, nb))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ints = [_ip(a) for a in (cfg["tag"], bs, be, nbr, sym)]
    fl, wl = np.ascontiguousarray(fld), np.ascontiguousarray(w)
    assert L.ref_atom_energy(P, 1, n, N, p(q), p(etype), p(ints[0]), p(ints[1]), p(ints[2]), nb, p(ints[3]), p(ints[4]), p(fl), p(wl),
                             p(en), p(ev), p(fcd), p(cd)) == 0
    o.phase(0); o.phase(5)
    e, _ = o.energies()
    for k_ref, k_orc in ((0, 0), (1, 3), (2, 1), (3, 2)):       # e_bond, e_lp, e_ov, e_un
        assert abs(en[k_ref] - e[k_orc]) <= 1e-10 * max(1.0, abs(e[k_orc])), (k_ref, en[k_ref], e[k_orc])
    assert abs(e[0]) > 1e3 and abs(e[1]) > 1.0
    bs, be, nbr, sym, fld2 = o.bonds()
    assert relerr(cd.T, fld2[:, 28:31]) < 1e-10
    assert relerr(fcd[:, 3], o.cddelta()) < 1e-10
    assert np.abs(fcd[:, :3]).max() == 0.0        # this routine produces no direct forces


@pytest.mark.parametrize("perturb,seed,vdw_type", [(0.0, 0, 1), (0.1, 7, 1), (0.1, 8, 3), (0.1, 9, 2)])
def test_nonbonded_equals_reference_MPE_serial_routine(perturb, seed, vdw_type, tmp_path):
    """a9: vdW_Coulomb_Energy_Full_C_test_err (reaxc_nonbonded_sw64.c:40-258), full list, owner computes; the shipped
    force field (shielded vdW, type 1) and variants that take the inner-wall branches (types 3 and 2)."""
    ff = H.FFIELD if vdw_type == 1 else H.ffield_variant(tmp_path / "ffield.v", vdw_type)
    L, P = _ref_lib(ff)
    orc = H.Oracle(ffield=ff)
    assert int(orc.params_dump()[1]) == vdw_type
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, qeq=True, oracle=orc)
    o = cfg["oracle"]
    n, x, q = cfg["n"], cfg["x"], cfg["q"]
    N = len(x)
    o.set_atoms(n, x, cfg["type"], cfg["tag"], q)
    o.build_neighbors(12.5)
    off, idx = o.get_neighbors()
    rows = np.repeat(np.arange(N), np.diff(off))
    loc = rows < n
    rows, cols = rows[loc], idx[loc]
    dv = x[cols] - x[rows]
    d = np.sqrt((dv ** 2).sum(1))
    sel = d <= 10.0
    rows, cols, dv, d = rows[sel], _ip(cols[sel]), np.ascontiguousarray(dv[sel]), np.ascontiguousarray(d[sel])
    far_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=far_off[1:])
    tap = np.zeros(8)
    L.ref_taper(P, tap.ctypes.data_as(C.c_void_p))
    etype = _ip([t - 1 for t in cfg["type"]])
    en = np.zeros(2); fcd = np.zeros((N, 4))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    xx, qq, tg = np.ascontiguousarray(x), np.ascontiguousarray(q), _ip(cfg["tag"])
    assert L.ref_nonbonded(P, n, N, p(xx), p(qq), p(etype), p(tg), p(far_off), p(cols), p(d), p(dv), p(tap), p(en), p(fcd)) == 0
    o.phase(0); o.phase(3)
    e, _ = o.energies()
    # the reference adds the FULL pair energy for every directed pair (SURVEY.md §8 "Semantics"): 2x the stock sum
    assert abs(en[0] - 2 * e[10]) <= 1e-10 * abs(2 * e[10]) and abs(en[1] - 2 * e[11]) <= 1e-10 * abs(2 * e[11])
    assert abs(e[10]) > 100 and abs(e[11]) > 100
    assert relerr(-fcd[:n, :3], o.forces()[:n]) < 1e-10
    assert np.abs(o.forces()[n:]).max() == 0.0 and np.abs(fcd[n:]).max() == 0.0     # ghosts receive no nonbonded force


# ---------------------------------------------------------------------------------------------------------------------
# fix qeq/reax: the reference's own FixQEqReaxSunway (fix_qeq_reax_sunway.cpp compiled unmodified against a LAMMPS-core
# stand-in, oracle/ref/ref_qeq.cpp) against the oracle's restatement: init_taper, init_matvec extrapolation, CG_v2 and
# its iteration counts, calculate_Q and the history shift.
def _ref_qeq(cfg, o, tol, s_hist, t_hist, swb=10.0):
    L = C.CDLL(LIBREF)
    n, x = cfg["n"], np.ascontiguousarray(cfg["x"])
    N = len(x)
    d = o.params_dump()
    ngp = int(d[2]); base = 3 + ngp + 18
    el = d[base:base + 4 * 30].reshape(4, 30)
    chi = np.concatenate([[0.0], el[:, 13]]); eta = np.concatenate([[0.0], el[:, 14]]); gam = np.concatenate([[0.0], el[:, 5]])
    off, idx = o.get_neighbors()
    off = np.ascontiguousarray(off[:n + 1], dtype=np.int64); idx = _ip(idx[:off[n]])
    sh, th = np.ascontiguousarray(s_hist.copy()), np.ascontiguousarray(t_hist.copy())
    q = np.zeros(N); s = np.zeros(n); t = np.zeros(n); mv = np.zeros(2, dtype=np.int32); tap = np.zeros(8)
    cnt = C.c_long(); rows = np.zeros(n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ty, tg, ow = _ip(cfg["type"]), _ip(cfg["tag"]), _ip(cfg["owner"])
    rc = L.ref_qeq_pre_force(n, N - n, 4, p(x), p(ty), p(tg), p(off), p(idx), p(ow), p(chi), p(eta), p(gam), C.c_double(0.0),
                             C.c_double(swb), C.c_double(tol), p(sh), p(th), p(q), p(s), p(t), p(mv), p(tap), C.byref(cnt), p(rows))
    assert rc == 0
    return dict(q=q, s=s, t=t, mv=(int(mv[0]), int(mv[1])), tap=tap, s_hist=sh, t_hist=th, nnz=cnt.value, rowsum=rows)


@pytest.mark.parametrize("perturb,seed,tol,history", [(0.0, 0, 1e-6, False), (0.1, 13, 1e-6, True), (0.1, 14, 1e-10, True)])
def test_qeq_equals_reference_fix_qeq_reax(perturb, seed, tol, history):
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, qeq=False)
    o = cfg["oracle"]
    n, x = cfg["n"], cfg["x"]
    o.set_atoms(n, x, cfg["type"], cfg["tag"], np.zeros(len(x)))
    o.build_neighbors(12.5)
    rng = np.random.default_rng(seed)
    sh = rng.normal(scale=0.05, size=(n, 5)) if history else np.zeros((n, 5))
    th = rng.normal(scale=0.05, size=(n, 5)) if history else np.zeros((n, 5))
    ref = _ref_qeq(cfg, o, tol, sh, th)
    o.qeq_init(0.0, 10.0, tol)
    o.qeq_set_hist(sh, th)
    mv = o.qeq_pre_force(cfg["owner"])
    # the pipelined-CG iteration counts of both solves are the reference's, exactly
    assert mv == ref["mv"] and max(mv) < 200 and min(mv) > 3
    offH, numH, colH, valH = o.qeq_H()
    assert int(numH.sum()) == ref["nnz"]
    rows = np.array([valH[offH[i]:offH[i] + numH[i]].sum() for i in range(n)])
    assert relerr(rows, ref["rowsum"]) < 1e-13            # H through the reference's calculate_H + init_taper
    s, t = o.qeq_st()
    assert relerr(s[:n], ref["s"]) < 1e-9 and relerr(t[:n], ref["t"]) < 1e-9
    assert np.abs(o.q() - ref["q"]).max() < 1e-10 and np.abs(ref["q"][:n]).max() > 0.1
    so, to = o.qeq_get_hist()
    assert relerr(so, ref["s_hist"]) < 1e-9 and relerr(to, ref["t_hist"]) < 1e-9
    assert np.array_equal(ref["s_hist"][:, 1:], sh[:, :4])  # history shifted by one


@pytest.mark.parametrize("perturb,seed,scale", [(0.0, 0, 1.0), (0.1, 15, 1.0), (0.1, 16, 0.9)])
def test_bond_order_corrections_equal_reference_BO_serial_body(perturb, seed, scale):
    """a6: the serial bond-order correction code of BO() (reaxc_bond_orders_sunway.cpp:460-774), which the live build
    skips with an early `return;` in favour of the slave-core kernels — compiled from the unmodified file with that one
    `return` defined away (oracle/ref/stubs/prelude_bo_serial.h) and run on the oracle's uncorrected bond list."""
    L, P = _ref_lib()
    cfg = H.static_config(1, 1, 1, perturb=perturb, seed=seed, scale=scale, qeq=False)
    o = cfg["oracle"]
    n, x = cfg["n"], cfg["x"]
    N = len(x)
    o.set_atoms(n, x, cfg["type"], cfg["tag"], np.zeros(N))
    o.build_neighbors(12.5)
    o.phase(0); o.phase(1)
    bs, be, nbr, sym, fld = o.bonds()
    nb = len(nbr)
    total_bop = np.ascontiguousarray(o.workspace()[:, 0])          # sum of BO' per atom after the bond-list build
    assert total_bop.max() > 2.0
    fref = np.ascontiguousarray(fld.copy()); wref = np.zeros((N, 16))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ints = [_ip(a) for a in ([t - 1 for t in cfg["type"]], cfg["tag"], bs, be, nbr, sym)]
    assert L.ref_bond_orders(P, n, N, p(ints[0]), p(ints[1]), p(ints[2]), p(ints[3]), nb, p(ints[4]), p(ints[5]), p(total_bop),
                             p(fref), p(wref)) == 0
    o.phase(4)
    _, _, _, _, forc = o.bonds()
    worc = o.workspace()
    for c in (4, 5, 6, 7) + tuple(range(17, 28)):                  # BO, BO_s, BO_pi, BO_pi2, C1dbo .. C4dbopi2
        den = max(np.abs(forc[:, c]).max(), 1e-300)
        assert np.abs(fref[:, c] - forc[:, c]).max() / den < 1e-11, c
    assert np.abs(forc[:, 4] - fld[:, 4]).max() > 1e-3             # the corrections did change the bond orders
    for c in range(15):                                            # total_bo, Delta_boc, Deltap, ..., dDelta_lp_temp
        den = max(np.abs(worc[:, c]).max(), 1e-300)
        assert np.abs(wref[:, c] - worc[:, c]).max() / den < 1e-11, c


# ---------------------------------------------------------------------------------------------------------------------
# f1 / f2: the reference's own output fixes (fix_reaxc_bonds_sunway.cpp, fix_reaxc_species_sunway.cpp compiled unmodified
# against the LAMMPS-core stand-in, oracle/ref/ref_analysis.cpp) write their files from the oracle's state; the oracle's
# restated writers (oracle/orc_analysis.cpp) must produce the same bytes.
@pytest.mark.parametrize("scale,T", [(1.0, 300.0), (0.9, 3000.0)])
def test_bond_table_file_equals_reference_fix_reaxc_bonds(scale, T, tmp_path):
    L = C.CDLL(LIBREF)
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=scale)
    v = H.maxwell_velocities(t, T, 31)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-8)
    o.md_run(3)
    n = o.nlocal
    xall, ty, tg, owner = o.md_ghosts()
    N = len(xall)
    o.n, o.N = n, N
    bs, be, nbr, sym, fld = o.bonds()
    w = o.workspace()
    q = np.ascontiguousarray(o.q())
    path = str(tmp_path / "bonds.ref")
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    BO, tbo, nlp = np.ascontiguousarray(fld[:, 4]), np.ascontiguousarray(w[:, 0]), np.ascontiguousarray(w[:, 8])
    ints = [_ip(a) for a in (ty, tg, bs, be, nbr)]
    assert L.ref_bonds_write(path.encode(), C.c_long(3), C.c_double(0.3), n, N - n, 4, p(ints[0]), p(ints[1]), p(q), p(ints[2]),
                             p(ints[3]), len(nbr), p(ints[4]), p(BO), p(tbo), p(nlp)) == 0
    ref = open(path).read()
    assert ref == o.md_bonds_text(3) and ref.count("\n") == 384 + 8


@pytest.mark.parametrize("scale,T,cut", [(1.0, 300.0, None), (0.8, 4000.0, (1, 4, 0.9))])
def test_species_file_equals_reference_fix_reaxc_species(scale, T, cut, tmp_path):
    L = C.CDLL(LIBREF)
    box, x, t, tag = H.tatb_cell(1, 1, 1, scale=scale)
    v = H.maxwell_velocities(t, T, 99)
    o = H.Oracle()
    o.md_init(box, x, v, t, tag, dt=0.0625, qeq_tol=1e-8, every=5)
    bc = np.full((5, 5), 0.30)
    if cut:
        bc[cut[0], cut[1]] = bc[cut[1], cut[0]] = cut[2]
    o.md_species_init(1, 5, 5, bocut=bc)
    o.md_species_step(0)
    found = False
    for step in range(1, 6):
        found = o.md_species_step(step)
        if step < 5:
            o.md_run(1)
    assert found
    sp = o.md_species_get()
    ids, avg = o.md_species_raw()
    xall, ty, tg, owner = o.md_ghosts()
    n, N = o.nlocal, len(xall)
    path = str(tmp_path / "species.ref")
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cuts = np.array([[cut[0], cut[1], cut[2]]], dtype=np.float64) if cut else np.zeros((0, 3))
    nmole = C.c_int(); every = C.c_int()
    cl = np.zeros(n, dtype=np.int32)
    ints = [_ip(a) for a in (ty, tg, owner)]
    assert L.ref_species_write(path.encode(), 1, 5, 5, n, N - n, 4, p(ints[0]), p(ints[1]), p(ints[2]), p(ids), p(avg), len(cuts),
                               p(cuts) if len(cuts) else None, C.byref(nmole), p(cl), C.byref(every)) == 0
    assert every.value == 5                                   # the reference keeps `every 5` for nevery*nrepeat = 5
    assert nmole.value == sp["nmole"] and np.array_equal(cl, sp["cluster"])
    assert open(path).read() == o.md_species_text(5)
    if cut:
        assert sp["nmole"] > 16
    # `position 5 <file>`: the reference's WritePos on the averaged q, x, y, z columns against the oracle's restatement
    box6 = np.array([0.0, 0.0, 0.0, box[0], box[1], box[2]])
    pos_orc, qxyz = o.md_species_pos(5, box6)
    assert np.abs(qxyz[:, 1:]).max() > 1.0 and np.abs(qxyz[:, 0]).max() > 1e-3
    pos_path = str(tmp_path / "species.pos")
    assert L.ref_species_write_pos(str(tmp_path / "species2.ref").encode(), 1, 5, 5, n, N - n, 4, p(ints[0]), p(ints[1]), p(ints[2]),
                                   p(ids), p(avg), len(cuts), p(cuts) if len(cuts) else None, C.byref(nmole), p(cl),
                                   C.byref(every), pos_path.encode(), 5, p(np.ascontiguousarray(qxyz)), p(box6)) == 0
    ref_pos = open(pos_path).read()
    assert ref_pos == pos_orc
    assert ref_pos.count("\n") == sp["nmole"] + 3
