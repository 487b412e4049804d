"""CPU tests of the product's host side: C ABI surface, ffield/control parsing vs the oracle, error behaviour.

No compute entry point is called here (no GPU in this container); librxb200.so is only loaded and its host-only parser run.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from sw_reaxff_b200 import api

ROOT = H.ROOT


def lib():
    return api.load_library()


def parse_dump(control, ffield, elements, lgvdw=0, enobonds=1):
    L = lib()
    L.rxb_parse_dump.restype = C.c_long
    arr = (C.c_char_p * len(elements))(*[e.encode() for e in elements])
    n = L.rxb_parse_dump(control.encode() if control else None, ffield.encode(), len(elements), arr, lgvdw, enobonds, None, C.c_long(0))
    if n < 0:
        raise RuntimeError(L.rxb_last_error().decode())
    out = np.zeros(n)
    L.rxb_parse_dump(control.encode() if control else None, ffield.encode(), len(elements), arr, lgvdw, enobonds,
                     out.ctypes.data_as(C.c_void_p), C.c_long(n))
    return out


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rxb200.h")).read()
    declared = sorted(set(re.findall(r"\b(rxb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 28
    L = lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/rxb200.h but not exported by librxb200.so"
    assert set(api.SYMBOLS) <= set(declared)


def test_parser_matches_oracle_bit_for_bit():
    mine = parse_dump(H.CONTROL, H.FFIELD, H.ELEMENTS)
    ref = H.Oracle().params_dump()
    assert mine.shape == ref.shape
    assert np.array_equal(mine, ref)  # parameters are parsed text: exact equality


def test_parser_element_permutation_and_null_control():
    mine = parse_dump(None, H.FFIELD, ["N", "O", "H", "C"])
    ref = H.Oracle(control=None, elements=["N", "O", "H", "C"]).params_dump()
    assert np.array_equal(mine, ref)
    # defaults of pair_style reax/c NULL (pair_reaxc_sunway.cpp:208-232): bond_cut 5, hbond_cut 7.5
    nt, ngp = int(mine[0]), int(mine[2])
    ctl = mine[3 + ngp:3 + ngp + 10]
    assert ctl[3] == 5.0 and ctl[4] == 7.5 and ctl[0] == pytest.approx(1e-4) and ctl[2] == 10.0


def test_known_tatb_parameters():
    d = parse_dump(H.CONTROL, H.FFIELD, H.ELEMENTS)
    nt, vdw_type, ngp = int(d[0]), int(d[1]), int(d[2])
    assert (nt, vdw_type, ngp) == (4, 1, 39)
    gp = d[3:3 + ngp]
    assert gp[28] == 1.5591 and gp[12] == 10.0
    ctl = d[3 + ngp:3 + ngp + 10]
    assert list(ctl[:8]) == [pytest.approx(1e-4), 0.0, 10.0, 4.5, 6.0, 0.3, 0.001, 1e-5]
    tap = d[3 + ngp + 10:3 + ngp + 18]
    r = 10.0
    assert abs(np.polyval(tap[::-1], r)) < 1e-12 and abs(np.polyval(tap[::-1], 0.0) - 1.0) < 1e-15


def test_parser_errors_are_reported_not_fatal(tmp_path):
    with pytest.raises(RuntimeError, match="Cannot open"):
        parse_dump(H.CONTROL, str(tmp_path / "nope.reax"), H.ELEMENTS)
    bad = tmp_path / "control.bad"
    bad.write_text("nbrhood_cutoff 4.5\nnot_a_keyword 1\n")
    with pytest.raises(RuntimeError, match="unknown parameter"):
        parse_dump(str(bad), H.FFIELD, H.ELEMENTS)
    with pytest.raises(RuntimeError, match="Non-existent ReaxFF type"):
        parse_dump(H.CONTROL, H.FFIELD, ["C", "H", "O", "Xx"])


def test_no_cpu_fallback_without_gpu():
    if H_have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(api.RxbError, match="no CUDA device|CUDA"):
        api.Rxb(0)


def H_have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_product_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/ (task rule ③)."""
    pkg = os.path.join(ROOT, "sw_reaxff_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                for pat in ("liboracle", "orc_", "oracle/", "import oracle", "from oracle", "import helpers"):
                    assert pat not in txt, (dirpath, fn, pat)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "librxb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_lookup_tables_equal_oracle_and_track_the_analytic_form(tmp_path):
    """Tabulated long-range mode (control: tabulate_long_range N; reaxc_lookup_sunway.cpp:157-285, dead code in the
    reference): the product's table builder (host C++, device layout) against the oracle's restatement."""
    import helpers as H
    N = 2000
    ctl = H.control_variant(tmp_path / "control.tab", N)
    L = lib()
    L.rxb_lookup_dump.restype = C.c_long
    arr = (C.c_char_p * 4)(*[e.encode() for e in H.ELEMENTS])
    n = C.c_int()
    cnt = L.rxb_lookup_dump(ctl.encode(), H.FFIELD.encode(), 4, arr, C.byref(n), None, C.c_long(0))
    assert cnt == 16 * (N + 2) * 16 and n.value == N + 2
    mine = np.zeros(cnt)
    L.rxb_lookup_dump(ctl.encode(), H.FFIELD.encode(), 4, arr, C.byref(n), mine.ctypes.data_as(C.c_void_p), C.c_long(cnt))
    mine = mine.reshape(4, 4, N + 2, 4, 4)
    o = H.Oracle(control=ctl)
    o.L.orc_lookup_tables.restype = C.c_int
    assert o.L.orc_lookup_tables(o.h, None) == N + 2
    orc = np.zeros((4, 4, 5, N + 2, 4))
    o.L.orc_lookup_tables(o.h, orc.ctypes.data_as(C.c_void_p))
    for t_mine, t_orc in ((0, 2), (1, 4), (2, 1), (3, 3)):          # CEvd, CEclmb, e_vdW, e_ele
        for i in range(4):
            for j in range(i, 4):
                a = mine[i, j, 1:N + 1, t_mine, :]; b = orc[i, j, t_orc, 1:N + 1, :]
                scale = np.abs(b).max(axis=0) + 1e-300
                assert (np.abs(a - b) / scale).max() < 1e-9, (t_mine, i, j)
                assert np.array_equal(mine[j, i, :, t_mine, :], mine[i, j, :, t_mine, :])
    # tabulate 0 -> no tables
    assert L.rxb_lookup_dump(H.CONTROL.encode(), H.FFIELD.encode(), 4, arr, C.byref(n), None, C.c_long(0)) == 0
