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
