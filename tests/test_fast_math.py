"""Accuracy of the straight-line fp64 exp / log / inverse cube root used by the pair and bonded kernels
(sw_reaxff_b200/csrc/rxb_math.cuh).  The header compiles unchanged for the host; every operation in it is an explicit
fma / add / mul on IEEE doubles, so the numbers measured here are the numbers the sm_100a kernels produce (the cube root's
fp32 seed aside, which is covered by perturbing the seed)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_check(tmp_path):
    exe = str(tmp_path / "fast_math_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "sw_reaxff_b200", "csrc"),
                    os.path.join(ROOT, "tests", "fast_math_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout
    return {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}


def test_fast_math_accuracy(tmp_path):
    r = _run_check(tmp_path)
    assert r["exp_rel"] < 3e-16            # <= 1.5 ulp over [-700, 700]
    assert r["log_abs_over_max1"] < 4e-16  # absolute error <= ~1 ulp of max(1, |log x|)
    assert r["log_abs_near_1"] < 2e-16
    assert r["pow_rel"] < 2e-15            # exp(p log r): what the vdW kernel evaluates
    assert r["rcbrt_rel"] < 3e-16
    assert r["rcbrt_seed1e-5_rel"] < 3e-16
    assert r["edge_log"] < 4e-16 and r["edge_exp"] < 3e-16   # powers of two, interval boundaries, the clamp ends
    assert r["mono_bad"] == 0                                # r -> r^p never decreases beyond rounding
    assert r["exp0"] == 1.0 and r["log1"] == 0.0


def test_tables_are_the_generated_ones():
    """rxb_math_tables.h is reproducible from gen_math_tables.py."""
    gen = os.path.join(ROOT, "sw_reaxff_b200", "csrc", "gen_math_tables.py")
    out = subprocess.run(["python", gen], check=True, stdout=subprocess.PIPE, text=True).stdout
    with open(os.path.join(ROOT, "sw_reaxff_b200", "csrc", "rxb_math_tables.h")) as f:
        assert f.read() == out
