"""CPU test: the product's branch-free FP64 math (kf_bayesopt_b200/csrc/kfbo_math.cuh, the code the
CUDA kernels run) compiled as plain host C++ and checked against long-double libm."""
import os
import subprocess

from conftest import ROOT


def test_custom_math_accuracy(tmp_path):
    exe = str(tmp_path / "math_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "math_check.cpp"), "-lm"])
    out = subprocess.check_output([exe], text=True)
    got = {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}
    # ulps against the correctly rounded result; the normals' absolute error (|n| <= 8.5)
    assert got["rcp_ulp"] <= 1.0
    assert got["sqrt_ulp"] <= 2.0
    # degree-3 economised polynomial (KFBO_LOG_DEGREE 30): <= 15 ulp of L, reached only where L < 2^-10
    assert got["neg2log_ulp"] <= 15.0
    # sin/cos of the Box-Muller angle: absolute error in units of 2^-53 (half an ulp of 1)
    assert got["sin_ulp"] <= 3.0 and got["cos_ulp"] <= 3.0
    assert got["na_abs"] < 4e-15 and got["nb_abs"] < 4e-15
