"""CPU: the two scalar kernels of math_b200/csrc/device_math.cuh that replace libdevice calls in
the link functions (exp_nonpos, log1p_nonneg), restated on the host with the same operations in
the same order (profiles/numerics/*.c), stay within their documented error against the x87
long-double libm -- and the restatements use the constants the device code uses."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NUM = os.path.join(ROOT, "profiles", "numerics")
CUH = open(os.path.join(ROOT, "math_b200", "csrc", "device_math.cuh")).read()


def run(src, tmp_path):
    exe = str(tmp_path / "chk")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(NUM, src),
                           "-lm"])
    return subprocess.check_output([exe], text=True)


def test_exp_nonpos_within_one_ulp(tmp_path):
    out = run("check_exp_nonpos.c", tmp_path)
    worst = float(re.search(r"custom ([0-9.]+)", out).group(1))
    assert worst < 1.0, out
    assert "f(0)=1 " in out and "f(-709)=0 " in out, out
    src = open(os.path.join(NUM, "check_exp_nonpos.c")).read()
    for const in ("1.4426950408889634074", "6.93147180369123816490e-01",
                  "1.90821492927058770002e-10", "6227020800.0"):
        assert const in src and const in CUH, const


def test_log1p_nonneg_within_two_ulp(tmp_path):
    out = run("check_log1p_nonneg.c", tmp_path)
    worst = float(re.search(r"fast ([0-9.]+)", out).group(1))
    assert worst < 1.6, out
    assert "f(0)=0" in out, out
    assert "const double c = e - (u - 1.0);" in CUH
