"""The bookkeeping of K2's speculative bisection (kspace_neutrinos_b200/csrc/ksn_qag_spec.h -- the header nvcc compiles
into k2_delta_nu_spec_kernel) replayed on the CPU: whatever the speculation width, result, error estimate, status and the
count of rule applications equal the sequential QAG loop's and the oracle's mini-GSL gsl_integration_qag bit for bit."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replay_equals_sequential_qag(tmp_path):
    exe = str(tmp_path / "qag_spec_host")
    inc = ["-I", os.path.join(ROOT, "kspace_neutrinos_b200", "csrc"), "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle")]
    obj = str(tmp_path / "mini_gsl.o")
    subprocess.run(["gcc", "-O2", *inc, "-c", os.path.join(ROOT, "oracle", "mini_gsl.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-O2", *inc, os.path.join(ROOT, "tests", "qag_spec_host.cpp"), obj, "-lm", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ALL IDENTICAL" in r.stdout and "DIFFERS" not in r.stdout
    # the point of the exercise: an oscillatory integrand (the hybrid-neutrino regime) needs far fewer passes at M = 4
    m = re.search(r"cos2000 .*?M=1: trips\s+(\d+).*?M=4: trips\s+(\d+)", r.stdout, re.S)
    assert m and int(m.group(2)) * 3 < int(m.group(1)), r.stdout
