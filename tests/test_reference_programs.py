"""The reference's OWN cmocka test programs (omega_nu_single_test.c, transfer_init_test.c, delta_pow_test.c,
powerspectrum_test.c, delta_tot_table_test.c -- reference Makefile:18-21, BASELINE.json configs[0]) run against the
product instead of the reference's .c files.  The sources are compiled where they lie under /root/reference (never
copied) against include/*.h; cmocka, <mpi.h>, FFTW2 and GSL come from the oracle's shims (test infrastructure).

CPU part (this file):
 * the three host-only programs, linked to the product's libkspace_neutrinos_b200.so (oracle/_ref/dropin_*_test, built by
   oracle/Makefile), must pass as they are;
 * the two programs that reach the kernels must, without a GPU, stop at the device boundary with the product's loud
   "no CPU path" error;
 * all five must pass when the product's HOST layer (kspace_neutrinos_b200/src/*.c) is linked with the test-only CPU
   stand-in for the device entry points (tests/device_standin.c, built on the oracle): that exercises the host layer's
   state machine, table layout and save/resume files exactly as the reference's tests poke them.
GPU part: tests/test_zz_reference_programs_gpu.py."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kspace_neutrinos_b200")
ORACLE = os.path.join(ROOT, "oracle")
REFDIR = os.path.join(ORACLE, "_ref")
REF = "/root/reference"
PROGRAMS = ("omega_nu_single", "transfer_init", "delta_pow", "powerspectrum", "delta_tot_table")


def fixtures_dir(tmp_path):
    """A scratch working directory holding testdata/ and camb_linear/ (test_save_resume writes into testdata/,
    delta_tot_table_test.c:113)."""
    src = REF if os.path.isdir(os.path.join(REF, "testdata")) else os.path.join(REFDIR, "fixtures")
    if not os.path.isdir(os.path.join(src, "testdata")):
        pytest.skip("reference fixtures not available (oracle/_ref/fixtures is built where /root/reference exists)")
    shutil.copytree(os.path.join(src, "testdata"), tmp_path / "testdata")
    os.symlink(os.path.join(src, "camb_linear"), tmp_path / "camb_linear")
    return str(tmp_path)


def dropin(name):
    exe = os.path.join(REFDIR, f"dropin_{name}_test")
    if not os.path.exists(exe) and os.path.isdir(REF):
        subprocess.run(["make", "-C", ORACLE, "-s", "dropin"], capture_output=True)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    return exe


def run(exe, cwd, env=None, timeout=600):
    r = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=timeout, env={**os.environ, **(env or {})})
    return r.returncode, r.stdout + r.stderr


def n_run_failed(out):
    line = [l for l in out.splitlines() if l.startswith("[==========]")]
    assert line, out[-2000:]
    w = line[-1].split()
    return int(w[1]), int(w[w.index("run,") + 1])


@pytest.mark.parametrize("name,ntests", [("omega_nu_single", 9), ("transfer_init", 1), ("delta_pow", 2)])
def test_host_only_reference_programs_pass_against_the_product_library(name, ntests, tmp_path):
    rc, out = run(dropin(name), fixtures_dir(tmp_path))
    assert rc == 0, out[-3000:]
    assert n_run_failed(out) == (ntests, 0), out[-3000:]


@pytest.mark.parametrize("name", ["powerspectrum", "delta_tot_table"])
def test_device_programs_stop_loudly_without_a_gpu(name, tmp_path, ksn):
    if ksn.ksn_device_available():
        pytest.skip("a GPU is visible: covered by tests/test_zz_reference_programs_gpu.py")
    rc, out = run(dropin(name), fixtures_dir(tmp_path))
    assert rc != 0 and "this library has no CPU path" in out, out[-2000:]
    assert "[       OK ]" not in out.split("no usable B200 device")[-1]


@pytest.fixture(scope="module")
def standin_build(tmp_path_factory):
    if not os.path.isdir(REF):
        pytest.skip("needs the reference's test sources under /root/reference")
    out = tmp_path_factory.mktemp("standin")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c")))
    objs = []
    common = ["gcc", "-O2", "-g", "-DPERIODIC", "-DKSN_HAVE_MPI", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ORACLE, "shim"),
              "-I", ORACLE, "-I", os.path.join(PKG, "src")]
    for s in srcs + [os.path.join(ROOT, "tests", "device_standin.c")] + [os.path.join(ORACLE, f) for f in ("ksn_oracle.c", "mini_gsl.c", "mini_cmocka.c", "mini_mpi.c")]:
        o = str(out / (os.path.basename(s)[:-2] + ".o"))
        r = subprocess.run(common + ["-DDOUBLEPRECISION_FFTW", "-c", s, "-o", o], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        objs.append(o)

    def build(name, double=True):
        exe = str(out / f"{name}_{'d' if double else 's'}_test")
        with open(os.path.join(REF, f"{name}_test.c")) as src:
            r = subprocess.run(common + (["-DDOUBLEPRECISION_FFTW"] if double else []) + ["-x", "c", "-", "-x", "none", *objs, "-o", exe, "-lm", "-lpthread"],
                               stdin=src, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        return exe
    return build


@pytest.mark.parametrize("name,double,ntests", [("omega_nu_single", True, 9), ("transfer_init", True, 1), ("delta_pow", True, 2),
                                                ("powerspectrum", True, 1), ("powerspectrum", False, 1), ("delta_tot_table", True, 7)])
def test_reference_programs_pass_against_the_host_layer_with_the_cpu_standin(standin_build, name, double, ntests, tmp_path):
    rc, out = run(standin_build(name, double), fixtures_dir(tmp_path))
    assert rc == 0, out[-3000:]
    assert n_run_failed(out) == (ntests, 0), out[-3000:]
