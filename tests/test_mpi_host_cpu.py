"""The host C layer as an MPI PM code builds it (-DKSN_HAVE_MPI), run as two ranks over the oracle's fork-based mini-MPI:
parameter block, transfer table and collective binding must reach the rank that has neither the parameters nor the file
(interface_common.c:54-75).  No GPU involved: nothing here computes on the grid."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kspace_neutrinos_b200")


def test_two_rank_mpi_build_of_the_host_layer(tmp_path):
    lib = os.path.join(PKG, "libkspace_neutrinos_b200.so")
    assert os.path.exists(lib), "build the product first (__graft_entry__.build())"
    exe = str(tmp_path / "mpi_host_flow")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c")))
    cmd = ["gcc", "-O1", "-g", "-Wall", "-DKSN_HAVE_MPI", "-DDOUBLEPRECISION_FFTW",
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "src"),
           os.path.join(ROOT, "tests", "mpi_host_flow.c"), *srcs, os.path.join(ROOT, "oracle", "mini_mpi.c"),
           "-L", PKG, "-lkspace_neutrinos_b200", f"-Wl,-rpath,{PKG}", "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "MPI HOST FLOW OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
