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
    cmd = ["gcc", "-O1", "-g", "-Wall", "-Werror=implicit-function-declaration", "-DKSN_HAVE_MPI", "-DDOUBLEPRECISION_FFTW",
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "src"),
           os.path.join(ROOT, "tests", "mpi_host_flow.c"), *srcs, os.path.join(ROOT, "oracle", "mini_mpi.c"),
           "-L", PKG, "-lkspace_neutrinos_b200", f"-Wl,-rpath,{PKG}", "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "MPI HOST FLOW OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def test_slab_split_pm_steps_over_mini_mpi(tmp_path):
    """add_nu_power_to_rhogrid on x-slabs over 1, 2 and 3 (uneven) ranks: host layer (-DKSN_HAVE_MPI) + the test-only CPU
    stand-in for the device entry points.  Same integrator state on every rank, same corrected grid for every split."""
    import numpy as np
    exe = str(tmp_path / "mpi_host_step")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c")))
    orc = [os.path.join(ROOT, "oracle", f) for f in ("ksn_oracle.c", "mini_gsl.c", "mini_mpi.c")]
    cmd = ["gcc", "-O2", "-g", "-Wall", "-Werror=implicit-function-declaration", "-DKSN_HAVE_MPI", "-DDOUBLEPRECISION_FFTW",
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "src"),
           os.path.join(ROOT, "tests", "mpi_host_step.c"), os.path.join(ROOT, "tests", "device_standin.c"), *srcs, *orc,
           "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    res = {}
    # the backend bootstrap (iface_common.c bind_comm): no device -> host call-back (1) on every rank; pretended peer
    # memory -> (2) after the handle exchange and the trial sum; one rank failing -> all ranks fall back together
    for env, ranks, want in (({"KSN_STANDIN_P2P": "1"}, 3, 2), ({"KSN_STANDIN_P2P": "1", "KSN_STANDIN_P2P_FAIL_RANK": "1"}, 3, 1),
                             ({"KSN_STANDIN_P2P": "1", "KSN_COMM": "mpi"}, 2, 1), ({"KSN_STANDIN_P2P": "1", "KSN_COMM": "nccl"}, 2, 1)):
        out = str(tmp_path / "boot.bin")
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat"), str(ranks), out], capture_output=True, text=True,
                           timeout=300, env={**os.environ, **env})
        assert r.returncode == 0 and "MPI HOST STEP OK" in r.stdout and f"BACKEND {want}" in r.stdout, (env, r.stdout[-1000:] + r.stderr[-2000:])
    for ranks in (1, 2, 3):
        out = str(tmp_path / f"out{ranks}.bin")
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat"), str(ranks), out], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "MPI HOST STEP OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
        assert f"BACKEND {0 if ranks == 1 else 1}" in r.stdout, r.stdout[-1000:]
        raw = open(out, "rb").read()
        n, nk, ia = np.frombuffer(raw[:12], dtype=np.int32)
        dnu = np.frombuffer(raw[12:12 + 8 * nk], dtype=np.float64)
        grid = np.frombuffer(raw[12 + 8 * nk:], dtype=np.float64)
        assert grid.size == 2 * n * n * (n // 2 + 1)
        res[ranks] = (int(nk), int(ia), dnu, grid)
    assert res[1][0] > 0 and res[1][1] == 4                     # 0.0205 is too close to 0.02 to be kept (delta_tot_table.c:229)
    for ranks in (2, 3):
        assert res[ranks][:2] == res[1][:2]
        # the cross-rank sum adds the slab sums in a different order than one rank's sweep: last-bit differences only
        np.testing.assert_allclose(res[ranks][2], res[1][2], rtol=1e-12)
        np.testing.assert_allclose(res[ranks][3], res[1][3], rtol=1e-12)
    assert not np.array_equal(res[1][3][2:], np.zeros_like(res[1][3][2:]))


def test_communicator_passed_to_total_powerspectrum_is_bound_without_InitOmegaNu(tmp_path):
    """The KSPACE_NEUTRINOS_2-off hook (gadget-2 patch 0004) calls compute_total_power_spectrum(..., comm) and nothing else:
    the collective must be bound from that argument (powerspectrum.c:91-95 reduces on the communicator it is handed)."""
    import numpy as np
    exe = str(tmp_path / "mpi_total_power")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c")))
    orc = [os.path.join(ROOT, "oracle", f) for f in ("ksn_oracle.c", "mini_gsl.c", "mini_mpi.c")]
    cmd = ["gcc", "-O2", "-g", "-Wall", "-Werror=implicit-function-declaration", "-DKSN_HAVE_MPI", "-DDOUBLEPRECISION_FFTW",
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "src"),
           os.path.join(ROOT, "tests", "mpi_total_power.c"), os.path.join(ROOT, "tests", "device_standin.c"), *srcs, *orc,
           "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    res = {}
    for ranks in (1, 2, 3):
        out = str(tmp_path / f"tp{ranks}.bin")
        r = subprocess.run([exe, str(ranks), out], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "MPI TOTAL POWER OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
        res[ranks] = np.fromfile(out)
    nb = (res[1].size - 1) // 3
    for ranks in (2, 3):
        assert res[ranks][0] == res[1][0]
        assert np.array_equal(res[ranks][1 + 2 * nb:], res[1][1 + 2 * nb:])            # counts
        np.testing.assert_allclose(res[ranks][1:1 + 2 * nb], res[1][1:1 + 2 * nb], rtol=1e-12)
