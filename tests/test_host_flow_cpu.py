"""The product's HOST layer (kspace_neutrinos_b200/src/*.c) run on the CPU, with the test-only stand-in for the device
entry points (tests/device_standin.c, built on the oracle) in place of the CUDA objects: the reference-facing flow --
kspace_params -> InitOmegaNu -> allocate_kspace_memory -> add_nu_power_to_rhogrid per step, compute_neutrino_power_from_cdm,
save/read state files, get/set_nu_state -- against the reference's own sources (oracle/_ref/libksref_double.so) on the
same inputs.  Says nothing about the kernels (tests -m gpu do); it pins the glue either side of them where no GPU exists."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi
from tests import refs
from tests.test_step_gpu import _run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kspace_neutrinos_b200")
ORACLE = os.path.join(ROOT, "oracle")
HOST_API = [n for n in capi.PROTOTYPES if not n.startswith("ksn_") or n in ("ksn_set_default_hubble", "ksn_set_quiet", "ksn_global_omnu", "ksn_global_transfer")]


@pytest.fixture(scope="module")
def standin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hoststandin") / "libksn_hoststandin.so")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c"))) + [os.path.join(ROOT, "tests", "device_standin.c")] + \
        [os.path.join(ORACLE, f) for f in ("ksn_oracle.c", "mini_gsl.c")]
    r = subprocess.run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-DDOUBLEPRECISION_FFTW", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ORACLE, "shim"), "-I", ORACLE, "-I", os.path.join(PKG, "src"), *srcs, "-o", out, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    h = C.CDLL(out, mode=C.RTLD_LOCAL)
    refs._attach(h, [n for n in HOST_API if hasattr(h, n)])
    h.ksn_set_quiet(1)
    return h


@pytest.fixture(scope="module")
def ref():
    r = refs.ref_lib(True)
    if r is None:
        pytest.skip("oracle/_ref not built")
    return r


@pytest.mark.parametrize("n,hybrid,masses", [(32, False, (0.15, 0.15, 0.15)), (32, True, (0.15, 0.15, 0.15)), (24, False, (0.2, 0.1, 0.3))])
def test_add_nu_power_host_flow_matches_reference(standin, ref, n, hybrid, masses):
    g = refs.random_grid(n, seed=3 + n)
    times = (0.01, 0.02, 0.0205, 0.05, 0.2, 0.34, 0.345, 0.5)
    want = _run(ref, "add_nu_power_to_rhogrid", g, times, False, hybrid=hybrid, masses=masses)
    got = _run(standin, "add_nu_power_to_rhogrid_f64", g, times, False, hybrid=hybrid, masses=masses)
    for (ia_r, nk_r, g_r, dn_r), (ia_g, nk_g, g_g, dn_g) in zip(want, got):
        assert (ia_r, nk_r) == (ia_g, nk_g)
        np.testing.assert_allclose(dn_g, dn_r, rtol=1e-10, atol=0)
        np.testing.assert_allclose(g_g, g_r, rtol=1e-10, atol=0)


def test_state_files_round_trip_and_cross_load(standin, ref, tmp_path):
    """save_nu_state / read_all_nu_state (delta_tot_table.c:254-374): the file the product's host layer writes after a few
    steps is byte-identical to the reference's, and each library resumes from the other's file to the same delta_nu."""
    n = 32
    g = refs.random_grid(n, seed=11)
    times = (0.01, 0.02, 0.05, 0.1)
    files = {}
    for name, libh, fn in (("ref", ref, "add_nu_power_to_rhogrid"), ("got", standin, "add_nu_power_to_rhogrid_f64")):
        _run(libh, fn, g, times, False)
        files[name] = tmp_path / f"{name}_delta_tot_nu.txt"
        libh.save_nu_state(str(files[name]).encode())
        assert files[name].exists()
        libh.save_nu_state(str(files[name]).encode())             # a second save moves the old file to .bak (:339-348)
        assert (tmp_path / f"{name}_delta_tot_nu.txt.bak").exists()
    a, b = files["ref"].read_bytes(), files["got"].read_bytes()
    assert len(a) > 0 and a == b                                   # '# %le' rows, one per stored scale factor (:254-262)
    # resume: each library reads the OTHER one's file, then takes the same next step
    res = {}
    for name, libh, fn, other in (("ref", ref, "add_nu_power_to_rhogrid", "got"), ("got", standin, "add_nu_power_to_rhogrid_f64", "ref")):
        om, dt = refs.init_module(libh, n)
        libh.read_all_nu_state(C.byref(dt), str(files[other]).encode())
        assert dt.ia == len(times)
        gg = g.copy()
        getattr(libh, fn)(0.15, refs.BOX, gg.ctypes.data_as(C.c_void_p), n, 0, n, 0)
        res[name] = (dt.ia, np.array([dt.delta_nu_last[i] for i in range(dt.nk)]), gg)
    assert res["ref"][0] == res["got"][0]
    np.testing.assert_allclose(res["got"][1], res["ref"][1], rtol=1e-5)      # the files carry 7 digits of the history
    np.testing.assert_allclose(res["got"][2], res["ref"][2], rtol=1e-6)


def test_compute_neutrino_power_from_cdm_host_flow(standin, ref):
    """interface_common.c:106-123 (host supplies P(k), empty bins dropped) through the product's host layer."""
    n = 64
    kk, _, _ = refs.load_golden_state()
    nk_in = n // 2
    keff = np.ascontiguousarray(kk[::9][:nk_in])
    P = (1e5 * (keff / keff[0]) ** -0.7) ** 2
    nmodes = np.ones(nk_in, dtype=np.int64)
    nmodes[[3, 17]] = 0
    nm = nmodes.ctypes.data_as(C.POINTER(C.c_long))

    def run(libh):
        refs.init_module(libh, n)
        out = []
        for a in (0.01, 0.03, 0.0305, 0.2):
            d = libh.compute_neutrino_power_from_cdm(a, refs.dptr(keff), refs.dptr(P), nm, nk_in, 0)
            out.append((d.nbins, d.norm, np.array([d.logkk[i] for i in range(d.nbins)]), np.array([d.delta_ratio[i] for i in range(d.nbins)])))
            libh.free_d_pow(C.byref(d))
        return out
    for (nb_g, norm_g, lk_g, r_g), (nb_w, norm_w, lk_w, r_w) in zip(run(standin), run(ref)):
        assert nb_g == nb_w == nk_in - 2
        assert norm_g == pytest.approx(norm_w, rel=1e-12)
        np.testing.assert_allclose(lk_g, lk_w, rtol=1e-14)
        np.testing.assert_allclose(r_g, r_w, rtol=1e-10)


def test_get_and_set_nu_state(standin, ref):
    """get_nu_state / set_nu_state (interface_common.c:150-180): the arrays a host checkpoints, moved from the reference
    into the product's host layer, continue to the same delta_nu."""
    n = 32
    g = refs.random_grid(n, seed=12)
    times = (0.01, 0.02, 0.05, 0.1)
    _run(ref, "add_nu_power_to_rhogrid", g, times, False)
    sf, dtot = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    nk, ia = C.c_size_t(), C.c_size_t()
    ref.get_nu_state(C.byref(sf), C.byref(dtot), C.byref(nk), C.byref(ia))
    assert ia.value == len(times) and nk.value > 0
    sf_np = np.array([sf[i] for i in range(ia.value)])
    dt_np = np.array([dtot[i] for i in range(nk.value * ia.value)])
    res = {}
    for name, libh, fn in (("ref", ref, "add_nu_power_to_rhogrid"), ("got", standin, "add_nu_power_to_rhogrid_f64")):
        om, dt = refs.init_module(libh, n)                         # a fresh module, as after a restart
        libh.set_nu_state(refs.dptr(sf_np), refs.dptr(dt_np), nk.value, ia.value, 0)
        assert dt.ia == ia.value and dt.nk == nk.value
        gg = g.copy()
        getattr(libh, fn)(0.15, refs.BOX, gg.ctypes.data_as(C.c_void_p), n, 0, n, 0)
        res[name] = (dt.ia, np.array([dt.delta_nu_last[i] for i in range(dt.nk)]), gg)
    assert res["got"][0] == res["ref"][0] == ia.value + 1
    np.testing.assert_allclose(res["got"][1], res["ref"][1], rtol=1e-10)
    np.testing.assert_allclose(res["got"][2], res["ref"][2], rtol=1e-10)


def test_config1_like_transfer_at_a_0p1_host_flow(standin, ref):
    """BASELINE.json configs[1] in miniature (3 x 0.1 eV, CAMB ics_transfer_0.1.dat, TimeTransfer = 0.1; 32^3 instead of
    256^3): init step, a kept row, a dropped row."""
    n = 32
    tfile = os.path.join(refs.GOLDEN, "camb_ics_transfer_0.1.dat")
    g = refs.random_grid(n, seed=256)
    times = (0.1, 0.12, 0.1205)
    kw = dict(masses=(0.1, 0.1, 0.1), time_transfer=0.1, transfer=tfile)
    want = _run(ref, "add_nu_power_to_rhogrid", g, times, False, **kw)
    got = _run(standin, "add_nu_power_to_rhogrid_f64", g, times, False, **kw)
    assert [w[0] for w in want] == [1, 2, 2]
    for (ia_r, nk_r, g_r, dn_r), (ia_g, nk_g, g_g, dn_g) in zip(want, got):
        assert (ia_r, nk_r) == (ia_g, nk_g)
        np.testing.assert_allclose(dn_g, dn_r, rtol=1e-10, atol=0)
        np.testing.assert_allclose(g_g, g_r, rtol=1e-10, atol=0)


def test_float_grid_host_flow(standin):
    """The float-grid build of the reference (no DOUBLEPRECISION_FFTW) against the product's _f32 entry through the host
    layer: 1e-5 (north-star tolerance for a float grid)."""
    ref_s = refs.ref_lib(False)
    if ref_s is None:
        pytest.skip("oracle/_ref not built")
    n = 32
    g = refs.random_grid(n, seed=9, dtype=np.float32)
    times = (0.01, 0.03, 0.0305, 0.1)
    want = _run(ref_s, "add_nu_power_to_rhogrid", g, times, False)
    got = _run(standin, "add_nu_power_to_rhogrid_f32", g, times, False)
    for (ia_r, nk_r, g_r, dn_r), (ia_g, nk_g, g_g, dn_g) in zip(want, got):
        assert (ia_r, nk_r) == (ia_g, nk_g)
        np.testing.assert_allclose(dn_g, dn_r, rtol=1e-5, atol=0)
        np.testing.assert_allclose(g_g, g_r, rtol=1e-5, atol=0)


def test_total_power_and_output_files_host_flow(standin, ref, tmp_path):
    """compute_total_power_spectrum + save_total_power + save_neutrino_power (interface_gadget.c:114-144,196-224,
    delta_tot_table.c:355-374).  save_total_power's KSPACE_NEUTRINOS_2 switch is a compile-time one in the reference
    (:199-219; oracle/_ref is built without it, as the reference's own Makefile builds); the product decides at run time:
    integrator initialised -> the KSPACE_NEUTRINOS_2 branch (total = get_delta_tot(delta_nu_last, delta_cdm_last)),
    otherwise the plain branch.  Both are checked: the plain one byte for byte against the reference's file, the other
    against the reference's own get_delta_tot on the reference's state."""
    n = 32
    g = refs.random_grid(n, seed=8)
    out = {}
    # (a) no neutrino steps taken: plain branch on both sides
    for name, libh, tot in (("ref", ref, "compute_total_power_spectrum"), ("got", standin, "compute_total_power_spectrum_f64")):
        refs.init_module(libh, n)
        # init_module forgets delta_cdm_last; the reference writes through it unconditionally here (:136): give it one
        keep = (C.c_double * (n // 2))()
        C.c_void_p.in_dll(libh, "delta_cdm_last").value = C.addressof(keep)
        gg = g.copy()
        getattr(libh, tot)(0.02, refs.BOX, gg.ctypes.data_as(C.c_void_p), n, 0, n, 0)
        C.c_void_p.in_dll(libh, "delta_cdm_last").value = None
        d = tmp_path / (name + "_plain")
        d.mkdir()
        assert libh.save_total_power(0.02, 3, str(d).encode()) == 0
        out[name] = (d / "powerspec_tot_003.txt").read_bytes()
    assert len(out["ref"]) > 100 and out["got"] == out["ref"]
    # (b) after two PM steps
    for name, libh, step, tot in (("ref", ref, "add_nu_power_to_rhogrid", "compute_total_power_spectrum"),
                                  ("got", standin, "add_nu_power_to_rhogrid_f64", "compute_total_power_spectrum_f64")):
        om, dt = refs.init_module(libh, n)
        gg = g.copy()
        p = gg.ctypes.data_as(C.c_void_p)
        getattr(libh, step)(0.01, refs.BOX, p, n, 0, n, 0)
        getattr(libh, step)(0.02, refs.BOX, p, n, 0, n, 0)
        d = tmp_path / name
        d.mkdir()
        assert libh.save_neutrino_power(0.02, 7, str(d).encode()) == 0
        getattr(libh, tot)(0.02, refs.BOX, p, n, 0, n, 0)
        assert libh.save_total_power(0.02, 3, str(d).encode()) == 0
        out[name] = ((d / "powerspec_nu_007.txt").read_bytes(), (d / "powerspec_tot_003.txt").read_bytes())
        assert libh.save_neutrino_power(0.02, 8, str(d / "no" / "such" / "dir").encode()) == -1      # fopen failure -> -1
        if name == "ref":
            # what a -DKSPACE_NEUTRINOS_2 build of the reference would have written (interface_gadget.c:199-219)
            last = C.POINTER(C.c_double).in_dll(libh, "delta_cdm_last")
            OmegaNua3 = libh.OmegaNu_nopart(0.02) * 0.02 ** 3
            rows = out[name][1].decode().split("\n")
            want = rows[:3]
            for i in range(dt.nk):
                tot_i = libh.get_delta_tot(dt.delta_nu_last[i], last[i], OmegaNua3, dt.Omeganonu, libh.OmegaNu(1.0), 0.0)
                want.append("%s %s" % (rows[3 + i].split()[0], "%g" % (tot_i * tot_i)))
            want_tot = ("\n".join(want) + "\n").encode()
    assert len(out["ref"][0]) > 100 and out["got"][0] == out["ref"][0]
    assert out["got"][1] == want_tot
