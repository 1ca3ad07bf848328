"""Pins the CPU oracle (oracle/ksn_oracle.c + mini-GSL): against the reference's own known answers
 everywhere, and against the reference sources compiled unmodified (oracle/_ref)
where that library exists.  No GPU needed."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi
from tests import refs

O = refs.orc()
KT = 8.61734e-5 * ((4 / 11.) ** (1 / 3.) * 1.00328) * refs.T_CMB0


def _orc_ps(grid, nrbins, start=0, nslab=None):
    n = grid.shape[1]
    nslab = grid.shape[0] if nslab is None else nslab
    p, k = np.zeros(nrbins), np.zeros(nrbins)
    c = np.zeros(nrbins, dtype=np.int64)
    nret = O.orc_total_powerspectrum(n, grid.ctypes.data_as(C.c_void_p), 1 if grid.dtype == np.float64 else 0, nrbins, start, nslab,
                                     refs.dptr(p), c.ctypes.data_as(capi.c_longlong_p), refs.dptr(k))
    return nret, p, c, k


def test_powerspectrum_kat():
    """powerspectrum_test.c:19-47"""
    for dt in (np.float64, np.float32):
        nret, p, c, k = _orc_ps(refs.kat_grid_4().astype(dt), 15)
        assert nret == 9 and list(c[:9]) == [6, 12, 8, 3, 12, 12, 3, 6, 1]
        assert abs(k[2] - 1.73205) < 1e-5
        assert abs(p[0] - 0.254834) < 1e-5 * 0.04 and abs(p[1] - 0.00212722) < 1e-5 * 0.005 and abs(p[2] - 0.00323766) < 1e-5 * 0.003


def test_specialJ_kat():
    """delta_tot_table_test.c:155-170"""
    J = O.orc_specialJ
    assert J(0, -1, 0) == 1
    for x, v in ((1, 0.2117), (2, 0.0223807), (0.5, 0.614729), (0.3, 0.829763)):
        assert abs(J(x, -1, 0) - v) < 1e-3
    assert abs(J(0, 1, 0) - 0.940437) < 1e-4
    assert abs(J(0.5, 1e-2, 0.5) - 0.614729 / 0.5) < 1e-3
    assert abs(J(0.5, 1, 0.5) - 0.556557 / 0.5) < 1e-4
    assert abs(J(1, 0.1, 0.5) - 0.211662 / 0.5) < 1e-4


def test_fslength_and_omega_kat():
    """delta_tot_table_test.c:183-191, omega_nu_single_test.c (omega_nu(1), nufrac_low)"""
    c = refs.orc_cosmo()
    assert abs(O.orc_fslength(C.byref(c), math.log(0.5), 0.0, 299792.) / 1272.92 / (0.45 / KT) - 1) < 1e-5
    assert abs(O.orc_fslength(C.byref(c), math.log(0.1), math.log(0.5), 299792.) / 5427.8 / (0.6 / KT) - 1) < 1e-5
    assert abs(O.orc_omega_nu(C.byref(c), 1.0) / (0.45 / 93.14 / 0.49) - 1) < 1e-3
    assert abs(O.orc_nufrac_low(1.0) - 0.0595634) < 1e-5 and abs(O.orc_nufrac_low(0.5) - 0.00941738) < 1e-5
    # a^-3 scaling of non-relativistic neutrinos
    assert abs(O.orc_omega_nu(C.byref(c), 0.5) / (8 * O.orc_omega_nu(C.byref(c), 1.0)) - 1) < 5e-3


def test_transfer_kat():
    """transfer_init_test.c:12-37"""
    tl, tt = capi.c_double_p(), capi.c_double_p()
    path = os.path.join(refs.GOLDEN, "ics_transfer_99.dat").encode()
    assert O.orc_transfer_read(path, 512000.0, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt)) == 271
    assert abs(tt[0] - 0.508479) < 1e-6 and abs(tt[30] - 0.0122563) < 1e-6
    assert O.orc_transfer_read(path, 512000.0e3, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt)) == 336


def _orc_resume(masses=(0.15, 0.15, 0.15), hybrid=False, time=0.33333333):
    kk, delta_nu, delta_tot = refs.load_golden_state()
    c = refs.orc_cosmo(masses, hybrid=hybrid)
    c0 = refs.orc_cosmo()
    OmegaNua3 = O.orc_omega_nu(C.byref(c0), 0.01) * 0.01 ** 3
    fnu = OmegaNua3 / (refs.OMEGA0 - O.orc_omega_nu(C.byref(c0), 1.0) + OmegaNua3)
    dcdm = (delta_tot - fnu * delta_nu) / (1 - fnu)
    d = refs.OrcDtot()
    O.orc_dtot_alloc(C.byref(d), len(kk), 0.01, 1.0, refs.OMEGA0, C.byref(c), refs.UNIT_TIME, refs.UNIT_LENGTH)
    assert O.orc_dtot_read(C.byref(d), os.path.join(refs.GOLDEN, "delta_tot_nu.txt").encode()) == 25
    tl, tt = capi.c_double_p(), capi.c_double_p()
    nt = O.orc_transfer_read(os.path.join(refs.GOLDEN, "ics_transfer_99.dat").encode(), refs.BOX, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt))
    O.orc_dtot_init(C.byref(d), len(kk), refs.dptr(kk), refs.dptr(dcdm), tl, tt, nt, time)
    return dict(c=c, d=d, kk=kk, dcdm=dcdm, delta_nu=delta_nu, tl=tl, tt=tt, nt=nt)


def test_get_delta_nu_update_golden():
    """delta_tot_table_test.c:193-228 against testdata/powerspec_nu_004.txt"""
    s = _orc_resume()
    out = np.zeros(len(s["kk"]))
    assert s["d"].ia == 25
    assert O.orc_get_delta_nu_update(C.byref(s["d"]), 0.33333333, len(out), refs.dptr(s["kk"]), refs.dptr(s["dcdm"]), refs.dptr(out), s["tl"], s["tt"], s["nt"]) == 0
    assert s["d"].ia == 25 and np.all(np.abs(out / s["delta_nu"] - 1) < 1e-2)
    s["d"].ia -= 1
    O.orc_get_delta_nu_update(C.byref(s["d"]), 0.33333333, len(out), refs.dptr(s["kk"]), refs.dptr(s["dcdm"]), refs.dptr(out), s["tl"], s["tt"], s["nt"])
    assert s["d"].ia == 25 and np.all(np.abs(out / s["delta_nu"] - 1) < 3e-2)


def test_reproduce_linear_first_steps():
    """test_reproduce_linear (delta_tot_table_test.c:315-363), first 12 of the 99 CAMB steps (the full run is
    the GPU test; the CPU oracle takes ~3 s per late step)."""
    z = np.load(os.path.join(refs.GOLDEN, "camb_linear_steps.npz"))
    c = refs.orc_cosmo()
    d = refs.OrcDtot()
    O.orc_dtot_alloc(C.byref(d), 200, 0.01, 1.0, refs.OMEGA0, C.byref(c), refs.UNIT_TIME, refs.UNIT_LENGTH)
    tl, tt = capi.c_double_p(), capi.c_double_p()
    nt = O.orc_transfer_read(os.path.join(refs.GOLDEN, "camb_ics_transfer_0.01.dat").encode(), 512000.0, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt))
    k0, d0 = np.ascontiguousarray(z["keffs"][0]), np.ascontiguousarray(z["delta_cdm"][0])
    O.orc_dtot_init(C.byref(d), 200, refs.dptr(k0), refs.dptr(d0), tl, tt, nt, 0.01)
    for i in range(12):
        k, dc = np.ascontiguousarray(z["keffs"][i]), np.ascontiguousarray(z["delta_cdm"][i])
        out = np.zeros(200)
        assert O.orc_get_delta_nu_update(C.byref(d), float(z["a"][i]), 200, refs.dptr(k), refs.dptr(dc), refs.dptr(out), tl, tt, nt) == 0
        assert d.ia == i + 1
        acc = 0.05 if i < 8 else 0.02
        assert np.all(np.abs(z["delta_nu_camb"][i] - out) < acc * out)


def test_k3_oracle_matches_numpy_restatement():
    n = 16
    g = refs.random_grid(n, seed=4)
    rng = np.random.default_rng(0)
    logkk = np.sort(rng.uniform(np.log(2 * np.pi / refs.BOX * 1.5), np.log(2 * np.pi / refs.BOX * 11), 9))
    ratio = rng.random(9)
    want = refs.k3_numpy(g, 0, refs.BOX, logkk, ratio, 0.05)
    got = g.copy()
    O.orc_scale_modes(got.ctypes.data_as(C.c_void_p), 1, n, 0, n, refs.BOX, refs.dptr(logkk), refs.dptr(ratio), 9, 0.05)
    np.testing.assert_allclose(got, want, rtol=1e-14)


# ----------------------------------------------------------------------------- against the reference sources
needs_ref = pytest.mark.skipif(refs.ref_lib(True) is None, reason="oracle/_ref not built (no /root/reference on this box)")


@needs_ref
@pytest.mark.parametrize("n,nrbins,dtype", [(8, 4, np.float64), (16, 8, np.float64), (32, 16, np.float64), (48, 100, np.float64), (32, 16, np.float32),
                                            (192, 96, np.float64)])   # 192: the corner mode sits on the last bin edge (1-ulp case)
def test_k1_oracle_equals_reference(n, nrbins, dtype):
    ref = refs.ref_lib(dtype == np.float64)
    g = refs.random_grid(n, seed=n, dtype=dtype)
    r = refs.total_powerspectrum(ref, g, nrbins)
    o = _orc_ps(g, nrbins)
    assert r[0] == o[0] and np.array_equal(r[2], o[2])
    np.testing.assert_allclose(o[1][:o[0]], r[1][:r[0]], rtol=1e-13 if dtype == np.float64 else 1e-6)
    np.testing.assert_allclose(o[3][:o[0]], r[3][:r[0]], rtol=1e-13)
    # ragged slab
    r = refs.total_powerspectrum(ref, np.ascontiguousarray(g[3:7]), nrbins, startslab=3, nslab=4)
    o = _orc_ps(np.ascontiguousarray(g[3:7]), nrbins, 3, 4)
    assert np.array_equal(r[2][:r[0]], o[2][:o[0]])


@needs_ref
@pytest.mark.parametrize("masses,hybrid", [((0.15, 0.15, 0.15), False), ((0.2, 0.1, 0.3), False), ((0.15, 0.15, 0.15), True), ((0.0, 0.0, 0.06), False)])
def test_whole_step_oracle_equals_reference(masses, hybrid):
    """interface_gadget.c:158-194 end to end: same grids in, same grids out (to rounding: the reference
    is built with -ffast-math)."""
    import tests.test_step_gpu as step
    ref = refs.ref_lib(True)
    n = 32
    g = refs.random_grid(n, seed=17)
    times = (0.01, 0.02, 0.0205, 0.05, 0.2, 0.34, 0.345)
    want = step._run(ref, "add_nu_power_to_rhogrid", g, times, False, hybrid=hybrid, masses=masses)
    m = refs.orc_module(n, masses=masses, hybrid=hybrid)
    cur = g.copy()
    for (ia, nk, gref, dnu), a in zip(want, times):
        assert O.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, cur.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
        assert (m.dtot.ia, m.dtot.nk) == (ia, nk)
        np.testing.assert_allclose(np.array([m.dtot.delta_nu_last[i] for i in range(nk)]), dnu, rtol=1e-12)
        np.testing.assert_allclose(cur, gref, rtol=1e-12)
