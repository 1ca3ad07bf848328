"""K2 (the linear-response integral) on the GPU: reference known answers, the reference sources
(oracle/_ref, mini-GSL + exact host hubble_function calls) to 1e-10, and the CAMB fixtures."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi
from tests import refs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def state(gpu):
    om = refs.make_omnu(gpu)
    refs.set_background(gpu, om)
    kk, delta_nu, delta_tot = refs.load_golden_state()
    delta_cdm = refs.golden_delta_cdm(gpu, om, delta_nu, delta_tot)
    transfer = refs.load_transfer(gpu)
    return dict(om=om, kk=kk, delta_nu=delta_nu, delta_cdm=delta_cdm, transfer=transfer)


def test_fslength_device_matches_host_and_kat(gpu, state):
    """test_fslength (delta_tot_table_test.c:183-191) through the device table of 1/(aH)."""
    kT = 8.61734e-5 * ((4 / 11.) ** (1 / 3.) * 1.00328) * refs.T_CMB0
    hub = capi.HUBBLE_FN(lambda a, _u: gpu.hubble_function(a))
    capi.check(gpu.ksn_set_background(hub, None, math.log(0.01) - 0.01, 0.01, 16384))
    lo = np.array([math.log(0.5), math.log(0.1), math.log(0.01), math.log(0.9)])
    out = np.zeros(4)
    capi.check(gpu.ksn_fslength_device(refs.dptr(lo), 4, 0.0, 299792., refs.dptr(out)))
    assert abs(out[0] / 1272.92 / (0.45 / kT) - 1) < 1e-5
    for i in range(4):
        host = gpu.fslength(lo[i], 0.0, 299792.)
        assert abs(out[i] / host - 1) < 1e-11
    capi.check(gpu.ksn_fslength_device(refs.dptr(lo[1:2].copy()), 1, math.log(0.5), 299792., refs.dptr(out)))
    assert abs(out[0] / 5427.8 / (0.6 / kT) - 1) < 1e-5
    gpu.ksn_invalidate_background()


@pytest.mark.parametrize("masses,nkinks", [((0.15, 0.15, 0.15), 1), ((0.2, 0.1, 0.3), 3)])
def test_background_table_patches_the_kinks_of_hubble(gpu, masses, nkinks):
    """H(a) has a slope discontinuity where Omega_nu switches from its spline table to the non-relativistic series
    (omega_nu_single.c:180-199, a = 100 kT/m per distinct mass).  The device table must find each one, cover it with a
    refined patch, and then reproduce the host's fslength ACROSS the kink to rounding (a plain 4-point table is off by
    ~1e-11 there, which is what flipped adaptive-quadrature decisions against the CPU path in long runs)."""
    om = refs.make_omnu(gpu, masses)
    refs.set_background(gpu, om)
    hub = capi.HUBBLE_FN(lambda a, _u: gpu.hubble_function(a))
    capi.check(gpu.ksn_set_background(hub, None, math.log(0.01) - 0.01, 0.01, 16384))
    npatch, flagged = C.c_int(), C.c_int()
    capi.check(gpu.ksn_background_info(C.byref(npatch), C.byref(flagged)))
    assert npatch.value == nkinks and flagged.value == 3 * nkinks       # a kink spoils the three cells whose stencil straddles it
    kT = 8.61734e-5 * ((4 / 11.) ** (1 / 3.) * 1.00328) * refs.T_CMB0
    worst = 0.0
    for m in set(masses):
        a_sw = 100 * kT / m
        lo = np.array([math.log(a_sw) - d for d in (0.3, 0.02, 1e-3, 1e-5)])
        out = np.zeros(len(lo))
        capi.check(gpu.ksn_fslength_device(refs.dptr(lo), len(lo), math.log(a_sw) + 0.01, 299792., refs.dptr(out)))
        for i in range(len(lo)):
            worst = max(worst, abs(out[i] / gpu.fslength(lo[i], math.log(a_sw) + 0.01, 299792.) - 1))
    assert worst < 2e-14, worst
    gpu.ksn_invalidate_background()


def _resume(libh, om, st, time):
    d = refs.new_delta_tot(libh, om, len(st["kk"]))
    libh.read_all_nu_state(C.byref(d), os.path.join(refs.GOLDEN, "delta_tot_nu.txt").encode())
    libh.delta_tot_init(C.byref(d), len(st["kk"]), refs.dptr(st["kk"]), refs.dptr(st["delta_cdm"]), C.byref(st["transfer"]), time)
    return d


def test_get_delta_nu_update_golden(gpu, state):
    """test_get_delta_nu_update (delta_tot_table_test.c:193-228): resume at a=1/3, compare with the saved P_nu."""
    d = _resume(gpu, state["om"], state, 0.33333333)
    assert d.ia == 25
    out = np.zeros(len(state["kk"]))
    gpu.get_delta_nu_update(C.byref(d), 0.33333333, len(out), refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), refs.dptr(out), C.byref(state["transfer"]))
    assert d.ia == 25
    assert np.all(np.abs(out / state["delta_nu"] - 1) < 1e-2)
    d.ia -= 1
    gpu.get_delta_nu_update(C.byref(d), 0.33333333, len(out), refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), refs.dptr(out), C.byref(state["transfer"]))
    assert d.ia == 25
    assert np.all(np.abs(out / state["delta_nu"] - 1) < 3e-2)


@pytest.mark.parametrize("masses,hybrid", [((0.15, 0.15, 0.15), False), ((0.2, 0.1, 0.3), False), ((0.15, 0.15, 0.15), True)])
def test_update_sequence_matches_reference(gpu, state, masses, hybrid):
    """A run of get_delta_nu_update steps (kept rows, dropped rows, repeated a) against the reference
    sources: delta_nu to 1e-10 relative (north-star tolerance), identical row bookkeeping."""
    ref = refs.ref_lib(True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    outs = {}
    for name, libh in (("ref", ref), ("gpu", gpu)):
        om = refs.make_omnu(libh, masses)
        if hybrid:
            m = (C.c_double * 3)(*masses)
            libh.init_hybrid_nu(C.byref(om.hybnu), m, 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
        refs.set_background(libh, om)
        tr = refs.load_transfer(libh)
        st = dict(state, om=om, transfer=tr)
        d = _resume(libh, om, st, 0.3)
        res = []
        for a in (0.3, 0.3005, 0.312, 0.312, 0.33, 0.335, 0.36):
            out = np.zeros(len(state["kk"]))
            libh.get_delta_nu_update(C.byref(d), a, len(out), refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), refs.dptr(out), C.byref(tr))
            res.append((d.ia, out.copy(), np.array([d.delta_tot[5][i] for i in range(d.ia)])))
        outs[name] = res
    for (ia_r, o_r, row_r), (ia_g, o_g, row_g) in zip(outs["ref"], outs["gpu"]):
        assert ia_r == ia_g
        np.testing.assert_allclose(o_g, o_r, rtol=1e-10, atol=0)
        np.testing.assert_allclose(row_g, row_r, rtol=1e-10, atol=0)


def test_reproduce_linear_theory(gpu):
    """test_reproduce_linear (delta_tot_table_test.c:315-363): 99 get_delta_nu_update steps a = 0.01 ... 0.99 fed with
    CAMB's delta_cdm must track CAMB's neutrino transfer function (5 % -> 2 % from step 8 -> 1.2 % from step 22) and
    store one row per step."""
    z = np.load(os.path.join(refs.GOLDEN, "camb_linear_steps.npz"))
    om = refs.make_omnu(gpu)
    refs.set_background(gpu, om)
    tr = refs.load_transfer(gpu, os.path.join(refs.GOLDEN, "camb_ics_transfer_0.01.dat"), box=512000.0)
    d = refs.new_delta_tot(gpu, om, 200)
    k0, d0 = np.ascontiguousarray(z["keffs"][0]), np.ascontiguousarray(z["delta_cdm"][0])
    gpu.delta_tot_init(C.byref(d), 200, refs.dptr(k0), refs.dptr(d0), C.byref(tr), 0.01)
    acc = 0.05
    for i in range(99):
        if i == 8:
            acc = 2e-2
        if i == 22:
            acc = 1.2e-2
        k, dc = np.ascontiguousarray(z["keffs"][i]), np.ascontiguousarray(z["delta_cdm"][i])
        out = np.zeros(200)
        gpu.get_delta_nu_update(C.byref(d), float(z["a"][i]), 200, refs.dptr(k), refs.dptr(dc), refs.dptr(out), C.byref(tr))
        assert d.ia == i + 1
        assert np.all(np.abs(z["delta_nu_camb"][i] - out) < acc * out), i


def test_update_sequence_matches_oracle_port(gpu, state):
    """Same as the reference-sources test but against the oracle restatement, which exists on every box."""
    o = refs.orc()
    masses = (0.2, 0.1, 0.3)
    om = refs.make_omnu(gpu, masses)
    refs.set_background(gpu, om)
    tr = refs.load_transfer(gpu)
    d = _resume(gpu, om, dict(state, om=om, transfer=tr), 0.3)
    oc = refs.orc_cosmo(masses)
    od = refs.OrcDtot()
    n = len(state["kk"])
    o.orc_dtot_alloc(C.byref(od), n, 0.01, 1.0, refs.OMEGA0, C.byref(oc), refs.UNIT_TIME, refs.UNIT_LENGTH)
    o.orc_dtot_read(C.byref(od), os.path.join(refs.GOLDEN, "delta_tot_nu.txt").encode())
    tl, tt = capi.c_double_p(), capi.c_double_p()
    nt = o.orc_transfer_read(os.path.join(refs.GOLDEN, "ics_transfer_99.dat").encode(), refs.BOX, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt))
    o.orc_dtot_init(C.byref(od), n, refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), tl, tt, nt, 0.3)
    for a in (0.3, 0.3005, 0.312, 0.312, 0.33, 0.36):
        g, w = np.zeros(n), np.zeros(n)
        gpu.get_delta_nu_update(C.byref(d), a, n, refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), refs.dptr(g), C.byref(tr))
        assert o.orc_get_delta_nu_update(C.byref(od), a, n, refs.dptr(state["kk"]), refs.dptr(state["delta_cdm"]), refs.dptr(w), tl, tt, nt) == 0
        assert d.ia == od.ia
        np.testing.assert_allclose(g, w, rtol=1e-10, atol=0)


@pytest.mark.parametrize("masses", [(0.1, 0.1, 0.1), (0.2, 0.1, 0.3)])
def test_benchmark_shape_matches_oracle(gpu, masses):
    """K2 at the shape bench.py times (BASELINE configs 3-4): ~780 k bins up to the grid's Nyquist corner, a 98-row
    delta_tot history, hybrid neutrinos on and a > NuPartTime -- the regime where the highest-k bins need > 100
    applications of the 61-point rule.  delta_nu against the oracle restatement, three consecutive steps."""
    o = refs.orc()
    nk, rows = 783, 98
    om = refs.make_omnu(gpu, masses)
    gpu.init_hybrid_nu(C.byref(om.hybnu), (C.c_double * 3)(*masses), 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
    refs.set_background(gpu, om)
    tr = refs.load_transfer(gpu)
    kk = np.geomspace(2 * np.pi / refs.BOX * 1.01, 2 * np.pi / refs.BOX * 1700, nk)
    dcdm = 1e5 * (kk / kk[0]) ** -0.8
    d = refs.new_delta_tot(gpu, om, nk)
    gpu.delta_tot_init(C.byref(d), nk, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.01)
    oc = refs.orc_cosmo(masses, hybrid=True)
    od = refs.OrcDtot()
    o.orc_dtot_alloc(C.byref(od), nk, 0.01, 1.0, refs.OMEGA0, C.byref(oc), refs.UNIT_TIME, refs.UNIT_LENGTH)
    tl, tt = capi.c_double_p(), capi.c_double_p()
    nt = o.orc_transfer_read(os.path.join(refs.GOLDEN, "ics_transfer_99.dat").encode(), refs.BOX, refs.UNIT_LENGTH, refs.UNIT_LENGTH * 1e3, C.byref(tl), C.byref(tt))
    o.orc_dtot_init(C.byref(od), nk, refs.dptr(kk), refs.dptr(dcdm), tl, tt, nt, 0.01)
    for i in range(1, rows):                       # the same synthetic history in both tables: delta_tot grows like a
        d.scalefact[i] = od.scalefact[i] = math.log(0.01 * (i + 1))
        for k in range(nk):
            d.delta_tot[k][i] = d.delta_tot[k][0] * (i + 1)
            od.delta_tot[k * od.namax + i] = od.delta_tot[k * od.namax] * (i + 1)
    d.ia = od.ia = rows
    worst, deepest = 0.0, 0
    for a in (0.981, 0.982, 0.9915):               # two dropped rows, then a kept one
        g, w = np.zeros(nk), np.zeros(nk)
        gpu.get_delta_nu_update(C.byref(d), a, nk, refs.dptr(kk), refs.dptr(dcdm), refs.dptr(g), C.byref(tr))
        deepest = max(deepest, gpu.ksn_last_k2_max_passes())
        assert o.orc_get_delta_nu_update(C.byref(od), a, nk, refs.dptr(kk), refs.dptr(dcdm), refs.dptr(w), tl, tt, nt) == 0
        assert d.ia == od.ia
        worst = max(worst, float(np.max(np.abs(g / w - 1))))
    assert deepest > 60                             # the deep-bisection regime was really entered
    assert worst < 1e-10, worst


def _benchmark_state(gpu, om, nk=783, rows=98):
    """The synthetic 98-row history of test_benchmark_shape_matches_oracle, product side only."""
    tr = refs.load_transfer(gpu)
    kk = np.geomspace(2 * np.pi / refs.BOX * 1.01, 2 * np.pi / refs.BOX * 1700, nk)
    dcdm = 1e5 * (kk / kk[0]) ** -0.8
    d = refs.new_delta_tot(gpu, om, nk)
    gpu.delta_tot_init(C.byref(d), nk, refs.dptr(kk), refs.dptr(dcdm), C.byref(tr), 0.01)
    for i in range(1, rows):
        d.scalefact[i] = math.log(0.01 * (i + 1))
        for k in range(nk):
            d.delta_tot[k][i] = d.delta_tot[k][0] * (i + 1)
    d.ia = rows
    return d, kk, dcdm, tr


@pytest.mark.parametrize("hybrid", [True, False])
@pytest.mark.parametrize("masses", [(0.1, 0.1, 0.1), (0.2, 0.1, 0.3)])
def test_speculative_bisection_is_bit_identical(gpu, masses, hybrid):
    """k2_delta_nu_spec_kernel<M> integrates the halves of the M worst intervals at once and replays QAG's loop over the
    cached results (csrc/ksn_qag_spec.h).  The replay makes the sequential loop's decisions in the sequential order, so
    delta_nu, the per-bin status and the count of rule applications must equal the sequential kernel's BIT FOR BIT, for
    every width, at the benchmark's shape (deepest bins > 100 rule applications with hybrid neutrinos)."""
    om = refs.make_omnu(gpu, masses)
    if hybrid:
        gpu.init_hybrid_nu(C.byref(om.hybnu), (C.c_double * 3)(*masses), 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
    refs.set_background(gpu, om)
    runs = {}
    old = os.environ.get("KSN_K2_SPEC")
    try:
        for width in (1, 2, 3, 4):
            os.environ["KSN_K2_SPEC"] = str(width)
            assert gpu.ksn_k2_spec_width() == width
            d, kk, dcdm, tr = _benchmark_state(gpu, om)
            outs = []
            for a in (0.981, 0.982, 0.9915):
                g = np.zeros(len(kk))
                gpu.get_delta_nu_update(C.byref(d), a, len(kk), refs.dptr(kk), refs.dptr(dcdm), refs.dptr(g), C.byref(tr))
                outs.append((g.copy(), gpu.ksn_last_k2_max_passes(), d.ia))
            runs[width] = outs
    finally:
        if old is None:
            os.environ.pop("KSN_K2_SPEC", None)
        else:
            os.environ["KSN_K2_SPEC"] = old
    if hybrid:
        assert max(p for _, p, _ in runs[1]) > 60
    for width in (2, 3, 4):
        for (g1, p1, ia1), (gm, pm, iam) in zip(runs[1], runs[width]):
            assert ia1 == iam and p1 == pm, (width, p1, pm)
            assert np.array_equal(g1, gm), (width, float(np.max(np.abs(gm / g1 - 1))))


@pytest.mark.parametrize("masses,hybrid", [((0.1, 0.1, 0.1), True), ((0.2, 0.1, 0.3), False)])
def test_prefetched_tables_give_bit_identical_delta_nu(gpu, masses, hybrid):
    """ksn_delta_nu_prefetch computes the a-only part of K2 (free-streaming table, spline factorisations) on a side stream
    ahead of the call -- the PM hook uses it to run them beside K1.  Same kernels on the same inputs: delta_nu must equal the
    un-prefetched call's BIT FOR BIT; a prefetch for other inputs (another a, a changed knot) must simply be ignored."""
    om = refs.make_omnu(gpu, masses)
    if hybrid:
        gpu.init_hybrid_nu(C.byref(om.hybnu), (C.c_double * 3)(*masses), 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
    refs.set_background(gpu, om)
    runs = {}
    for mode in ("plain", "prefetch", "other_a", "other_knot"):
        d, kk, dcdm, tr = _benchmark_state(gpu, om)
        outs = []
        for i, a in enumerate((0.9805, 0.981, 0.982, 0.9915)):
            if mode != "plain" and i > 0:                 # (the first call builds the background table the prefetch needs)
                sf = np.array([d.scalefact[j] for j in range(d.ia)] + [math.log(a)])
                if mode == "other_a":
                    sf[-1] = math.log(a * 1.00001)
                if mode == "other_knot":
                    sf[3] += 1e-12
                capi.check(gpu.ksn_delta_nu_prefetch(a * (1.00001 if mode == "other_a" else 1.0), d.TimeTransfer, d.light, refs.dptr(sf), len(sf), d.namax))
            g = np.zeros(len(kk))
            gpu.get_delta_nu_update(C.byref(d), a, len(kk), refs.dptr(kk), refs.dptr(dcdm), refs.dptr(g), C.byref(tr))
            if i > 0:
                assert gpu.ksn_last_k2_prefetch_used() == (1 if mode == "prefetch" else 0), (mode, i)
            outs.append((g.copy(), d.ia, gpu.ksn_last_k2_evals()))
        runs[mode] = outs
    for mode in ("prefetch", "other_a", "other_knot"):
        for (g0, ia0, ev0), (g1, ia1, ev1) in zip(runs["plain"], runs[mode]):
            assert ia0 == ia1 and ev0 == ev1, mode
            assert np.array_equal(g0, g1), (mode, float(np.max(np.abs(g1 / g0 - 1))))
