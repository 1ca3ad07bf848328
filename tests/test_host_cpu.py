"""Host-side C code of the product (no GPU): Omega_nu tables, CAMB reader, delta_pow lookup, scalar
specialJ/fslength, state files, bin-threshold tables.  Known answers are the reference's own test values
(SURVEY section 4); cross-checks use the CPU oracle as the checker."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi, host
from tests import refs

KT = 8.61734e-5 * ((4 / 11.) ** (1 / 3.) * 1.00328) * refs.T_CMB0


@pytest.fixture(scope="module")
def L(ksn):
    ksn.ksn_set_quiet(1)
    return ksn


def test_omega_nu_known_answers(L):
    """omega_nu_single_test.c: omega_nu(1) = m/93.14/h^2 (1e-3 here incl. the 1.00328 correction), a^-3 scaling,
    degeneracy bookkeeping, nufrac_low, hybrid jump."""
    om = refs.make_omnu(L, (0.15, 0.15, 0.15))
    assert list(om.nu_degeneracies) == [3, 0, 0]
    assert abs(L.get_omega_nu(C.byref(om), 1.0) / (0.45 / 93.14 / 0.49) - 1) < 1e-3
    assert abs(L.get_omega_nu(C.byref(om), 0.5) / (8 * L.get_omega_nu(C.byref(om), 1.0)) - 1) < 5e-3
    om2 = refs.make_omnu(L, (0.2, 0.1, 0.3))
    assert list(om2.nu_degeneracies) == [1, 1, 1]
    tot = sum(L.omega_nu_single(C.byref(om2), 0.3, i) for i in range(3))
    assert abs(tot / L.get_omega_nu(C.byref(om2), 0.3) - 1) < 1e-12
    assert abs(L.nufrac_low(1.0) - 0.0595634) < 1e-5 and abs(L.nufrac_low(0.5) - 0.00941738) < 1e-5
    m = (C.c_double * 3)(0.15, 0.15, 0.15)
    L.init_hybrid_nu(C.byref(om.hybnu), m, 700.0, 299792.458, 0.5, om.kBtnu)
    assert L.particle_nu_fraction(C.byref(om.hybnu), 0.4, 0) == 0
    assert L.particle_nu_fraction(C.byref(om.hybnu), 0.6, 0) == om.hybnu.nufrac_low[0] > 0
    assert L.get_omega_nu_nopart(C.byref(om), 0.6) < L.get_omega_nu(C.byref(om), 0.6)
    # massless neutrinos do not crash and are relativistic
    om0 = refs.make_omnu(L, (0.0, 0.0, 0.0))
    assert L.get_omega_nu(C.byref(om0), 0.1) > 0


def test_omega_nu_matches_oracle_table(L):
    """omega_nu_single_test.c:101-108 (table vs exact to 1e-6): here product vs oracle over 123 scale factors."""
    o = refs.orc()
    for masses in ((0.15, 0.15, 0.15), (0.2, 0.1, 0.3), (0.0, 0.0, 0.06)):
        om = refs.make_omnu(L, masses)
        oc = refs.orc_cosmo(masses)
        for a in np.geomspace(0.01, 1.0, 123):
            assert abs(L.get_omega_nu(C.byref(om), a) / o.orc_omega_nu(C.byref(oc), a) - 1) < 1e-12
            assert abs(L.get_omegag(C.byref(om), a) / o.orc_omegag(C.byref(oc), a) - 1) < 1e-14


def test_transfer_reader(L):
    """transfer_init_test.c:12-37"""
    t = refs.load_transfer(L)
    assert t.NPowerTable == 271 and abs(t.T_nu[0] - 0.508479) < 1e-6 and abs(t.T_nu[30] - 0.0122563) < 1e-6
    t2 = refs.load_transfer(L, box=512000.0e3)
    assert t2.NPowerTable == 336
    assert abs(t2.logk[0] - math.log(0.998289e-05 / 1e3)) < 1e-9


def test_delta_pow_lookup(L):
    """delta_pow_test.c:20-43"""
    kk, delta_nu, delta_tot = refs.load_golden_state()
    logk = np.log(kk)
    ratio = delta_nu / delta_tot
    d = capi.DeltaPow()
    L.init_delta_pow(C.byref(d), refs.dptr(logk), refs.dptr(ratio), len(kk), 1.0)
    for i in range(0, 50, 5):
        assert L.get_dnudcdm_powerspec(C.byref(d), logk[i]) == ratio[i]
    for i in range(50):
        x = (logk[5 * i + 1] + logk[5 * i + 2]) / 2
        y = (ratio[5 * i + 1] + ratio[5 * i + 2]) / 2
        assert abs(L.get_dnudcdm_powerspec(C.byref(d), x) - y) < 5e-3 * y
    assert L.get_dnudcdm_powerspec(C.byref(d), logk[0] - 0.01) == ratio[0]
    assert L.get_dnudcdm_powerspec(C.byref(d), logk[-1] + 0.01) == ratio[-1]
    o = refs.orc()
    for x in np.linspace(logk[0] - 0.5, logk[-1] + 0.5, 301):
        assert L.get_dnudcdm_powerspec(C.byref(d), x) == pytest.approx(o.orc_dnudcdm(refs.dptr(logk), refs.dptr(ratio), len(kk), 1.0, x), rel=1e-14)
    L.free_d_pow(C.byref(d))


def test_scalar_specialJ_and_fslength(L):
    """delta_tot_table_test.c:155-191 through the exported host functions"""
    assert L.specialJ(0, -1, 0) == 1
    assert abs(L.specialJ(1, -1, 0) - 0.2117) < 1e-3 and abs(L.specialJ(0, 1, 0) - 0.940437) < 1e-4
    assert abs(L.specialJ(0.5, 1, 0.5) - 0.556557 / 0.5) < 1e-4 and abs(L.specialJ(1, 0.1, 0.5) - 0.211662 / 0.5) < 1e-4
    om = refs.make_omnu(L)
    refs.set_background(L, om)
    assert abs(L.fslength(math.log(0.5), 0.0, 299792.) / 1272.92 / (0.45 / KT) - 1) < 1e-5
    assert abs(L.fslength(math.log(0.1), math.log(0.5), 299792.) / 5427.8 / (0.6 / KT) - 1) < 1e-5
    assert L.fslength(0.0, -1.0, 299792.) == 0
    o = refs.orc()
    oc = refs.orc_cosmo()
    for lo, hi in ((-4.6, -0.01), (-2.0, -1.0), (-0.3, 0.0)):
        assert abs(L.fslength(lo, hi, 299792.) / o.orc_fslength(C.byref(oc), lo, hi, 299792.) - 1) < 1e-13


def test_state_file_roundtrip(L, tmp_path):
    """test_save_resume (delta_tot_table_test.c:84-124) on a scratch copy; also get/set_nu_state."""
    om = refs.make_omnu(L)
    d = refs.new_delta_tot(L, om, 300)
    L.read_all_nu_state(C.byref(d), os.path.join(refs.GOLDEN, "delta_tot_nu.txt").encode())
    assert d.ia == 25 and abs(d.scalefact[0] / math.log(0.01) - 1) < 1e-5
    for i in range(1, d.ia):
        assert d.scalefact[i] > d.scalefact[i - 1]
    path = str(tmp_path / "delta_tot_nu.txt").encode()
    L.save_all_nu_state(C.byref(d), path)
    L.save_all_nu_state(C.byref(d), path)                 # second save rotates the first to .bak
    assert os.path.exists(path.decode() + ".bak")
    d2 = refs.new_delta_tot(L, om, 300)
    L.read_all_nu_state(C.byref(d2), path)
    assert (d2.ia, d2.nk) == (d.ia, d.nk)
    L.save_all_nu_state(C.byref(d2), path)
    d3 = refs.new_delta_tot(L, om, 300)
    L.read_all_nu_state(C.byref(d3), path)
    for i in range(d.ia):
        assert d2.scalefact[i] == d3.scalefact[i]
        for k in range(0, d.nk, 7):
            assert d2.delta_tot[k][i] == d3.delta_tot[k][i]     # %le text is idempotent after the first write
    assert d.delta_tot[0] and ((C.addressof(d.delta_tot[0].contents) - C.addressof(d.scalefact.contents)) == 8 * d.namax)


def test_set_kspace_vars(L):
    """interface_gadget.c:25-58: nine tags with Gadget-3 type ids (REAL=1, STRING=2, INT=3)."""
    tags = ((C.c_char * 50) * 20)()
    addr = (C.c_void_p * 20)()
    ids = (C.c_int * 20)()
    nt = L.set_kspace_vars(tags, addr, ids, 2)
    assert nt == 11
    names = [tags[i].value.decode() for i in range(2, nt)]
    assert names == ["KspaceTransferFunction", "TimeTransfer", "InputSpectrum_UnitLength_in_cm", "MNue", "MNum", "MNut", "HybridNeutrinosOn", "Vcrit", "NuPartTime"]
    assert list(ids[2:nt]) == [2, 1, 1, 1, 1, 1, 3, 1, 1]
    assert addr[3] == C.addressof(capi.kspace_params()) + capi.KspaceParams.TimeTransfer.offset


@pytest.mark.parametrize("n,nrbins", [(4, 15), (16, 8), (64, 32), (128, 64), (256, 128), (96, 200), (192, 96), (5, 4), (9, 8), (15, 7), (33, 16)])
def test_bin_thresholds_reproduce_reference_counts(L, n, nrbins):
    """The host-libm integer thresholds K1 searches on the device give exactly the reference's mode counts
    (checked with the oracle's K1 on a constant grid, whose counts are data independent, and -- where the reference
    sources are compiled here -- with the reference binary itself; 192 is a size where the corner mode's bin depends on
    how the -ffast-math build evaluates floor(binsperunit*log(kk))).  Odd sizes: the reference counts the z = dims/2 column
    once although it has a distinct conjugate there (powerspectrum.c:70-78), so its counts sum to less than n^3 - 1; the
    product follows it."""
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw)) == 0
    t = np.array([thr[i] for i in range(nrbins)], dtype=np.int64)
    assert t[0] == 0 and np.all(np.diff(t) >= 0)
    k = np.fft.fftfreq(n, 1.0 / n).round().astype(np.int64)
    kz = np.arange(n // 2 + 1)
    k2 = (k[:, None, None] ** 2 + k[None, :, None] ** 2 + kz[None, None, :] ** 2)
    mult = np.where((kz == 0) | (kz == n // 2), 1, 2)[None, None, :] * np.ones_like(k2)
    b = np.searchsorted(t, k2, side="right") - 1
    cnt = np.bincount(b[k2 > 0], weights=mult[k2 > 0], minlength=nrbins).astype(np.int64)
    g = np.ones((n, n, n // 2 + 1, 2))
    o = refs.orc()
    p, kk = np.zeros(nrbins), np.zeros(nrbins)
    c = np.zeros(nrbins, dtype=np.int64)
    m2 = C.c_double()
    o.orc_powerspectrum_sums(n, g.ctypes.data_as(C.c_void_p), 1, nrbins, 0, n, refs.dptr(p), refs.dptr(kk), c.ctypes.data_as(capi.c_longlong_p), C.byref(m2))
    assert np.array_equal(cnt, c) and (n % 2 or c.sum() == n ** 3 - 1)
    ref = refs.ref_lib(True)
    if ref is not None and n <= 192:
        r_n, _, r_c, _ = refs.total_powerspectrum(ref, g, nrbins)
        assert r_n == np.count_nonzero(cnt) and np.array_equal(r_c[:r_n], cnt[cnt > 0])
    for q in (1, n // 4, n // 2):
        assert iw[q] == pytest.approx(math.pi * q / (n * math.sin(math.pi * q / n)), rel=1e-15)


def test_finish_powerspectrum_matches_oracle(L):
    rng = np.random.default_rng(3)
    nb = 40
    p, k = rng.random(nb) * 100, rng.random(nb) * 50
    c = rng.integers(0, 5, nb).astype(np.int64)
    n1, p1, c1, k1 = host.finish_powerspectrum(p, k, c, 7.5)
    o = refs.orc()
    p2, k2, c2 = p.copy(), k.copy(), c.copy()
    n2 = o.orc_powerspectrum_finish(nb, 7.5, refs.dptr(p2), c2.ctypes.data_as(capi.c_longlong_p), refs.dptr(k2))
    assert n1 == n2 == int((c > 0).sum())
    np.testing.assert_array_equal(p1[:n1], p2[:n2])
    np.testing.assert_array_equal(k1[:n1], k2[:n2])
    np.testing.assert_array_equal(c1[:n1], c2[:n2])


def test_slab_partition():
    for n, r in ((2048, 8), (2048, 3), (256, 7), (4, 8)):
        slabs = host.slab_partition(n, r)
        assert sum(s.count for s in slabs) == n and slabs[0].start == 0
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert b.start == a.start + a.count and a.count - b.count in (0, 1)
    assert host.modes_in_slab(2048, host.Slab(0, 2048)) == 4299161600


def test_k1_tile_plan_for_the_named_grids(monkeypatch):
    """The tile shape K1 takes is host arithmetic over the shared-memory budget (227 KB opt-in, 148 SMs on B200): pin it
    for the BASELINE grids -- 2048: eight warps, 17 modes per lane, all 1024 bins in shared memory; 4096: the bin window
    (bins >= 1024 of 2048 in shared memory) is what keeps eight warps, six without it."""
    import ctypes as C
    from kspace_neutrinos_b200 import capi
    L = capi.lib()

    def plan(n):
        v = [C.c_int() for _ in range(5)]
        assert L.ksn_k1_tile_plan(n, n // 2, 232448, 148, *[C.byref(x) for x in v]) == 1
        return tuple(x.value for x in v)

    for k in ("KSN_K1_WIN", "KSN_K1_TILE"):
        monkeypatch.delenv(k, raising=False)
    assert plan(2048) == (8, 17, 2, 2, 0)
    assert plan(4096) == (8, 17, 2, 4, 1024)
    assert plan(1024)[4] == 0 and plan(256)[4] == 0            # never a window where everything fits
    monkeypatch.setenv("KSN_K1_WIN", "0")
    assert plan(4096) == (6, 17, 2, 4, 0)
    monkeypatch.setenv("KSN_K1_WIN", "2")                      # the tests' forced quarter window
    assert plan(256) == (8, 9, 4, 1, 64)
    monkeypatch.delenv("KSN_K1_WIN")

    def plan_f32(n):
        v = [C.c_int() for _ in range(5)]
        assert L.ksn_k1_tile_plan_ex(n, n // 2, 232448, 148, 4, *[C.byref(x) for x in v]) == 1
        return tuple(x.value for x in v)

    # float rows: issue-bound, so eight warps and the longest chunk -- a whole 2048 float row as ONE tile (compile-time
    # chunk of 33; measured 5.57 TB/s at 2048^3 against 4.10 for 15 warps x 9 modes, profiles/r2_optin_timings.txt)
    assert plan_f32(2048) == (8, 33, 2, 1, 0)
    assert plan_f32(1024) == (8, 17, 4, 1, 0)
    assert plan_f32(4096) == (8, 17, 4, 4, 1056)
    monkeypatch.setenv("KSN_K1_TILE", "15,9,2")
    assert plan_f32(2048) == (15, 9, 2, 4, 0)
    monkeypatch.setenv("KSN_K1_TILE", "4,9,3")
    assert plan(256)[:3] == (4, 9, 3)


def _golden_geometry(n):
    g = np.load(os.path.join(refs.GOLDEN, "geometry_counts.npz"))
    return g[f"count_{n}"], g[f"keffsum_{n}"]


@pytest.mark.parametrize("n", [1024, 2048, 4096])
def test_bin_thresholds_at_the_baseline_sizes(L, n):
    """At BASELINE.json's PMGRIDs the reference cannot be swept here; tests/golden/geometry_counts.npz holds its data-independent
    outputs (tools/make_golden_geometry.py: integer multiplicity histogram + the reference's bin expression, checked
    against the compiled reference at the sizes it can do).  The thresholds K1 searches on the device must sit exactly on
    the bin edges of that expression (this machine's libm), and the golden vector must be consistent with them."""
    nrbins = n // 2
    cnt, ksum = _golden_geometry(n)
    assert cnt.sum() == n ** 3 - 1 and len(cnt) == nrbins
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw)) == 0
    t = np.array([thr[i] for i in range(nrbins)], dtype=np.int64)
    libm = C.CDLL("libm.so.6")
    libm.log.restype = C.c_double
    libm.log.argtypes = [C.c_double]
    hb = 0.5 * ((nrbins - 1) / libm.log(n * 0.8660254037844386))
    k2max = 3 * (n // 2) ** 2
    for b in range(1, nrbins):
        if t[b] > k2max:
            assert cnt[b:].sum() == 0 or t[b] == k2max + 1
            continue
        assert math.floor(hb * libm.log(float(t[b]))) >= b
        assert math.floor(hb * libm.log(float(t[b] - 1))) < b if t[b] > 1 else True
    # an empty golden bin <=> no integer k2 = a^2+b^2+c^2 (a, b, c <= N/2) between its thresholds; cheap necessary check:
    # bins whose threshold interval is empty must be empty in the golden vector
    width = np.diff(np.append(t, k2max + 1))
    assert np.all(cnt[width <= 0] == 0)
    assert cnt[np.searchsorted(t, k2max, side="right") - 1] >= 1          # the corner mode's bin
    assert np.all(ksum[cnt > 0] / cnt[cnt > 0] >= np.sqrt(np.maximum(t[cnt > 0], 1)) * (1 - 1e-12))
    assert np.all(ksum[cnt > 0] / cnt[cnt > 0] < np.sqrt(t[cnt > 0] + width[cnt > 0].astype(float)))


@pytest.mark.parametrize("n", [1024, 2048])
def test_golden_counts_from_thresholds(L, n):
    """PMGRID = 1024 and 2048 (BASELINE configs[2], [3]): mode counts recomputed here from the product's thresholds and a
    numpy multiplicity histogram equal the golden vector bin for bin."""
    nrbins, H = n // 2, n // 2
    cnt, ksum = _golden_geometry(n)
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw)) == 0
    t = np.array([thr[i] for i in range(nrbins)], dtype=np.int64)
    w1 = np.where((np.arange(H + 1) == 0) | (np.arange(H + 1) == H), 1, 2).astype(np.int64)
    sq = np.arange(H + 1, dtype=np.int64) ** 2
    ab = (sq[:, None] + sq[None, :]).ravel()
    wab = (w1[:, None] * w1[None, :]).ravel()
    r2 = np.bincount(ab, weights=wab, minlength=2 * H * H + 1).astype(np.int64)     # pairs (i, j) per kx^2+ky^2
    nz = np.nonzero(r2)[0]
    got = np.zeros(nrbins, dtype=np.int64)
    for c in range(H + 1):
        b = np.searchsorted(t, nz + c * c, side="right") - 1
        add = np.bincount(b, weights=r2[nz] * w1[c], minlength=nrbins).astype(np.int64)
        if c == 0:
            add[0] -= 1                                                 # k = 0 is not a mode (powerspectrum.c:65)
        got += add
    assert np.array_equal(got, cnt)
