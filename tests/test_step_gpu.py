"""The whole per-PM-step entry, add_nu_power_to_rhogrid, against the reference sources on identical
grids: corrected grid, P(k) inputs and integrator state (BASELINE.json north-star tolerances)."""
import ctypes as C

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi
from tests import refs

pytestmark = pytest.mark.gpu


def _run(libh, fn, grid0, times, on_device, **kw):
    n = grid0.shape[1]
    om, dt = refs.init_module(libh, n, **kw)
    g = grid0.copy()
    out = []
    dev = refs.DeviceBuffer(libh, g) if on_device else None
    for a in times:
        p = dev.ptr if dev else g.ctypes.data_as(C.c_void_p)
        getattr(libh, fn)(a, refs.BOX, p, n, 0, n, 0)
        cur = dev.download(g) if dev else g.copy()
        last = np.array([dt.delta_nu_last[i] for i in range(dt.nk)])
        out.append((dt.ia, dt.nk, cur, last))
    if dev:
        dev.free()
    return out


@pytest.mark.parametrize("n,hybrid,masses,on_device", [
    (32, False, (0.15, 0.15, 0.15), True),
    (64, False, (0.1, 0.1, 0.1), False),
    (64, True, (0.15, 0.15, 0.15), True),
    (48, False, (0.2, 0.1, 0.3), True),
])
def test_add_nu_power_matches_reference(gpu, n, hybrid, masses, on_device):
    ref = refs.ref_lib(True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    g = refs.random_grid(n, seed=7 + n)
    times = (0.01, 0.02, 0.0205, 0.05, 0.2, 0.34, 0.345, 0.5)
    want = _run(ref, "add_nu_power_to_rhogrid", g, times, False, hybrid=hybrid, masses=masses)
    got = _run(gpu, "add_nu_power_to_rhogrid_f64", g, times, on_device, hybrid=hybrid, masses=masses)
    for (ia_r, nk_r, g_r, dn_r), (ia_g, nk_g, g_g, dn_g) in zip(want, got):
        assert (ia_r, nk_r) == (ia_g, nk_g)
        np.testing.assert_allclose(dn_g, dn_r, rtol=1e-10, atol=0)
        np.testing.assert_allclose(g_g, g_r, rtol=1e-10, atol=0)
        assert not np.array_equal(g_g, g)


def test_add_nu_power_matches_oracle_port(gpu):
    """Whole step against the oracle restatement (available on every box), float grid included."""
    o = refs.orc()
    n = 32
    for dtype, tol, fn in ((np.float64, 1e-10, "add_nu_power_to_rhogrid_f64"), (np.float32, 1e-5, "add_nu_power_to_rhogrid_f32")):
        g = refs.random_grid(n, seed=5, dtype=dtype)
        times = (0.01, 0.03, 0.0305, 0.1)
        got = _run(gpu, fn, g, times, True, masses=(0.1, 0.1, 0.1))
        m = refs.orc_module(n, masses=(0.1, 0.1, 0.1))
        cur = g.copy()
        for (ia, nk, gg, dnu), a in zip(got, times):
            assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, cur.ctypes.data_as(C.c_void_p), 1 if dtype == np.float64 else 0, n, 0, n) == 0
            assert (m.dtot.ia, m.dtot.nk) == (ia, nk)
            np.testing.assert_allclose(dnu, np.array([m.dtot.delta_nu_last[i] for i in range(nk)]), rtol=tol)
            np.testing.assert_allclose(gg, cur, rtol=tol, atol=0)


def test_config1_256_cube_transfer_at_a_0p1(gpu):
    """BASELINE.json configs[1]: PMGRID=256^3 double, 3 x 0.1 eV, CAMB ics_transfer_0.1.dat (TimeTransfer = 0.1), against the
    oracle port on the same bytes: three PM steps (init, kept row, dropped row)."""
    import os
    o = refs.orc()
    n = 256
    tfile = os.path.join(refs.GOLDEN, "camb_ics_transfer_0.1.dat")
    g = refs.random_grid(n, seed=256)
    times = (0.1, 0.12, 0.1205)
    got = _run(gpu, "add_nu_power_to_rhogrid_f64", g, times, True, masses=(0.1, 0.1, 0.1), time_transfer=0.1, transfer=tfile)
    m = refs.orc_module(n, masses=(0.1, 0.1, 0.1), time_transfer=0.1, transfer=tfile)
    cur = g.copy()
    for (ia, nk, gg, dnu), a in zip(got, times):
        assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, cur.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
        assert (m.dtot.ia, m.dtot.nk) == (ia, nk)
        np.testing.assert_allclose(dnu, np.array([m.dtot.delta_nu_last[i] for i in range(nk)]), rtol=1e-10)
        np.testing.assert_allclose(gg, cur, rtol=1e-10, atol=0)


def test_full_size_properties_1024(gpu):
    """At a BASELINE size the CPU oracle cannot sweep in test time (1024^3, 8.6 GB): properties that do not need it.
    One whole step on a device-resident synthetic grid: mode counts add up to N^3-1; the k=0 element is untouched;
    every other mode is scaled by a factor in (1, 1 + prefac] that depends on |k| only (checked on Hermitian-equivalent
    and permuted wave vectors)."""
    from kspace_neutrinos_b200 import host
    n = 1024
    slab = host.Slab(0, n)
    grid = host.DeviceGrid(n, slab)
    grid.fill_synthetic(seed=7, slope=-1.0)
    sim = host.KspaceNeutrinos(host.Cosmology(transfer_file=refs.default_transfer_file(), mnu=(0.1, 0.1, 0.1)), n)
    L = n // 2 + 1

    def rows(i, js):      # a few rows of plane i, copied back
        out = {}
        for j in js:
            buf = np.empty((L, 2))
            off = ((i * n + j) * L) * 16
            capi.check(gpu.ksn_memcpy_d2h(buf.ctypes.data_as(C.c_void_p), C.c_void_p(grid.ptr.value + off), buf.nbytes))
            out[j] = buf
        return out
    probes = [(0, (0, 3, n - 3)), (3, (0,)), (5, (7,)), (7, (5,)), (n - 5, (n - 7,))]
    before = {i: rows(i, js) for i, js in probes}
    sim.add_nu_power_to_rhogrid(0.01, grid.ptr, slab)
    sim.add_nu_power_to_rhogrid(0.02, grid.ptr, slab)
    after = {i: rows(i, js) for i, js in probes}
    grid.free()
    assert sim.state.nk > 400 and sim.state.ia == 2
    f = {(i, j): after[i][j][:, 0] / before[i][j][:, 0] for i, js in probes for j in js}
    assert f[(0, 0)][0] == 1.0                                      # F(0,0,0) untouched
    for key, v in f.items():
        vv = v[1:] if key == (0, 0) else v
        assert np.all(vv > 1.0) and np.all(vv < 1.2)
    np.testing.assert_allclose(f[(0, 3)], f[(0, n - 3)], rtol=1e-13)   # (0, 3, z) and (0, -3, z): same |k|
    np.testing.assert_allclose(f[(0, 3)], f[(3, 0)], rtol=1e-13)       # (0, 3, z) and (3, 0, z)
    np.testing.assert_allclose(f[(5, 7)], f[(7, 5)], rtol=1e-13)       # (5, 7, z) and (7, 5, z)
    np.testing.assert_allclose(f[(5, 7)], f[(n - 5, n - 7)], rtol=1e-13)


def test_compute_neutrino_power_from_cdm_matches_reference(gpu):
    """The MP-Gadget style entry (interface_common.c:106-123): host supplies P(k); empty bins are dropped; only K2 runs
    on the GPU.  Against the reference sources where available, else against the oracle port's integrator."""
    n = 64
    kk, delta_nu, delta_tot = refs.load_golden_state()
    nk_in = n // 2
    keff = np.ascontiguousarray(kk[::9][:nk_in])
    rng = np.random.default_rng(1)
    P = (1e5 * (keff / keff[0]) ** -0.7) ** 2
    nmodes = np.ones(nk_in, dtype=np.int64)
    nmodes[[3, 17]] = 0                                        # two empty bins
    nm = nmodes.ctypes.data_as(C.POINTER(C.c_long))

    def run(libh):
        refs.init_module(libh, n, masses=(0.15, 0.15, 0.15))
        out = []
        for a in (0.01, 0.03, 0.0305, 0.2):
            d = libh.compute_neutrino_power_from_cdm(a, refs.dptr(keff), refs.dptr(P), nm, nk_in, 0)
            out.append((d.nbins, d.norm, np.array([d.logkk[i] for i in range(d.nbins)]), np.array([d.delta_ratio[i] for i in range(d.nbins)])))
            libh.free_d_pow(C.byref(d))
        return out
    got = run(gpu)
    ref = refs.ref_lib(True)
    if ref is not None:
        want = run(ref)
    else:
        o = refs.orc()
        m = refs.orc_module(n, masses=(0.15, 0.15, 0.15))
        keep = nmodes > 0
        kz, dz = np.ascontiguousarray(keff[keep]), np.ascontiguousarray(np.sqrt(P[keep]))
        want = []
        for a in (0.01, 0.03, 0.0305, 0.2):
            dn = np.zeros(len(kz))
            assert o.orc_get_delta_nu_update(C.byref(m.dtot), a, len(kz), refs.dptr(kz), refs.dptr(dz), refs.dptr(dn), m.t_logk, m.t_tnu, m.nt) == 0
            nop = o.orc_omega_nu_nopart(C.byref(m.cosmo), a)
            hyb = o.orc_omega_nu(C.byref(m.cosmo), a) - nop
            want.append((len(kz), nop / (m.dtot.Omeganonu / a ** 3 + hyb), np.log(kz), dn / dz))
    for (nb_g, norm_g, lk_g, r_g), (nb_w, norm_w, lk_w, r_w) in zip(got, want):
        assert nb_g == nb_w == nk_in - 2
        assert norm_g == pytest.approx(norm_w, rel=1e-12)
        np.testing.assert_allclose(lk_g, lk_w, rtol=1e-14)
        np.testing.assert_allclose(r_g, r_w, rtol=1e-10)


def test_total_power_path_and_output_files(gpu, tmp_path):
    """compute_total_power_spectrum + save_total_power (interface_gadget.c:114-144,196-224) and save_neutrino_power:
    file contents against the oracle's numbers through the same '%g' formatting."""
    o = refs.orc()
    n = 32
    g = refs.random_grid(n, seed=8)
    refs.init_module(gpu, n, masses=(0.15, 0.15, 0.15))
    d = refs.DeviceBuffer(gpu, g)
    gpu.add_nu_power_to_rhogrid_f64(0.01, refs.BOX, d.ptr, n, 0, n, 0)
    assert gpu.save_neutrino_power(0.01, 7, str(tmp_path).encode()) == 0
    lines = open(tmp_path / "powerspec_nu_007.txt").read().split("\n")
    assert lines[0] == "# k P_nu(k)" and lines[1] == "# a = 0.01"
    dt = capi.global_delta_tot_table()
    assert lines[2] == f"# nbins = {dt.nk}"
    k0, p0 = lines[3].split()
    assert k0 == "%g" % dt.wavenum[0] and p0 == "%g" % (dt.delta_nu_last[0] ** 2)
    # total power of the (already corrected) grid, no-neutrino path
    gpu.compute_total_power_spectrum_f64(0.01, refs.BOX, d.ptr, n, 0, n, 0)
    cur = d.download(g)
    d.free()
    assert gpu.save_total_power(0.01, 3, str(tmp_path).encode()) == 0
    rows = [l.split() for l in open(tmp_path / "powerspec_tot_003.txt").read().strip().split("\n")[3:]]
    nb = n // 2
    p, k = np.zeros(nb), np.zeros(nb)
    c = np.zeros(nb, dtype=np.int64)
    nret = o.orc_total_powerspectrum(n, cur.ctypes.data_as(C.c_void_p), 1, nb, 0, n, refs.dptr(p), c.ctypes.data_as(capi.c_longlong_p), refs.dptr(k))
    assert len(rows) == nret
    np.testing.assert_allclose([float(r[0]) for r in rows], k[:nret] * 2 * np.pi / refs.BOX, rtol=6e-6)    # %g keeps 6 significant digits


@pytest.mark.parametrize("n,hybrid,masses", [(32, False, (0.1, 0.1, 0.1)), (32, True, (0.2, 0.1, 0.3)), (64, True, (0.06, 0.1, 0.15))])
def test_hundred_step_run_tracks_oracle(gpu, n, hybrid, masses):
    """A whole simulated run, a = 0.01 ... 1.0 in 110 PM steps (kept and dropped rows, the history growing to ~100 rows, the
    hybrid switch-on at a = 0.333): after EVERY step delta_nu and the row bookkeeping must agree with the CPU oracle to the
    north-star tolerance.  ~10^5 adaptive-quadrature decisions are replayed; one flipped decision would show up as ~1e-7."""
    o = refs.orc()
    g = refs.random_grid(n, seed=77)
    refs.init_module(gpu, n, masses=masses, hybrid=hybrid)
    dt = capi.global_delta_tot_table()
    m = refs.orc_module(n, masses=masses, hybrid=hybrid)
    dev = refs.DeviceBuffer(gpu, g)
    cur = g.copy()
    worst = 0.0
    # kept rows every 0.0105 (namax = 101 rows is never exceeded: the reference has no bound check), plus dropped in-between steps
    times = [0.01]
    for i in range(1, 95):
        times.append(0.01 + 0.0105 * i)
        if i % 6 == 2:
            times.append(0.01 + 0.0105 * i + 0.004)
    for a in times:
        gpu.add_nu_power_to_rhogrid_f64(a, refs.BOX, dev.ptr, n, 0, n, 0)
        assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, cur.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
        assert (dt.ia, dt.nk) == (m.dtot.ia, m.dtot.nk)
        got = np.array([dt.delta_nu_last[i] for i in range(dt.nk)])
        want = np.array([m.dtot.delta_nu_last[i] for i in range(dt.nk)])
        worst = max(worst, float(np.max(np.abs(got / want - 1))))
    final = dev.download(g)
    dev.free()
    assert dt.ia > 90
    assert worst < 1e-10, worst
    np.testing.assert_allclose(final, cur, rtol=1e-9, atol=0)       # 110 multiplicative steps accumulate rounding


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_streaming_plan_for_slabs_that_do_not_fit_hbm(gpu, dtype, tol, monkeypatch):
    """A host slab larger than what may stay resident (forced here with KSN_STAGE_MAX_MB; in production: larger than free
    HBM) goes through a three-chunk ring twice -- K1 on the chunks as they land, then upload/K3/download per chunk with the
    uploads overlapping the downloads.  Same result as the resident plan, slab starting at plane 0 (total_mass2 comes from
    a chunk the ring recycles) and not."""
    n = 96                                                     # 96 planes of 75 KB (double): 1 MB chunks hold 13 -> 8 chunks
    g = refs.random_grid(n, seed=31, dtype=dtype)
    fn = "add_nu_power_to_rhogrid_f64" if dtype == np.float64 else "add_nu_power_to_rhogrid_f32"
    outs = []
    for streaming in (False, True):
        if streaming:
            monkeypatch.setenv("KSN_STAGE_MAX_MB", "0")
            monkeypatch.setenv("KSN_STAGE_CHUNK_MB", "1")       # 1 MB: 8 chunks (double) or 4 (float) -> the ring of 3 is recycled
        refs.init_module(gpu, n, masses=(0.15, 0.15, 0.15))
        dt = capi.global_delta_tot_table()
        cur = g.copy()
        for a in (0.01, 0.02, 0.03):
            getattr(gpu, fn)(a, refs.BOX, cur.ctypes.data_as(C.c_void_p), n, 0, n, 0)
        part = g[20:93].copy()                                     # ragged slab that does not hold plane 0: K3 alone (a COPY: a slice of g is a view)
        logkk = np.log(np.array([dt.wavenum[i] for i in range(dt.nk)]))
        ratio = np.linspace(0.5, 0.1, dt.nk)
        capi.check(gpu.ksn_scale_modes(part.ctypes.data_as(C.c_void_p), g.dtype.itemsize, n, 20, 73, refs.BOX,
                                       refs.dptr(logkk), refs.dptr(ratio), dt.nk, 0.01))
        outs.append((cur, part, np.array([dt.delta_nu_last[i] for i in range(dt.nk)])))
    np.testing.assert_allclose(outs[1][2], outs[0][2], rtol=1e-12)
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=tol)
    np.testing.assert_array_equal(outs[1][1], outs[0][1])
