"""K3 (the mode-scaling pass of add_nu_power_to_rhogrid) against a numpy restatement of
interface_gadget.c:163-188 / delta_pow.c:19-37."""
import ctypes as C

import numpy as np
import pytest

from tests import refs

pytestmark = pytest.mark.gpu


def _table(n, box, nk=40, seed=0):
    rng = np.random.default_rng(seed)
    kmin, kmax = 2 * np.pi / box, np.sqrt(3) * (n / 2) * 2 * np.pi / box
    # knots strictly inside the mode range so both clamps are exercised
    logkk = np.sort(rng.uniform(np.log(kmin * 1.3), np.log(kmax * 0.8), nk))
    ratio = 0.2 + 0.6 * rng.random(nk)
    return logkk, ratio, 0.0123


@pytest.mark.parametrize("n,dtype,tol", [(4, np.float64, 1e-10), (16, np.float64, 1e-10), (64, np.float64, 1e-10), (96, np.float64, 1e-10), (64, np.float32, 1e-5)])
def test_scale_modes_matches_numpy(gpu, n, dtype, tol):
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    g = refs.random_grid(n, seed=n + 1, dtype=dtype)
    logkk, ratio, norm = _table(n, box, nk=max(3, n // 2))
    want = refs.k3_numpy(g, 0, box, logkk, ratio, norm)
    # device-resident
    d = refs.DeviceBuffer(gpu, g)
    capi.check(gpu.ksn_scale_modes(d.ptr, g.dtype.itemsize, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
    got = d.download(g)
    d.free()
    np.testing.assert_allclose(got, want, rtol=tol, atol=0)
    # host-resident (staged), ragged slab
    a, b = n // 4, n // 4 + max(1, n // 3)
    sub = np.ascontiguousarray(g[a:b])
    want_sub = refs.k3_numpy(sub, a, box, logkk, ratio, norm)
    capi.check(gpu.ksn_scale_modes(sub.ctypes.data_as(C.c_void_p), g.dtype.itemsize, n, a, b - a, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
    np.testing.assert_allclose(sub, want_sub, rtol=tol, atol=0)


def test_origin_untouched_and_clamps(gpu):
    from kspace_neutrinos_b200 import capi
    n, box = 32, refs.BOX
    g = refs.random_grid(n, seed=2)
    logkk = np.log(np.array([3.0, 5.0, 9.0]) * 2 * np.pi / box)     # knots at |n| = 3, 5, 9
    ratio = np.array([0.5, 0.25, 1.0])
    d = refs.DeviceBuffer(gpu, g)
    capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), 3, 0.1))
    got = d.download(g)
    d.free()
    assert np.array_equal(got[0, 0, 0], g[0, 0, 0])                      # k=0 is skipped
    np.testing.assert_allclose(got[1, 0, 0], g[1, 0, 0] * 1.05, rtol=1e-14)       # below the table: first knot
    np.testing.assert_allclose(got[0, 0, 3], g[0, 0, 3] * 1.05, rtol=1e-14)       # on a knot
    np.testing.assert_allclose(got[0, 5, 0], g[0, 5, 0] * 1.025, rtol=1e-14)
    np.testing.assert_allclose(got[16, 16, 16], g[16, 16, 16] * 1.1, rtol=1e-14)  # above the table: last knot


def _invwin(gpu, n):
    from kspace_neutrinos_b200 import capi
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert gpu.ksn_bin_tables(n, n // 2, C.byref(thr), C.byref(iw)) == 0
    return iw


@pytest.mark.parametrize("n,dtype,tol", [(4, np.float64, 1e-10), (32, np.float64, 1e-10), (96, np.float64, 1e-10), (256, np.float64, 1e-10), (64, np.float32, 1e-5)])
def test_fused_greens_function_matches_the_two_separate_passes(gpu, n, dtype, tol):
    """K3 fused with the PM Green's function of Gadget-2's pmforce_periodic (the loop right after the
    add_nu_power_to_rhogrid hook, gadget-2/0002 patch:116-125).  One pass must equal the neutrino correction followed by
    the numpy restatement of that loop -- device-resident whole grid (TMA kernel for 256), host-resident ragged slab,
    the mean F(0,0,0) zeroed like there."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    asmth2 = (2 * np.pi * 1.25 / n) ** 2            # Asmth = 1.25 mesh cells, in grid units as in pm_periodic.c
    g = refs.random_grid(n, seed=3 * n + 1, dtype=dtype)
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    iw = _invwin(gpu, n)
    want = refs.greens_numpy(refs.k3_numpy(g, 0, box, logkk, ratio, norm), 0, asmth2)
    d = refs.DeviceBuffer(gpu, g)
    capi.check(gpu.ksn_scale_modes_greens(d.ptr, g.dtype.itemsize, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
    got = d.download(g)
    # the switch must not leak into the plain entry
    d.upload(g)
    capi.check(gpu.ksn_scale_modes(d.ptr, g.dtype.itemsize, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
    plain = d.download(g)
    d.free()
    assert got[0, 0, 0, 0] == 0 and got[0, 0, 0, 1] == 0
    np.testing.assert_allclose(got, want, rtol=tol, atol=0)
    np.testing.assert_allclose(plain, refs.k3_numpy(g, 0, box, logkk, ratio, norm), rtol=tol, atol=0)
    a, b = n // 4, n // 4 + max(1, n // 3)
    sub = np.ascontiguousarray(g[a:b])
    want_sub = refs.greens_numpy(refs.k3_numpy(sub, a, box, logkk, ratio, norm), a, asmth2)
    capi.check(gpu.ksn_scale_modes_greens(sub.ctypes.data_as(C.c_void_p), g.dtype.itemsize, n, a, b - a, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
    np.testing.assert_allclose(sub, want_sub, rtol=tol, atol=0)


def test_fused_step_entry_equals_step_then_greens(gpu):
    """add_nu_power_and_greens_to_rhogrid (extension) == add_nu_power_to_rhogrid followed by the Green's-function loop,
    with the same integrator state afterwards."""
    from kspace_neutrinos_b200 import capi
    n = 32
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    g = refs.random_grid(n, seed=5)
    outs = []
    for fused in (False, True):
        refs.init_module(gpu, n, masses=(0.15, 0.15, 0.15))
        dt = capi.global_delta_tot_table()
        d = refs.DeviceBuffer(gpu, g)
        for a in (0.01, 0.02):
            d.upload(g)
            if fused:
                gpu.add_nu_power_and_greens_to_rhogrid_f64(a, refs.BOX, d.ptr, n, 0, n, asmth2, 0)
            else:
                gpu.add_nu_power_to_rhogrid_f64(a, refs.BOX, d.ptr, n, 0, n, 0)
        cur = d.download(g)
        d.free()
        outs.append((cur, dt.ia, np.array([dt.delta_nu_last[i] for i in range(dt.nk)])))
    assert outs[0][1] == outs[1][1]
    # (not bit-equal: the very first sweep of a geometry uses the kernel that also bins keff/count, whose power sums
    # differ from the tile kernel's in the last bits)
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-12, atol=0)
    np.testing.assert_allclose(outs[1][0], refs.greens_numpy(outs[0][0], 0, asmth2), rtol=1e-11, atol=0)


@pytest.mark.parametrize("start,greens", [(0, False), (0, True), (3000, False)])
def test_long_rows_split_between_ctas_at_pmgrid_4096(gpu, start, greens):
    """A row of PMGRID = 4096 (2049 modes, 32.8 KB) does not fit the one-row-per-CTA bulk-copy kernel: it is cut into two
    pieces of 1025 and 1024 modes, one CTA each.  Thin slab of the full-size grid against the numpy restatement, and bit
    for bit against the plain-load kernel (KSN_K3_NOSPLIT), which computes the same factors."""
    import os
    from kspace_neutrinos_b200 import capi
    n, nslab, box = 4096, 2, refs.BOX
    rng = np.random.default_rng(11 + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2))
    kmin, kmax = 2 * np.pi / box, np.sqrt(3) * (n / 2) * 2 * np.pi / box
    nk = 300                                                         # jittered, but no two knots closer than 0.6 of the mean spacing
    step = (np.log(kmax * 0.8) - np.log(kmin * 1.3)) / (nk - 1)
    logkk = np.log(kmin * 1.3) + step * (np.arange(nk) + rng.uniform(-0.2, 0.2, nk))
    ratio, norm = 0.2 + 0.6 * rng.random(nk), 0.0123
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    iw = _invwin(gpu, n)

    def run():
        d = refs.DeviceBuffer(gpu, g)
        if greens:
            capi.check(gpu.ksn_scale_modes_greens(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
        else:
            capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
        out = d.download(g)
        d.free()
        return out

    split = run()
    os.environ["KSN_K3_NOSPLIT"] = "1"
    try:
        plain = run()
    finally:
        del os.environ["KSN_K3_NOSPLIT"]
    assert np.array_equal(split, plain)
    want = refs.k3_numpy(g, start, box, logkk, ratio, norm)
    if greens:
        want = refs.greens_numpy(want, start, asmth2)
    np.testing.assert_allclose(split, want, rtol=1e-10, atol=0)


@pytest.mark.parametrize("n,start,nslab,dtype,tol", [(64, 0, 64, np.float64, 1e-10), (64, 7, 30, np.float32, 1e-5), (33, 0, 33, np.float64, 1e-10),
                                                     (1024, 510, 4, np.float64, 1e-10), (2048, 0, 2, np.float64, 1e-10),
                                                     (2048, 1023, 3, np.float64, 1e-10), (4096, 2047, 2, np.float64, 1e-10)])
def test_fused_greens_function_against_the_gadget2_loop(gpu, n, start, nslab, dtype, tol):
    """The fused pass against the C restatement of GADGET-2.0.7's own loop (oracle/gadget2_greens.c: pm_periodic.c,
    pmforce_periodic, 'multiply with Green's function for the potential') applied after the oracle's restatement of the
    reference's scaling loop (interface_gadget.c:163-188) -- at small grids and on planes of the full-width grids, so that
    every K3 instantiation (rows sharing a CTA, one row per CTA, row pieces) meets it with the Green's function on."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    rng = np.random.default_rng(5 * n + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(dtype)
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    want = np.ascontiguousarray(g).copy()
    refs.orc().orc_scale_modes(want.ctypes.data_as(C.c_void_p), 1 if dtype == np.float64 else 0, n, start, nslab, box,
                               refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm)
    want = refs.greens_gadget2(want, start, asmth2)
    d = refs.DeviceBuffer(gpu, g)
    capi.check(gpu.ksn_scale_modes_greens(d.ptr, g.dtype.itemsize, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, _invwin(gpu, n), asmth2))
    got = d.download(g)
    d.free()
    assert b"Green's function fused" in gpu.ksn_last_k3_kernel()
    if n >= 1152 and n <= 2302:
        assert b"k3_scale_row_kernel<whole>" in gpu.ksn_last_k3_kernel()
    if start == 0:
        assert got[0, 0, 0, 0] == 0 and got[0, 0, 0, 1] == 0
    np.testing.assert_allclose(got, want, rtol=tol, atol=0)
