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
