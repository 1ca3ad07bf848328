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
