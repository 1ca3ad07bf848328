"""The full-width parity harness (oracle/_ref/ref_slabs, used by tests/test_fullsize_parity_gpu.py) checked on the CPU at a
size where the whole grid fits: two ranks that tile a 32^3 grid must reproduce the in-process reference library
(oracle/_ref/libksref_double.so) -- spectrum, integrator state after every step, corrected grid -- and plane ranges that do
NOT tile the grid must give counts that add up from the ranges' own spectra."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import refs
from tests.test_fullsize_parity_gpu import REF_SLABS, TRANSFER, read_out


def _run(tmp_path, g, n, slabs, masses, hybrid, times, tag):
    inp, out = str(tmp_path / f"in{tag}.bin"), str(tmp_path / f"ref{tag}.bin")
    with open(inp, "wb") as f:
        for s, c in slabs:
            f.write(np.ascontiguousarray(g[s:s + c]).tobytes())
    args = [REF_SLABS, str(n), str(int(hybrid))] + [repr(m) for m in masses] + [TRANSFER, inp, out, str(len(slabs))]
    args += [str(x) for sl in slabs for x in sl] + [str(len(times))] + [repr(t) for t in times]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    return read_out(out), np.fromfile(out + ".grid")


def test_ref_slabs_reproduces_the_in_process_reference(tmp_path):
    if not os.path.exists(REF_SLABS):
        pytest.skip("oracle/_ref/ref_slabs not built")
    ref = refs.ref_lib(True)
    n, masses, times = 32, (0.2, 0.1, 0.3), (0.01, 0.02, 0.05)
    g = refs.random_grid(n, seed=9)
    got, grid = _run(tmp_path, g, n, [(0, 13), (13, 19)], masses, False, times, "a")
    nret, P, Cn, K = refs.total_powerspectrum(ref, g, n // 2)
    assert got["nret"] == nret and np.array_equal(got["C"], Cn[:nret].astype(np.float64))
    np.testing.assert_allclose(got["P"], P[:nret], rtol=1e-13)
    np.testing.assert_allclose(got["K"], K[:nret], rtol=1e-13)
    om, dt = refs.init_module(ref, n, masses=masses)
    want = g.copy()
    for i, a in enumerate(times):
        ref.add_nu_power_to_rhogrid(a, refs.BOX, want.ctypes.data_as(C.c_void_p), n, 0, n, 0)
        np.testing.assert_allclose(got["dnu"][i], np.array([dt.delta_nu_last[k] for k in range(dt.nk)]), rtol=1e-12)
    assert (got["nk"], got["ia"]) == (dt.nk, dt.ia)
    np.testing.assert_allclose(grid, want.reshape(-1), rtol=1e-12)


def test_ref_slabs_on_plane_ranges_that_do_not_tile_the_grid(tmp_path):
    if not os.path.exists(REF_SLABS):
        pytest.skip("oracle/_ref/ref_slabs not built")
    n = 32
    g = refs.random_grid(n, seed=10)
    both, _ = _run(tmp_path, g, n, [(0, 3), (15, 4)], (0.1, 0.1, 0.1), True, (0.01,), "b")
    first, _ = _run(tmp_path, g, n, [(0, 3)], (0.1, 0.1, 0.1), True, (0.01,), "c")
    # every plane holds n rows of multiplicity 1 + 2 (n/2 - 1) + 1; F(0,0,0) is not a mode
    kz_mult = np.r_[1, 2 * np.ones(n // 2 - 1), 1]
    assert both["C"].sum() == 7 * n * kz_mult.sum() - 1
    assert first["C"].sum() == 3 * n * kz_mult.sum() - 1
