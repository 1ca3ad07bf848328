"""GPU part of tests/test_reference_programs.py (named to sort last): the reference's own cmocka programs that reach
the kernels -- powerspectrum_test.c (K1 on the 4^3 known-answer grid, double and float builds) and
delta_tot_table_test.c (K2 behind get_delta_nu_update, save/resume, fslength, specialJ, the 99-step CAMB linear-theory
run) -- linked to the product's libkspace_neutrinos_b200.so (oracle/_ref/dropin_*_test; prebuilt, /root/reference is
not read at run time; their data files come from oracle/_ref/fixtures)."""
import pytest

from tests.test_reference_programs import dropin, fixtures_dir, n_run_failed, run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ntests", [("powerspectrum", 1), ("powerspectrum_single", 1), ("delta_tot_table", 7),
                                         ("omega_nu_single", 9), ("transfer_init", 1), ("delta_pow", 2)])
def test_reference_program_passes_against_the_product_on_the_gpu(gpu, name, ntests, tmp_path):
    rc, out = run(dropin(name), fixtures_dir(tmp_path))
    assert rc == 0, out[-3000:]
    assert n_run_failed(out) == (ntests, 0), out[-3000:]


@pytest.mark.parametrize("n", [1024, 2048])
def test_mode_counts_at_the_baseline_sizes_equal_the_golden_vector(gpu, n):
    """BASELINE.json configs[2], [3] at FULL size (8.6 / 68.8 GB grids resident in HBM): total_powerspectrum's
    data-independent outputs -- which bins are non-empty, the mode count of each (bit-exact) and keff (1e-10, the
    north-star tolerance) -- against tests/golden/geometry_counts.npz, i.e. against the reference's own bin expression and
    multiplicities (tools/make_golden_geometry.py; that generator equals the compiled reference at every size it can run)."""
    import ctypes as C
    import os

    import numpy as np

    from kspace_neutrinos_b200 import capi
    from tests import refs
    g = np.load(os.path.join(refs.GOLDEN, "geometry_counts.npz"))
    cnt, ksum = g[f"count_{n}"], g[f"keffsum_{n}"]
    nrbins = n // 2
    nbytes = n * n * (n // 2 + 1) * 16
    ptr = C.c_void_p()
    capi.check(gpu.ksn_device_malloc(C.byref(ptr), nbytes))
    try:
        capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, 0, n, 20261017, -1.0))
        shape = np.empty((n, n, 1, 1))         # only shape[0], shape[1] are read
        nret, power, count, keffs = refs.total_powerspectrum(gpu, shape, nrbins, fn="total_powerspectrum_f64", pointer=ptr)
        # a second sweep takes the cached-geometry (power-only) kernel: same counts and keff must come back
        nret2, power2, count2, keffs2 = refs.total_powerspectrum(gpu, shape, nrbins, fn="total_powerspectrum_f64", pointer=ptr)
    finally:
        gpu.ksn_device_free(ptr)
    keep = cnt > 0
    assert nret == nret2 == np.count_nonzero(keep)
    assert np.array_equal(count[:nret], cnt[keep]) and np.array_equal(count2[:nret], cnt[keep])
    np.testing.assert_allclose(keffs[:nret], ksum[keep] / cnt[keep], rtol=1e-10)
    np.testing.assert_array_equal(keffs2[:nret], keffs[:nret])
    assert np.all(power[:nret] > 0)
    np.testing.assert_allclose(power2[:nret], power[:nret], rtol=1e-10)      # first-call kernel vs tile kernel
