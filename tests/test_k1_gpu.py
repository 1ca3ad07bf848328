"""K1 (total_powerspectrum) on the GPU against the reference's known answers, the reference
sources compiled here (oracle/_ref) and size-independent properties.  Calls go through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from tests import refs

pytestmark = pytest.mark.gpu

# powerspectrum_test.c:38-45
KAT_COUNTS = [6, 12, 8, 3, 12, 12, 3, 6, 1]


def test_kat_4cube_host_pointer(gpu):
    g = refs.kat_grid_4()
    nret, power, count, keffs = refs.total_powerspectrum(gpu, g, 15, fn="total_powerspectrum_f64")
    assert nret == 9
    assert list(count[:9]) == KAT_COUNTS
    assert abs(keffs[2] - 1.73205) < 1e-5
    assert abs(power[0] - 0.254834) < 1e-5 * 0.04
    assert abs(power[1] - 0.00212722) < 1e-5 * 0.005
    assert abs(power[2] - 0.00323766) < 1e-5 * 0.003


def test_kat_4cube_device_pointer_and_float(gpu):
    g = refs.kat_grid_4()
    d = refs.DeviceBuffer(gpu, g)
    nret, power, count, keffs = refs.total_powerspectrum(gpu, g, 15, fn="total_powerspectrum", pointer=d.ptr)
    d.free()
    assert nret == 9 and list(count[:9]) == KAT_COUNTS
    assert abs(power[0] - 0.254834) < 1e-5 * 0.04
    g32 = g.astype(np.float32)
    nret, power, count, keffs = refs.total_powerspectrum(gpu, g32, 15, fn="total_powerspectrum_f32")
    assert nret == 9 and list(count[:9]) == KAT_COUNTS
    assert abs(power[1] - 0.00212722) < 1e-5 * 0.005


@pytest.mark.parametrize("n,nrbins", [(8, 4), (16, 8), (32, 16), (64, 32), (96, 48), (128, 64), (128, 200), (192, 96)])
def test_matches_reference_double(gpu, n, nrbins):
    ref = refs.ref_lib(True)
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here); covered by the oracle restatement test")
    g = refs.random_grid(n, seed=n)
    r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
    d = refs.DeviceBuffer(gpu, g)
    # first sweep of a geometry: k1_bin_kernel (power, keff, counts); second: the tile kernel (power) + cached geometry
    for sweep in ("first", "cached"):
        m_n, m_p, m_c, m_k = refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f64", pointer=d.ptr)
        assert m_n == r_n, sweep
        assert np.array_equal(m_c[:m_n], r_c[:r_n]), sweep             # mode counts: bit-exact
        np.testing.assert_allclose(m_p[:m_n], r_p[:r_n], rtol=1e-10, atol=0, err_msg=sweep)   # north-star tolerance, double grid
        np.testing.assert_allclose(m_k[:m_n], r_k[:r_n], rtol=1e-10, atol=0, err_msg=sweep)
    assert b"k1_tile_kernel" in gpu.ksn_last_k1_kernel()
    d.free()


def test_matches_reference_float(gpu):
    ref = refs.ref_lib(False)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    n, nrbins = 64, 32
    g = refs.random_grid(n, seed=5, dtype=np.float32)
    r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
    m_n, m_p, m_c, m_k = refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f32")
    assert m_n == r_n and np.array_equal(m_c[:m_n], r_c[:r_n])
    np.testing.assert_allclose(m_p[:m_n], r_p[:r_n], rtol=1e-5)    # float-grid tolerance of the north star
    np.testing.assert_allclose(m_k[:m_n], r_k[:r_n], rtol=1e-10)


def _sums(gpu, g, nrbins, startslab, nslab, pointer=None):
    n = g.shape[1]
    from kspace_neutrinos_b200 import capi
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert gpu.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw)) == 0
    power, keff = np.zeros(nrbins), np.zeros(nrbins)
    count = np.zeros(nrbins, dtype=np.int64)
    m2 = C.c_double()
    sub = np.ascontiguousarray(g[startslab:startslab + nslab])
    p = pointer if pointer is not None else sub.ctypes.data_as(C.c_void_p)
    capi.check(gpu.ksn_powerspectrum_sums(p, g.dtype.itemsize, n, nrbins, startslab, nslab, thr, iw, refs.dptr(power), refs.dptr(keff),
                                          count.ctypes.data_as(capi.c_longlong_p), C.byref(m2)))
    return power, keff, count, m2.value


@pytest.mark.parametrize("splits", [[0, 13, 64], [0, 1, 2, 40, 64], [0, 0, 64, 64]])
def test_slab_partition_sums_to_whole(gpu, splits):
    """Ragged and empty slabs: per-slab bin sums add up to the whole-grid sums (exactly for counts)."""
    n, nrbins = 64, 32
    g = refs.random_grid(n, seed=11)
    P, K, Cn, M2 = _sums(gpu, g, nrbins, 0, n)
    p = np.zeros(nrbins); k = np.zeros(nrbins); c = np.zeros(nrbins, dtype=np.int64); m2 = 0.0
    for a, b in zip(splits[:-1], splits[1:]):
        pp, kk, cc, mm = _sums(gpu, g, nrbins, a, b - a)
        p += pp; k += kk; c += cc; m2 += mm
    assert np.array_equal(c, Cn) and Cn.sum() == n ** 3 - 1
    assert m2 == M2 == float(n) ** 6
    np.testing.assert_allclose(p, P, rtol=1e-12)
    np.testing.assert_allclose(k, K, rtol=1e-12)


def test_bitwise_reproducible(gpu):
    n, nrbins = 128, 64
    g = refs.random_grid(n, seed=3)
    d = refs.DeviceBuffer(gpu, g)
    first = _sums(gpu, g, nrbins, 0, n, pointer=d.ptr)   # first call per geometry: the variant that also bins keff/count
    a = _sums(gpu, g, nrbins, 0, n, pointer=d.ptr)
    b = _sums(gpu, g, nrbins, 0, n, pointer=d.ptr)
    h = _sums(gpu, g, nrbins, 0, n)        # staged from host memory, chunked
    d.free()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    np.testing.assert_allclose(first[0], a[0], rtol=1e-13)
    np.testing.assert_allclose(h[0], a[0], rtol=1e-13)


def test_linearity_large(gpu):
    """At a size the CPU oracle would take minutes for: P(c*grid) == P(grid) (normalised by |F(0)|^2),
    sum(count) == N^3-1, and the last compacted bin holds exactly the corner mode."""
    n, nrbins = 512, 256
    from kspace_neutrinos_b200 import capi
    nbytes = n * n * (n // 2 + 1) * 16
    ptr = C.c_void_p()
    capi.check(gpu.ksn_device_malloc(C.byref(ptr), nbytes))
    capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, 0, n, 20261017, -1.0))
    shape = np.empty((n, n, 1, 1))         # only shape[0], shape[1] are read
    n1, p1, c1, k1 = refs.total_powerspectrum(gpu, shape, nrbins, fn="total_powerspectrum_f64", pointer=ptr)
    assert c1[:n1].sum() == n ** 3 - 1 and c1[n1 - 1] == 1
    # scale the grid by a k-independent factor with K3 (ratio table == const) and re-measure
    logkk = np.array([-20.0, 20.0]); ratio = np.array([1.0, 1.0])
    capi.check(gpu.ksn_scale_modes(ptr, 8, n, 0, n, 1.0, refs.dptr(logkk), refs.dptr(ratio), 2, 0.5))
    n2, p2, c2, k2 = refs.total_powerspectrum(gpu, shape, nrbins, fn="total_powerspectrum_f64", pointer=ptr)
    gpu.ksn_device_free(ptr)
    assert n2 == n1 and np.array_equal(c1, c2)
    # every mode except (0,0,0) was multiplied by 1.5 -> P grows by 2.25 relative to the unchanged |F(0)|^2
    np.testing.assert_allclose(p2[:n1], 2.25 * p1[:n1], rtol=1e-12)
    np.testing.assert_array_equal(k1, k2)


def test_odd_aligned_device_pointer(gpu):
    """A device slab that starts on an odd 16-byte boundary (the 256-bit pair loads must re-align per row)."""
    from kspace_neutrinos_b200 import capi
    n, nrbins = 64, 32
    g = refs.random_grid(n, seed=21)
    ptr = C.c_void_p()
    capi.check(gpu.ksn_device_malloc(C.byref(ptr), g.nbytes + 64))
    shifted = C.c_void_p(ptr.value + 16)
    capi.check(gpu.ksn_memcpy_h2d(shifted, g.ctypes.data_as(C.c_void_p), g.nbytes))
    a = _sums(gpu, g, nrbins, 0, n, pointer=shifted)
    a = _sums(gpu, g, nrbins, 0, n, pointer=shifted)         # second call: the pair kernel
    d = refs.DeviceBuffer(gpu, g)
    b = _sums(gpu, g, nrbins, 0, n, pointer=d.ptr)
    d.free()
    gpu.ksn_device_free(ptr)
    assert np.array_equal(a[2], b[2]) and a[3] == b[3]
    np.testing.assert_allclose(a[0], b[0], rtol=1e-13)


@pytest.mark.parametrize("n,start,nslab,tile", [(64, 0, 64, None), (128, 3, 77, None), (256, 0, 256, None), (256, 100, 31, "8,5,2"),
                                                (256, 0, 256, "4,9,3"), (512, 0, 64, None), (512, 200, 48, "6,17,2")])
def test_tile_pair_and_full_kernels_agree(gpu, n, start, nslab, tile, monkeypatch):
    """The three K1 kernels (k1_bin_kernel: all three sums, first call per geometry; k1_pair_kernel: 256-bit loads +
    segmented scans; k1_tile_kernel: TMA tiles + per-lane sequential walk, fast and general tiles) sum the same modes in
    different orders: same power sums to rounding, on whole grids, ragged slabs and forced tile shapes."""
    from kspace_neutrinos_b200 import capi
    nrbins = n // 2
    g = refs.random_grid(n, seed=n + start)[start:start + nslab].copy()
    d = refs.DeviceBuffer(gpu, g)
    monkeypatch.setenv("KSN_NO_GEOM_CACHE", "1")
    full = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    assert b"k1_bin_kernel" in gpu.ksn_last_k1_kernel()
    monkeypatch.delenv("KSN_NO_GEOM_CACHE")
    _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)              # fills the geometry cache
    monkeypatch.setenv("KSN_K1_PAIR", "1")
    pair = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    assert b"k1_pair_kernel" in gpu.ksn_last_k1_kernel()
    monkeypatch.delenv("KSN_K1_PAIR")
    if tile:
        monkeypatch.setenv("KSN_K1_TILE", tile)
    tl = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    name = gpu.ksn_last_k1_kernel()
    assert b"k1_tile_kernel" in name, name
    if tile:
        w, c, s = tile.split(",")
        assert f"({w} warps x {c} modes per lane, {s} stages".encode() in name, name
    tl2 = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    d.free()
    assert np.array_equal(tl[0], tl2[0])                             # run-to-run bitwise
    assert np.array_equal(full[2], tl[2]) and np.array_equal(full[2], pair[2])
    live = full[0] != 0
    assert np.array_equal(live, tl[0] != 0) and np.array_equal(live, pair[0] != 0)
    np.testing.assert_allclose(tl[0][live], full[0][live], rtol=2e-13)
    np.testing.assert_allclose(pair[0][live], full[0][live], rtol=2e-13)
    assert tl[3] == full[3] == pair[3]


def test_host_grid_staged_in_chunks_matches_device_grid(gpu):
    """A host-resident slab larger than one 256 MB staging chunk: K1 runs on each chunk as it lands and accumulates
    into the same per-CTA partials; the sums must equal the device-resident sweep of the same bytes."""
    from kspace_neutrinos_b200 import capi
    n, nrbins, nslab = 512, 256, 320                  # 320 planes x 2.1 MB = 673 MB -> 3 chunks
    nbytes = nslab * n * (n // 2 + 1) * 16
    ptr = C.c_void_p()
    capi.check(gpu.ksn_device_malloc(C.byref(ptr), nbytes))
    capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, 100, nslab, 7, -1.0))
    host = np.empty((nslab, n, n // 2 + 1, 2))
    capi.check(gpu.ksn_memcpy_d2h(host.ctypes.data_as(C.c_void_p), ptr, nbytes))
    _sums(gpu, host, nrbins, 100, nslab, pointer=ptr)                 # geometry
    dev = _sums(gpu, host, nrbins, 100, nslab, pointer=ptr)
    hst = _sums(gpu, host, nrbins, 100, nslab, pointer=host.ctypes.data_as(C.c_void_p))
    gpu.ksn_device_free(ptr)
    assert np.array_equal(dev[2], hst[2])
    live = dev[0] != 0
    np.testing.assert_allclose(hst[0][live], dev[0][live], rtol=1e-13)


def test_full_size_slabs_kernels_agree(gpu):
    """At the PMGRID of BASELINE configs 4 and 5 (2048, 4096), on slabs as one of 8 GPUs would hold them (a few planes
    here): the tile kernel against the scan-based kernel, and the multiplicity-weighted mode count of the slab."""
    from kspace_neutrinos_b200 import capi
    import os
    for n, start, nslab in ((2048, 1000, 24), (4096, 2040, 12), (4096, 0, 6)):
        nrbins = n // 2
        nbytes = nslab * n * (n // 2 + 1) * 16
        ptr = C.c_void_p()
        capi.check(gpu.ksn_device_malloc(C.byref(ptr), nbytes))
        capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, start, nslab, 11, -1.0))
        shape = np.empty((nslab, n, 1, 1))
        first = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        tile = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        assert b"k1_tile_kernel" in gpu.ksn_last_k1_kernel()
        os.environ["KSN_K1_PAIR"] = "1"
        try:
            pair = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        finally:
            del os.environ["KSN_K1_PAIR"]
        gpu.ksn_device_free(ptr)
        # every stored mode counted with its Hermitian multiplicity: N per (x, y) pair, minus the mean
        assert first[2].sum() == nslab * n * n - (1 if start == 0 else 0)
        live = first[0] != 0
        np.testing.assert_allclose(tile[0][live], first[0][live], rtol=2e-13)
        np.testing.assert_allclose(pair[0][live], first[0][live], rtol=2e-13)


@pytest.mark.parametrize("n,start,nslab", [(256, 0, 256), (256, 100, 31), (512, 0, 64), (512, 250, 24)])
def test_bin_window_kernel_agrees_small(gpu, n, start, nslab, monkeypatch):
    """k1_tile_kernel<CT, true> keeps only the upper bins of a warp's private copy in shared memory and the rarely used
    lower ones in a global-memory array of the warp's own.  KSN_K1_WIN=2 forces the window on small grids with only a
    QUARTER of the bins in shared memory, so the global-memory path is exercised hard; sums must equal the plain tile
    kernel's and the three-sum kernel's to rounding, run to run bit for bit."""
    nrbins = n // 2
    g = refs.random_grid(n, seed=2 * n + start)[start:start + nslab].copy()
    d = refs.DeviceBuffer(gpu, g)
    monkeypatch.setenv("KSN_NO_GEOM_CACHE", "1")
    full = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    monkeypatch.delenv("KSN_NO_GEOM_CACHE")
    _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)              # fills the geometry cache
    monkeypatch.setenv("KSN_K1_WIN", "0")
    plain = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    assert b"k1_tile_kernel" in gpu.ksn_last_k1_kernel() and b"in shared memory" not in gpu.ksn_last_k1_kernel()
    monkeypatch.setenv("KSN_K1_WIN", "2")
    win = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    name = gpu.ksn_last_k1_kernel()
    assert b"k1_tile_kernel" in name and f"of {nrbins} in shared memory".encode() in name, name
    win2 = _sums(gpu, g, nrbins, start, nslab, pointer=d.ptr)
    d.free()
    assert np.array_equal(win[0], win2[0])
    assert np.array_equal(full[2], win[2]) and win[3] == full[3]
    live = full[0] != 0
    assert np.array_equal(live, win[0] != 0)
    np.testing.assert_allclose(win[0][live], plain[0][live], rtol=2e-13)
    np.testing.assert_allclose(win[0][live], full[0][live], rtol=2e-13)


def test_bin_window_kernel_agrees_at_pmgrid_4096(gpu, monkeypatch):
    """At PMGRID = 4096 (2048 bins, 16 KB per warp) the window is what makes room for eight warps.  Slabs of the full-size
    grid -- one holding the k_x = k_y = 0 axis, whose rows reach the lowest bins -- against the six-warp kernel without
    the window and against the scan-based kernel."""
    from kspace_neutrinos_b200 import capi
    n, nrbins = 4096, 2048
    for start, nslab in ((0, 6), (2040, 12)):
        nbytes = nslab * n * (n // 2 + 1) * 16
        ptr = C.c_void_p()
        capi.check(gpu.ksn_device_malloc(C.byref(ptr), nbytes))
        capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, start, nslab, 13, -1.0))
        shape = np.empty((nslab, n, 1, 1))
        first = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        monkeypatch.setenv("KSN_K1_WIN", "0")
        plain = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        assert b"(6 warps x 17 modes per lane" in gpu.ksn_last_k1_kernel(), gpu.ksn_last_k1_kernel()
        monkeypatch.setenv("KSN_K1_WIN", "1")
        win = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        name = gpu.ksn_last_k1_kernel()
        assert b"(8 warps x 17 modes per lane" in name and b"bins >= 1024 of 2048 in shared memory" in name, name
        monkeypatch.delenv("KSN_K1_WIN")
        monkeypatch.setenv("KSN_K1_PAIR", "1")
        pair = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        monkeypatch.delenv("KSN_K1_PAIR")
        gpu.ksn_device_free(ptr)
        live = first[0] != 0
        assert np.array_equal(live, win[0] != 0)
        np.testing.assert_allclose(win[0][live], plain[0][live], rtol=2e-13)
        np.testing.assert_allclose(win[0][live], pair[0][live], rtol=2e-13)
        np.testing.assert_allclose(win[0][live], first[0][live], rtol=2e-13)


@pytest.mark.parametrize("n,start,nslab", [(32, 0, 32), (2048, 1023, 2)])
def test_synthetic_grid_equals_the_cpu_generator_of_the_reference_arm(gpu, n, start, nslab):
    """bench.py's GPU arm fills its slabs with ksn_fill_synthetic_grid; the reference's CPU arm (oracle/ref_bench.c) with
    the restatement in oracle/synthetic_grid.h -- the same counter-based field, equal up to the last bits of log/sincos."""
    from kspace_neutrinos_b200 import capi
    want = np.zeros((nslab, n, n // 2 + 1, 2))
    refs.orc().orc_fill_synthetic_grid(refs.dptr(want), n, start, nslab, 20261017, -1.0)
    d = refs.DeviceBuffer(gpu, want)
    capi.check(gpu.ksn_fill_synthetic_grid(d.ptr, 8, n, start, nslab, 20261017, -1.0))
    got = d.download(want)
    d.free()
    if start == 0:
        assert got[0, 0, 0, 0] == float(n) ** 3 and got[0, 0, 0, 1] == 0
    scale = np.hypot(want[..., 0], want[..., 1])[..., None] + 1e-300
    assert np.max(np.abs(got - want) / scale) < 1e-13
