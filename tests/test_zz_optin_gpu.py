"""Opt-in kernels written after round 1's GPU minutes were spent (float grids -- K3: KSN_K3_F32_TMA=1, flat bulk-copy
chunks; K1: KSN_K1_F32_TILE=1, the tile kernel on float rows; the K1 bin window with a bin's home chosen per tile,
KSN_K1_WIN=3; K2 with one k bin per thread-block cluster, KSN_K2_CLUSTER=2|3|4; K3's flat-chunk kernel on double grids,
KSN_K3_FLAT=1; the collective bootstrap of the -DKSN_HAVE_MPI host layer on two GPUs).  They have NOT run on a B200 yet, so these tests are skipped unless
KSN_TEST_UNVERIFIED=1 (tools/gpu_round2_check.sh sets it); the default float path stays the verified one until then.
Each test compares the opt-in kernel with the numpy restatement / the reference AND, bit for bit where the arithmetic is
the same, with the default float kernel.  Also here, for the same reason: odd PMGRID on the device (default kernels, a
case the GPU suite did not cover in round 1)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import refs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("KSN_TEST_UNVERIFIED") != "1", reason="opt-in kernels: set KSN_TEST_UNVERIFIED=1")]


def _table(n, box, nk=40, seed=0):
    rng = np.random.default_rng(seed)
    kmin, kmax = 2 * np.pi / box, 2 * np.pi / box * np.sqrt(3) * n / 2
    logkk = np.sort(np.log(kmin) + (np.log(kmax) - np.log(kmin)) * rng.random(nk))
    return logkk, 0.01 + 0.1 * rng.random(nk), 0.8


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("n,start,nslab", [(4, 0, 4), (6, 0, 6), (64, 0, 64), (64, 5, 17), (126, 120, 6), (256, 0, 256), (2048, 1000, 6), (4096, 3000, 2)])
def test_k3_float_bulk_copy_kernel_equals_the_plain_float_kernel(gpu, n, start, nslab):
    """Same factor arithmetic, same narrowing to float: the two kernels must agree bit for bit; and with numpy to the
    float-grid tolerance."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(n + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(np.float32)
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    outs = []
    for knob in (None, "1"):
        with _env(KSN_K3_F32_TMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes(d.ptr, 4, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs.append(d.download(g))
            d.free()
    np.testing.assert_array_equal(outs[1], outs[0])
    if n <= 256:
        np.testing.assert_allclose(outs[1], refs.k3_numpy(g, start, box, logkk, ratio, norm), rtol=1e-5, atol=0)


def test_k3_float_bulk_copy_kernel_with_the_greens_function(gpu):
    from kspace_neutrinos_b200 import capi
    n, box = 64, refs.BOX
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    g = refs.random_grid(n, seed=77, dtype=np.float32)
    logkk, ratio, norm = _table(n, box)
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert gpu.ksn_bin_tables(n, n // 2, C.byref(thr), C.byref(iw)) == 0
    outs = []
    for knob in (None, "1"):
        with _env(KSN_K3_F32_TMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes_greens(d.ptr, 4, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
            outs.append(d.download(g))
            d.free()
    np.testing.assert_array_equal(outs[1], outs[0])
    assert outs[1][0, 0, 0, 0] == 0 and outs[1][0, 0, 0, 1] == 0


@pytest.mark.parametrize("n,nrbins", [(8, 4), (64, 32), (96, 48), (128, 64), (256, 128)])
def test_k1_float_tile_kernel_matches_the_pair_kernel_and_the_reference(gpu, n, nrbins):
    """K1 on float rows through the tile kernel: counts bit-exact, power within the north star's float-grid tolerance of
    the reference's float build, and within 3e-6 of the default float kernel (the tile kernel does not reproduce the
    reference's float roundings of the window product, the pair kernel does)."""
    g = refs.random_grid(n, seed=n + 1, dtype=np.float32)
    res = {}
    for knob in (None, "1"):
        with _env(KSN_K1_F32_TILE=knob):
            d = refs.DeviceBuffer(gpu, g)
            refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f32", pointer=d.ptr)      # first sweep: geometry
            res[knob] = refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f32", pointer=d.ptr)
            name = gpu.ksn_last_k1_kernel()
            d.free()
            if knob:
                assert b"k1_tile_kernel" in name and b"float" in name, name
    (n0, p0, c0, k0), (n1, p1, c1, k1) = res[None], res["1"]
    assert n0 == n1 and np.array_equal(c0[:n0], c1[:n1])
    np.testing.assert_array_equal(k1[:n1], k0[:n0])
    np.testing.assert_allclose(p1[:n1], p0[:n0], rtol=3e-6)
    ref = refs.ref_lib(False)
    if ref is not None:
        r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
        assert r_n == n1 and np.array_equal(r_c[:r_n], c1[:n1])
        np.testing.assert_allclose(p1[:n1], r_p[:r_n], rtol=1e-5)


def test_k1_float_tile_kernel_on_odd_slab_offsets_and_thin_slabs(gpu):
    """Odd rows of a float grid start 8 bytes off the bulk-copy granule (the kernel copies from one mode earlier): slabs
    that start on any plane, one plane thin, at PMGRID 2048 and 4096 (bin window)."""
    from tests.test_k1_gpu import _sums
    for n, start, nslab in ((64, 3, 1), (64, 0, 64), (2048, 1029, 2), (4096, 3, 1)):
        rng = np.random.default_rng(n)
        sub = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(np.float32)
        nrbins = n // 2
        out = {}
        for knob in (None, "1"):
            with _env(KSN_K1_F32_TILE=knob):
                class G:                                    # _sums slices g[start:start+nslab] and reads g.shape[1], g.dtype
                    shape = (start + nslab, n, n // 2 + 1, 2)
                    dtype = sub.dtype

                    def __getitem__(self, sl):
                        return sub
                _sums(gpu, G(), nrbins, start, nslab)       # first call per geometry
                out[knob] = _sums(gpu, G(), nrbins, start, nslab)
        assert np.array_equal(out[None][2], out["1"][2])
        np.testing.assert_allclose(out["1"][0], out[None][0], rtol=3e-6)


@pytest.mark.parametrize("n,nrbins", [(5, 4), (9, 8), (15, 7), (33, 16)])
def test_odd_pmgrid_double_matches_the_reference(gpu, n, nrbins):
    """Odd PMGRID (legal, if unusual: the reference then counts the z = dims/2 column once, powerspectrum.c:70-78).  The
    CPU side is pinned in tests/test_host_cpu.py; this is the device side, not yet run on a B200 -- hence in this file."""
    ref = refs.ref_lib(True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    g = refs.random_grid(n, seed=n)
    r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
    d = refs.DeviceBuffer(gpu, g)
    for sweep in ("first", "cached"):
        m_n, m_p, m_c, m_k = refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f64", pointer=d.ptr)
        assert m_n == r_n and np.array_equal(m_c[:m_n], r_c[:r_n]), sweep
        np.testing.assert_allclose(m_p[:m_n], r_p[:r_n], rtol=1e-10, atol=0, err_msg=sweep)
        np.testing.assert_allclose(m_k[:m_n], r_k[:r_n], rtol=1e-10, atol=0, err_msg=sweep)
    from kspace_neutrinos_b200 import capi
    logkk, ratio, norm = _table(n, refs.BOX, nk=max(3, n // 2))
    capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, 0, n, refs.BOX, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
    got = d.download(g)
    d.free()
    np.testing.assert_allclose(got, refs.k3_numpy(g, 0, refs.BOX, logkk, ratio, norm), rtol=1e-10, atol=0)


@pytest.mark.parametrize("n,start,nslab", [(256, 0, 256), (256, 100, 7), (512, 0, 40), (4096, 0, 6), (4096, 2040, 12)])
def test_k1_bin_window_with_the_home_chosen_per_tile_is_bit_identical(gpu, n, start, nslab):
    """KSN_K1_WIN=3 / 4 (k1_tile_kernel<.., 2>): same bins, same update order as the window kernel that chooses per update
    (KSN_K1_WIN=1 / 2, verified in round 1) -- the sums must be bit-identical.  Small grids force a quarter-size window
    (values 2 / 4); slabs at 4096 include the one with the k_x = k_y = 0 axis, whose rows reach the cold bins."""
    from kspace_neutrinos_b200 import capi
    from tests.test_k1_gpu import _sums
    nrbins = n // 2
    ptr = C.c_void_p()
    capi.check(gpu.ksn_device_malloc(C.byref(ptr), nslab * n * (n // 2 + 1) * 16))
    capi.check(gpu.ksn_fill_synthetic_grid(ptr, 8, n, start, nslab, 13, -1.0))
    shape = np.empty((nslab, n, 1, 1))
    _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)                 # geometry
    per_update, per_tile = ("2", "4") if n < 4096 else ("1", "3")
    with _env(KSN_K1_WIN=per_update):
        a = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        assert b"in shared memory" in gpu.ksn_last_k1_kernel() and b"per tile" not in gpu.ksn_last_k1_kernel()
    with _env(KSN_K1_WIN=per_tile):
        b = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
        assert b"home chosen per tile" in gpu.ksn_last_k1_kernel(), gpu.ksn_last_k1_kernel()
        b2 = _sums(gpu, shape, nrbins, start, nslab, pointer=ptr)
    gpu.ksn_device_free(ptr)
    np.testing.assert_array_equal(b[0], a[0])
    np.testing.assert_array_equal(b2[0], b[0])
    assert np.array_equal(a[2], b[2]) and a[3] == b[3]


@pytest.mark.parametrize("hybrid", [True, False])
@pytest.mark.parametrize("masses", [(0.1, 0.1, 0.1), (0.2, 0.1, 0.3)])
def test_k2_one_bin_per_cluster_is_bit_identical(gpu, masses, hybrid):
    """k2_delta_nu_cluster_kernel<M> (KSN_K2_CLUSTER=2|3|4): the M groups of a speculative pass are the M CTAs of a
    thread-block cluster, the interval list sits in CTA 0's shared memory.  Same arithmetic and the same replay as the
    one-CTA kernels, so delta_nu, the rule count and the table row count must equal the sequential kernel's bit for bit."""
    from tests.test_k2_gpu import _benchmark_state
    om = refs.make_omnu(gpu, masses)
    if hybrid:
        gpu.init_hybrid_nu(C.byref(om.hybnu), (C.c_double * 3)(*masses), 500.0, 2.99792458e10 / 1e5, 0.333, om.kBtnu)
    refs.set_background(gpu, om)
    runs = {}
    for cl in (0, 2, 3, 4):
        with _env(KSN_K2_SPEC="1", KSN_K2_CLUSTER=str(cl) if cl else None):
            d, kk, dcdm, tr = _benchmark_state(gpu, om)
            outs = []
            for a in (0.981, 0.982, 0.9915):
                g = np.zeros(len(kk))
                gpu.get_delta_nu_update(C.byref(d), a, len(kk), refs.dptr(kk), refs.dptr(dcdm), refs.dptr(g), C.byref(tr))
                outs.append((g.copy(), gpu.ksn_last_k2_max_passes(), gpu.ksn_last_k2_max_trips(), d.ia))
            runs[cl] = outs
    if hybrid:
        assert max(p for _, p, _, _ in runs[0]) > 60
    for cl in (2, 3, 4):
        for (g1, p1, t1, ia1), (gm, pm, tm, iam) in zip(runs[0], runs[cl]):
            assert ia1 == iam and p1 == pm, (cl, p1, pm)
            assert np.array_equal(g1, gm), (cl, float(np.max(np.abs(gm / g1 - 1))))
            assert tm <= t1                       # fewer passes through the integrand than sequential bisections


@pytest.mark.parametrize("n,start,nslab,greens", [(4, 0, 4, False), (64, 0, 64, False), (64, 7, 9, True), (96, 0, 96, False), (256, 0, 256, True),
                                                   (1024, 500, 8, False), (1024, 0, 3, True)])
def test_k3_flat_chunk_kernel_on_double_grids_with_short_rows_is_bit_identical(gpu, n, start, nslab, greens):
    """KSN_K3_FLAT=1: where several rows share a CTA (PMGRID <= 1150) the flat-chunk kernel finds row and z of a mode
    without a division per mode.  Same factor arithmetic as k3_scale_tma_kernel<double, false, false>: bit-identical."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(3 * n + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2))
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert gpu.ksn_bin_tables(n, n // 2, C.byref(thr), C.byref(iw)) == 0
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    outs = []
    for knob in (None, "1"):
        with _env(KSN_K3_FLAT=knob):
            d = refs.DeviceBuffer(gpu, g)
            if greens:
                capi.check(gpu.ksn_scale_modes_greens(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
            else:
                capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs.append(d.download(g))
            d.free()
    np.testing.assert_array_equal(outs[1], outs[0])
    if not greens and n <= 256:
        np.testing.assert_allclose(outs[1], refs.k3_numpy(g, start, box, logkk, ratio, norm), rtol=1e-10, atol=0)


@pytest.mark.parametrize("comm", ["p2p", "nccl", "mpi"])
def test_mpi_build_picks_its_collective_on_two_gpus(gpu, comm, tmp_path):
    """The -DKSN_HAVE_MPI host layer chooses the collective for the bin sums on the communicator it is handed
    (src/iface_common.c: bind_comm): two ranks of the fork-based mini-MPI (test infrastructure), one GPU each, linked to
    the REAL device library, take PM steps on x-slabs with each starting point of the list (KSN_COMM) and must reproduce
    the one-rank run."""
    import glob
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    PKG = os.path.join(ROOT, "kspace_neutrinos_b200")
    exe = str(tmp_path / "mpi_host_step_gpu")
    srcs = sorted(glob.glob(os.path.join(PKG, "src", "*.c")))
    cmd = ["gcc", "-O2", "-g", "-Wall", "-DKSN_HAVE_MPI", "-DDOUBLEPRECISION_FFTW",
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "src"), os.path.join(ROOT, "tests", "mpi_host_step.c"), *srcs, os.path.join(ROOT, "oracle", "mini_mpi.c"),
           "-L", PKG, "-lkspace_neutrinos_b200", f"-Wl,-rpath,{PKG}", "-lm", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = {k: v for k, v in os.environ.items() if k not in ("LOCAL_RANK", "KSN_DEVICE", "RANK", "WORLD_SIZE")}
    res = {}
    for ranks in (1, 2):
        out = str(tmp_path / f"out{ranks}.bin")
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat"), str(ranks), out], capture_output=True, text=True,
                           timeout=300, env={**env, "KSN_COMM": comm})
        assert r.returncode == 0 and "MPI HOST STEP OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
        if ranks == 2:
            assert "BACKEND 102" in r.stdout, r.stdout[-500:]
        raw = open(out, "rb").read()
        n, nk, ia = np.frombuffer(raw[:12], dtype=np.int32)
        res[ranks] = (int(nk), int(ia), np.frombuffer(raw[12:12 + 8 * nk], dtype=np.float64), np.frombuffer(raw[12 + 8 * nk:], dtype=np.float64))
    assert res[1][:2] == res[2][:2]
    np.testing.assert_allclose(res[2][2], res[1][2], rtol=1e-10)
    np.testing.assert_allclose(res[2][3], res[1][3], rtol=1e-10)


@pytest.mark.parametrize("n,start,nslab", [(64, 0, 64), (256, 0, 256), (2048, 1000, 4)])
def test_k3_float_short_series_factor_stays_within_one_float_rounding(gpu, n, start, nslab):
    """KSN_K3_F32_TMA=2: ln(1+u) stops at u^4 for the narrow bins (truncation <= 7e-9 of a small correction term) -- the
    factor then differs from the full one by far less than the float rounding of the product, so results agree with the
    default float kernel to one float ulp, and mostly bit for bit."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(n + 17)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(np.float32)
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    outs = []
    for knob in (None, "2"):
        with _env(KSN_K3_F32_TMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes(d.ptr, 4, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs.append(d.download(g))
            d.free()
    np.testing.assert_allclose(outs[1], outs[0], rtol=1.3e-7, atol=0)
    assert np.mean(outs[1] == outs[0]) > 0.9
