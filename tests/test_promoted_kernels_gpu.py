"""Kernels that were written at the end of round 1 and verified on a B200 at the start of round 2 (profiles/r2_optin_*):
float grids through the bulk-copy kernels (K1: the tile kernel on float rows; K3: flat chunks, with the factor in float
where the table allows it), K3's flat-chunk kernel on double grids with short rows, odd PMGRID on the device, the
collective bootstrap of the -DKSN_HAVE_MPI host layer on two GPUs.  Each bulk-copy kernel is compared with the numpy
restatement / the reference AND, bit for bit where the arithmetic is the same, with the plain-load kernel that the
library keeps as the fall-back (KSN_K1_PAIR=1, KSN_K3_NOTMA=1)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import refs

pytestmark = pytest.mark.gpu


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
    float-grid tolerance.  (This table is far too steep for the all-float factor: both kernels evaluate it in double.)"""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(n + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(np.float32)
    logkk, ratio, norm = _table(n, box, nk=min(40, max(3, n // 2)))
    outs = []
    for knob in ("1", None):
        with _env(KSN_K3_NOTMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes(d.ptr, 4, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs.append(d.download(g))
            d.free()
            assert (b"plain loads" if knob else b"k3_scale_flat_kernel<float>") in gpu.ksn_last_k3_kernel(), gpu.ksn_last_k3_kernel()
            assert b"factor in float" not in gpu.ksn_last_k3_kernel()
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
    for knob in ("1", None):
        with _env(KSN_K3_NOTMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes_greens(d.ptr, 4, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
            outs.append(d.download(g))
            d.free()
    np.testing.assert_array_equal(outs[1], outs[0])
    assert outs[1][0, 0, 0, 0] == 0 and outs[1][0, 0, 0, 1] == 0


@pytest.mark.parametrize("n,nrbins", [(8, 4), (64, 32), (96, 48), (128, 64), (256, 128)])
def test_k1_float_tile_kernel_matches_the_pair_kernel_and_the_reference(gpu, n, nrbins):
    """K1 on float rows through the tile kernel (default): counts bit-exact, power within the north star's float-grid
    tolerance of the reference's float build, and within 3e-6 of the scan-based float kernel (the tile kernel does not
    reproduce the reference's float roundings of the window product, the pair kernel does)."""
    g = refs.random_grid(n, seed=n + 1, dtype=np.float32)
    res = {}
    for knob in ("1", None):
        with _env(KSN_K1_PAIR=knob):
            d = refs.DeviceBuffer(gpu, g)
            refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f32", pointer=d.ptr)      # first sweep: geometry
            res[knob] = refs.total_powerspectrum(gpu, g, nrbins, fn="total_powerspectrum_f32", pointer=d.ptr)
            name = gpu.ksn_last_k1_kernel()
            d.free()
            assert (b"k1_pair_kernel<float>" in name) if knob else (b"k1_tile_kernel<float>" in name), name
    (n0, p0, c0, k0), (n1, p1, c1, k1) = res["1"], res[None]
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
        for knob in ("1", None):
            with _env(KSN_K1_PAIR=knob):
                class G:                                    # _sums slices g[start:start+nslab] and reads g.shape[1], g.dtype
                    shape = (start + nslab, n, n // 2 + 1, 2)
                    dtype = sub.dtype

                    def __getitem__(self, sl):
                        return sub
                _sums(gpu, G(), nrbins, start, nslab)       # first call per geometry
                out[knob] = _sums(gpu, G(), nrbins, start, nslab)
        assert np.array_equal(out[None][2], out["1"][2])
        np.testing.assert_allclose(out[None][0], out["1"][0], rtol=3e-6)


@pytest.mark.parametrize("n,nrbins", [(5, 4), (9, 8), (15, 7), (33, 16)])
def test_odd_pmgrid_double_matches_the_reference(gpu, n, nrbins):
    """Odd PMGRID (legal, if unusual: the reference then counts the z = dims/2 column once, powerspectrum.c:70-78).  The
    CPU side is pinned in tests/test_host_cpu.py; this is the device side."""
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


@pytest.mark.parametrize("n,start,nslab,greens", [(4, 0, 4, False), (64, 0, 64, False), (64, 7, 9, True), (96, 0, 96, False), (256, 0, 256, True),
                                                   (1024, 500, 8, False), (1024, 0, 3, True)])
def test_k3_flat_chunk_kernel_on_double_grids_with_short_rows_is_bit_identical(gpu, n, start, nslab, greens):
    """Where several rows share a CTA (PMGRID <= 1150) the flat-chunk kernel finds row and z of a mode without a division
    per mode.  Same factor arithmetic as the plain-load kernel: bit-identical."""
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
    for knob in ("1", None):
        with _env(KSN_K3_NOTMA=knob):
            d = refs.DeviceBuffer(gpu, g)
            if greens:
                capi.check(gpu.ksn_scale_modes_greens(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm, iw, asmth2))
            else:
                capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs.append(d.download(g))
            d.free()
            assert (b"plain loads" if knob else b"k3_scale_flat_kernel<double>") in gpu.ksn_last_k3_kernel(), gpu.ksn_last_k3_kernel()
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


def _smooth_table(n, box, nk=None, seed=1, jitter=0.002):
    """A table like the one a PM step produces: one knot per P(k) bin (logarithmic bins, as narrow as the grid's own:
    u_max ~ 1.5 % at PMGRID 2048), delta_nu/delta_cdm falling smoothly with k plus a per-bin jitter, small norm."""
    rng = np.random.default_rng(seed)
    nk = nk or n // 2
    kmin, kmax = 2 * np.pi / box, np.sqrt(3) * (n / 2) * 2 * np.pi / box
    logkk = np.linspace(np.log(kmin * 1.1), np.log(kmax * 0.97), nk)
    ratio = 0.9 / (1 + (np.exp(logkk) / np.exp(logkk[nk // 3])) ** 2) * (1 + jitter * rng.standard_normal(nk))
    return logkk, ratio, 0.0073


@pytest.mark.parametrize("n,start,nslab", [(2048, 0, 2), (2048, 1023, 2), (4096, 2047, 1), (1024, 100, 4), (256, 0, 256)])
def test_k3_short_series_for_smooth_tables_equals_the_full_series(gpu, n, start, nslab):
    """k3_upload_table lets the double passes stop ln(1+u) at u^5 where |B| u_max^6 / 6 <= 1e-14 on every narrow segment
    (four FP64 instructions less per mode; the bins of PMGRID >= 2048 are that narrow): against KSN_K3_EXACT=1 (always to
    u^9) the corrected modes may differ by 1e-14 of their value, and against the oracle's restatement of
    interface_gadget.c:163-188 they hold the 1e-10 bar."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(n + start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2))
    logkk, ratio, norm = _smooth_table(n, box)
    outs = {}
    for knob in ("1", None):
        with _env(KSN_K3_EXACT=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs[knob] = d.download(g)
            d.free()
            name = gpu.ksn_last_k3_kernel()
            if knob:
                assert b"series to u^9" in name, name
            elif n >= 2048:
                assert b"series to u^5" in name, name
    np.testing.assert_allclose(outs[None], outs["1"], rtol=3e-14, atol=0)
    want = g.copy()
    refs.orc().orc_scale_modes(want.ctypes.data_as(C.c_void_p), 1, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm)
    np.testing.assert_allclose(outs[None], want, rtol=1e-10, atol=0)
    assert not np.array_equal(want, g)


@pytest.mark.parametrize("n,start,nslab", [(2048, 0, 2), (2048, 1023, 4), (4096, 2047, 2), (1024, 100, 4), (256, 0, 256), (64, 0, 64)])
def test_k3_float_grids_with_the_factor_in_float(gpu, n, start, nslab):
    """Float grids, smooth table (|B| <= 1/64, |norm ratio| <= 1/32): the factor - 1 is evaluated in float and applied as
    fmaf(x, delta, x) -- no FP64 instruction in the pass.  Against the double evaluation (KSN_K3_EXACT=1) the corrected
    modes agree to one float rounding and mostly bit for bit; against the oracle (float build) to the float-grid tolerance."""
    from kspace_neutrinos_b200 import capi
    box = refs.BOX
    rng = np.random.default_rng(n + 7 * start)
    g = rng.standard_normal((nslab, n, n // 2 + 1, 2)).astype(np.float32)
    logkk, ratio, norm = _smooth_table(n, box)
    outs = {}
    for knob in ("1", None):
        with _env(KSN_K3_EXACT=knob):
            d = refs.DeviceBuffer(gpu, g)
            capi.check(gpu.ksn_scale_modes(d.ptr, 4, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
            outs[knob] = d.download(g)
            d.free()
            name = gpu.ksn_last_k3_kernel()
            assert (b"factor in float" in name) == (knob is None), name
    np.testing.assert_allclose(outs[None], outs["1"], rtol=1.3e-7, atol=0)
    assert np.mean(outs[None] == outs["1"]) > 0.9
    want = g.copy()
    refs.orc().orc_scale_modes(want.ctypes.data_as(C.c_void_p), 0, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm)
    np.testing.assert_allclose(outs[None], want, rtol=1e-5, atol=0)
    # the plain-load kernel takes the same path: bit-identical
    with _env(KSN_K3_NOTMA="1"):
        d = refs.DeviceBuffer(gpu, g)
        capi.check(gpu.ksn_scale_modes(d.ptr, 4, n, start, nslab, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
        plain = d.download(g)
        d.free()
        assert b"plain loads" in gpu.ksn_last_k3_kernel() and b"factor in float" in gpu.ksn_last_k3_kernel()
    np.testing.assert_array_equal(plain, outs[None])


def test_k3_tables_with_knots_closer_than_the_finest_lookup_cells(gpu):
    """keff values are data-dependent means; nothing keeps two of them apart, and gsl_interp has no minimum spacing
    (delta_pow.c:19-37).  Knots closer than (range)/16384 in log2(k^2) used to be rejected; now the segment search steps
    over them."""
    from kspace_neutrinos_b200 import capi
    n, box = 128, refs.BOX
    rng = np.random.default_rng(5)
    g = rng.standard_normal((n, n, n // 2 + 1, 2))
    logkk, ratio, norm = _table(n, box, nk=30)
    # three pairs of nearly coincident knots, one of them straddling an integer k^2
    logkk = np.sort(np.concatenate([logkk, logkk[[5, 11, 20]] + np.array([1e-9, 3e-7, 1e-12])]))
    ratio = 0.2 + 0.6 * rng.random(len(logkk))
    d = refs.DeviceBuffer(gpu, g)
    capi.check(gpu.ksn_scale_modes(d.ptr, 8, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm))
    got = d.download(g)
    d.free()
    want = g.copy()
    refs.orc().orc_scale_modes(want.ctypes.data_as(C.c_void_p), 1, n, 0, n, box, refs.dptr(logkk), refs.dptr(ratio), len(logkk), norm)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=0)
