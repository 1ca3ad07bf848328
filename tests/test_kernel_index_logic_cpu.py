"""Index arithmetic of the float-grid kernels, restated in Python and checked against direct enumeration (CPU, no GPU):
the kernels themselves are tested on the GPU (tests/test_promoted_kernels_gpu.py); the part of them that is pure integer
logic -- which bytes a bulk copy covers, where a thread finds its mode -- is pinned here.

* k3_scale_flat_kernel (csrc/k3_scale.cu): the slab is cut into flat chunks of an even mode count; a thread derives
  (plane, row, z) of mode e of its chunk from the chunk's first mode.
* k1_tile_kernel<float, ..> (csrc/k1_powerspec.cu): a tile of an odd float row starts 8 bytes off the 16-byte granule of
  cp.async.bulk; it is copied from one mode earlier and to an even mode count, and lane l reads its chunk `sh` modes into
  the stage."""
import pytest

K3_EPT = 9
K3_FLAT_MAX_ROWS = 64          # rows a chunk may touch (csrc/k3_scale.cu)


def k3_flat_chunk(L, threads):
    """k3_launch: whole row pairs (~16-18 KB of them) where a pair fits one CTA, else the largest even piece."""
    capf, pair = threads * K3_EPT, 2 * L
    if pair <= capf:
        return pair * max(1, min(capf // pair, (K3_FLAT_MAX_ROWS - 2) // 2, 18432 // (pair * 8)))
    return capf & ~1


@pytest.mark.parametrize("n", [4, 6, 8, 16, 30, 64, 126, 128, 254])
@pytest.mark.parametrize("nplanes", [1, 2, 3])
def test_k3_flat_chunks_cover_the_slab_and_find_every_mode(n, nplanes):
    L, plane0 = n // 2 + 1, 5
    total = nplanes * n * L
    if total % 2:
        pytest.skip("odd mode total: the host keeps the plain kernel")
    chunk = k3_flat_chunk(L, 256)
    assert chunk % 2 == 0 and chunk <= 256 * K3_EPT
    nct = (total + chunk - 1) // chunk
    seen = 0
    for b in range(nct):
        e0 = b * chunk
        nel = min(chunk, total - e0)
        assert e0 % 2 == 0 and nel % 2 == 0                       # 16-byte aligned start, whole 16-byte granules
        r0 = e0 // L
        zb = e0 - r0 * L
        pl0 = r0 // n
        j0 = r0 - pl0 * n
        rows = (zb + nel + L - 1) // L
        assert rows <= K3_FLAT_MAX_ROWS
        magic = (0xffffffff // L + 1) & 0xffffffff                # row of a mode by a multiply instead of a division
        rowtab = []
        for rl in range(rows):                                    # the kernel's per-row table (threads 0 .. rows-1)
            j, gi = j0 + rl, plane0 + pl0
            if j >= n:
                q = j // n
                j -= q * n
                gi += q
            rowtab.append((gi, j))
        for e in range(nel):                                      # the kernel's per-mode code
            zz = zb + e
            rl = (zz * magic) >> 32
            z = zz - rl * L
            gi, j = rowtab[rl]
            g = e0 + e
            row = g // L
            assert (gi - plane0, j, z) == (row // n, row % n, g % L)
            seen += 1
    assert seen == total


def test_k3_flat_chunk_sizes_at_the_named_grids():
    assert k3_flat_chunk(1025, 256) == 2050          # PMGRID 2048 float: two whole rows, 16400 B (the double kernel's block)
    assert k3_flat_chunk(2049, 256) == 2304          # PMGRID 4096 float: pieces that ignore row boundaries
    assert k3_flat_chunk(513, 256) == 2052           # PMGRID 1024 float: four rows


@pytest.mark.parametrize("n,C", [(4, 1), (8, 1), (16, 1), (64, 5), (64, 9), (96, 5), (128, 9), (256, 13), (256, 17), (126, 5), (2048, 33)])
def test_k1_float_tile_copies_are_aligned_in_bounds_and_indexed_right(n, C):
    L = n // 2 + 1
    TE = 32 * C
    T = (L + TE - 1) // TE
    stage_elems = (((TE + 8) * 8 + 127) // 128 * 128) // 8       # k1_tile_smem with 8 bytes per mode
    for nplanes in (1, 2):
        nrows = nplanes * n
        total = nrows * L
        assert total % 2 == 0                                    # (even PMGRID; the host keeps the scan kernel otherwise)
        rows = range(nrows) if n <= 256 else list(range(0, 9)) + list(range(nrows - 9, nrows))
        for r in rows:
            for t in range(T):
                z0 = t * TE
                ln = min(TE, L - z0)
                g0 = r * L + z0
                sh = g0 & 1
                nel = (ln + sh + 1) & ~1
                src = g0 - sh
                assert src % 2 == 0 and src >= 0 and src + nel <= total and nel <= stage_elems
                for lane in (0, 1, 15, 16, 31):
                    for e in (0, C - 1):
                        z = z0 + lane * C + e
                        idx = sh + lane * C + e                   # where the lane reads
                        assert idx < stage_elems
                        if z < L:
                            assert idx < nel and src + idx == r * L + z


@pytest.mark.parametrize("n,nrbins", [(8, 4), (64, 32), (96, 48)])
def test_k1_float_tile_arithmetic_stays_within_the_float_grid_tolerance(n, nrbins):
    """The float tile kernel forms |F|^2 in float (as the reference's float build does) but keeps the separable DOUBLE
    window (m_z iwz^4) (iwx iwy)^4, whereas the reference rounds iwx iwy iwz and its square to float (powerspectrum.c:8-24
    with fftw_real = float).  Restated in numpy on the product's own tables and compared with the compiled float
    reference: the deviation must stay far inside the north star's 1e-5 (measured: below 1e-6)."""
    import ctypes as C

    import numpy as np

    from kspace_neutrinos_b200 import capi
    from tests import refs
    ref = refs.ref_lib(False)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    L = capi.lib()
    g = refs.random_grid(n, seed=n + 1, dtype=np.float32)
    r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, g, nrbins)
    thr = C.POINTER(C.c_uint)()
    iw = capi.c_double_p()
    assert L.ksn_bin_tables(n, nrbins, C.byref(thr), C.byref(iw)) == 0
    t = np.array([thr[i] for i in range(nrbins)], dtype=np.int64)
    w = np.array([iw[i] for i in range(n // 2 + 1)])
    k = np.fft.fftfreq(n, 1.0 / n).round().astype(np.int64)
    kz = np.arange(n // 2 + 1)
    k2 = k[:, None, None] ** 2 + k[None, :, None] ** 2 + kz[None, None, :] ** 2
    b = np.searchsorted(t, k2, side="right") - 1
    mult = np.where((kz == 0) | (kz == n // 2), 1.0, 2.0)
    wz = mult * w[kz] ** 4
    wxy = (w[np.abs(k)][:, None] * w[np.abs(k)][None, :]) ** 4
    re, im = g[..., 0], g[..., 1]
    pp = (im * im + re * re).astype(np.float32).astype(np.float64)
    val = pp * wz[None, None, :] * wxy[:, :, None]
    val[0, 0, 0] = 0.0
    P = np.bincount(b.ravel(), weights=val.ravel(), minlength=nrbins)
    cnt = np.bincount(b[k2 > 0].ravel(), weights=(mult[None, None, :] * np.ones_like(k2))[k2 > 0].ravel(), minlength=nrbins)
    m2 = float(re[0, 0, 0]) ** 2 + float(im[0, 0, 0]) ** 2
    keep = cnt > 0
    assert r_n == keep.sum() and np.array_equal(r_c[:r_n], cnt[keep].astype(np.int64))
    np.testing.assert_allclose(P[keep] / m2 / cnt[keep], r_p[:r_n], rtol=2e-6)
