"""The premise of the roofline target: the grid never leaves HBM.  A PM step as a GPU-resident code
would run it -- density field -> forward r2c FFT by cuFFT (through torch.fft, plumbing only) -> the library's fused
neutrino correction + Green's function on the DEVICE pointer -> inverse FFT -- against the same pipeline on the host with
numpy's FFT, the CPU oracle's add_nu_power_to_rhogrid and the numpy restatement of Gadget-2's Green's-function loop."""
import ctypes as C

import numpy as np
import pytest

from kspace_neutrinos_b200 import capi
from tests import refs

pytestmark = pytest.mark.gpu


def test_device_resident_pm_step_matches_host_pipeline(gpu):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device for torch")
    n = 64
    asmth2 = (2 * np.pi * 1.25 / n) ** 2
    rng = np.random.default_rng(5)
    rho = 1.0 + 0.1 * rng.standard_normal((n, n, n))            # mean 1: F(0,0,0) = n^3, the "total mass" K1 normalises by
    times = (0.01, 0.02, 0.03)

    # host pipeline: numpy FFT -> oracle step -> Green's function -> inverse FFT
    o = refs.orc()
    m = refs.orc_module(n, masses=(0.15, 0.15, 0.15))
    want = []
    for a in times:
        f = np.fft.rfftn(rho)
        g = np.empty((n, n, n // 2 + 1, 2))
        g[..., 0], g[..., 1] = f.real, f.imag
        assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, g.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
        g = refs.greens_numpy(g, 0, asmth2)
        want.append(np.fft.irfftn(g[..., 0] + 1j * g[..., 1], s=(n, n, n), axes=(0, 1, 2)))

    # device pipeline: the grid is produced, corrected and consumed in HBM
    refs.init_module(gpu, n, masses=(0.15, 0.15, 0.15))
    dt = capi.global_delta_tot_table()
    drho = torch.from_numpy(rho).cuda()
    got = []
    for a in times:
        f = torch.fft.rfftn(drho).contiguous()                    # complex128 [n][n][n/2+1], re/im interleaved: the library's layout
        assert f.dtype == torch.complex128 and f.shape == (n, n, n // 2 + 1)
        torch.cuda.synchronize()                                  # the library launches on its own stream
        assert gpu.ksn_pointer_is_device(C.c_void_p(f.data_ptr())) == 1
        gpu.add_nu_power_and_greens_to_rhogrid_f64(a, refs.BOX, C.c_void_p(f.data_ptr()), n, 0, n, asmth2, 0)
        got.append(torch.fft.irfftn(f, s=(n, n, n)).cpu().numpy())
    assert dt.ia == m.dtot.ia == 3
    for g_, w_ in zip(got, want):
        scale = np.max(np.abs(w_))
        assert scale > 0
        np.testing.assert_allclose(g_, w_, rtol=0, atol=1e-10 * scale)
