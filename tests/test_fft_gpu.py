"""The device-resident slab FFT in front of the hook (ksn_fft_*): real space in x-slabs -> k space in
y-slabs in FFTW's transposed order, the layout of  rfftwnd_mpi(..., FFTW_TRANSPOSED_ORDER)  in Gadget-2's pmforce_periodic
(gadget-2/0002 patch:116) -- against numpy's rfftn; the round trip; and the whole device-resident PM step (density ->
FFT -> add_nu_power_to_rhogrid on the y-slab -> Green's function) against the host pipeline with the CPU oracle.
More than one rank: processes that share cuda:0 (CUDA IPC works within one device), so the exchange kernel, the handle
exchange and the flag protocol are exercised on the one-GPU box too; tests/test_multi_gpu.py runs it on separate GPUs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import refs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def fft_atol(want):
    """round-off of a double-precision FFT whose largest output (the mean mode) dwarfs the rest: a few eps per butterfly
    level of the largest partial sums"""
    n3 = want.shape[1] ** 3
    return 3e-15 * np.log2(n3) * np.max(np.abs(want))


def transposed_rfftn(rho):
    """F[ky][kx][kz]: what FFTW_TRANSPOSED_ORDER leaves behind for rho[x][y][z]"""
    f = np.transpose(np.fft.rfftn(rho), (1, 0, 2))
    out = np.empty(f.shape + (2,))
    out[..., 0], out[..., 1] = f.real, f.imag
    return out


@pytest.mark.parametrize("n", [4, 8, 18, 30, 64, 96])
def test_one_rank_forward_matches_numpy_and_round_trips(gpu, n):
    from kspace_neutrinos_b200 import host
    rng = np.random.default_rng(n)
    rho = rng.standard_normal((n, n, n))
    fft = host.SlabFFT(n)
    assert (fft.xslab, fft.yslab) == (host.Slab(0, n), host.Slab(0, n))
    fft.upload_real(rho)
    fft.forward()
    got = fft.download_kspace()
    want = transposed_rfftn(rho)
    np.testing.assert_allclose(got, want, rtol=0, atol=fft_atol(want))
    fft.inverse()
    back = fft.download_real()
    fft.free()
    np.testing.assert_allclose(back, rho * float(n) ** 3, rtol=0, atol=1e-11 * n ** 3)


def test_k1_on_the_fft_output_equals_the_reference_on_numpys(gpu):
    """The slab the transform leaves behind is what total_powerspectrum sweeps (powerspectrum.c:56-89 with i = the y index)."""
    from kspace_neutrinos_b200 import host
    ref = refs.ref_lib(True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    n = 48
    rho = 1.0 + 0.2 * np.random.default_rng(3).standard_normal((n, n, n))
    fft = host.SlabFFT(n)
    fft.upload_real(rho)
    fft.forward()
    m_n, m_p, m_c, m_k = refs.total_powerspectrum(gpu, np.empty((n, n, 1, 1)), n // 2, fn="total_powerspectrum_f64", pointer=fft.kspace)
    fft.free()
    r_n, r_p, r_c, r_k = refs.total_powerspectrum(ref, transposed_rfftn(rho), n // 2)
    assert m_n == r_n and np.array_equal(m_c[:m_n], r_c[:r_n])
    np.testing.assert_allclose(m_p[:m_n], r_p[:r_n], rtol=1e-10)


WORKER = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.environ["KSN_ROOT"])
import torch.distributed as dist
from kspace_neutrinos_b200 import capi, host
from tests import refs
from tests.test_fft_gpu import transposed_rfftn, fft_atol
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = int(os.environ["KSN_TEST_NGPU"])
L = capi.lib()
capi.check(L.ksn_init(rank % ngpu))            # one GPU per rank where there are enough, else the ranks share cuda:0
dist.init_process_group(backend="gloo")
host.init_p2p_from_torch(rank, world)
def gather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out
for n in (int(x) for x in os.environ["KSN_TEST_SIZES"].split(",")):
    rho = 1.0 + 0.1 * np.random.default_rng(100 + n).standard_normal((n, n, n))     # the same field on every rank
    fft = host.SlabFFT(n, rank, world, gather)
    assert fft.xslab == host.slab_partition(n, world)[rank] == fft.yslab
    xs, ys = fft.xslab, fft.yslab
    fft.upload_real(rho[xs.start:xs.start + xs.count])
    fft.forward()
    got = fft.download_kspace()
    want = transposed_rfftn(rho)
    np.testing.assert_allclose(got, want[ys.start:ys.start + ys.count], rtol=0, atol=fft_atol(want))
    # a second transform right behind the first: the barriers keep the rounds apart
    fft.upload_real(2 * rho[xs.start:xs.start + xs.count])
    fft.forward()
    np.testing.assert_allclose(fft.download_kspace(), 2 * want[ys.start:ys.start + ys.count], rtol=0, atol=2 * fft_atol(want))
    fft.inverse()
    np.testing.assert_allclose(fft.download_real(), 2 * rho[xs.start:xs.start + xs.count] * float(n) ** 3, rtol=0, atol=1e-10 * n ** 3)
    # the device-resident PM step on this rank's y-slab: FFT -> neutrino correction (+ the bin sums over NVLink peer memory)
    fft.upload_real(rho[xs.start:xs.start + xs.count])
    fft.forward()
    sim = host.KspaceNeutrinos(host.Cosmology(transfer_file=refs.default_transfer_file(), mnu=(0.15, 0.15, 0.15)), n, rank=rank)
    o = refs.orc()
    m = refs.orc_module(n, masses=(0.15, 0.15, 0.15))
    full = want.copy()
    for a in (0.01, 0.02):
        sim.add_nu_power_to_rhogrid(a, fft.kspace, ys)
        assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, full.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
    np.testing.assert_allclose(sim.delta_nu_last(), np.array([m.dtot.delta_nu_last[i] for i in range(m.dtot.nk)]), rtol=1e-10)
    sl = full[ys.start:ys.start + ys.count]
    np.testing.assert_allclose(fft.download_kspace(), sl, rtol=1e-10, atol=2 * fft_atol(want))
    dist.barrier()                                  # nobody frees a buffer a peer still has mapped
    fft.free()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}/{world} ok")
'''


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return max(1, sum(1 for l in out.splitlines() if l.startswith("GPU ")))
    except OSError:
        return 1


@pytest.mark.parametrize("world,sizes", [(2, "8,30,64"), (3, "16,50"), (4, "64")])
def test_slab_ranks_exchange_through_peer_memory(world, sizes, tmp_path):
    script = tmp_path / "fft_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KSN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", KSN_TEST_SIZES=sizes, KSN_TEST_NGPU=str(_ngpus() if _ngpus() >= world else 1))
    for k in ("LOCAL_RANK", "KSN_DEVICE"):
        env.pop(k, None)
    port = 29900 + (os.getpid() % 500) + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world
