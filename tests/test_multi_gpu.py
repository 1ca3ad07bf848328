"""Slab-sharded run over several GPUs of one box, one process per GPU, with either collective backend -- the library's own
NCCL all-reduce of the bin sums, or the peer-memory exchange fused into the final-reduce kernel (csrc/ksn_p2p.cuh): every
rank must end up with the 1-rank answer -- mode counts exactly, P(k)/delta_nu/grid to 1e-10 -- and, with the peer-memory
backend, with the SAME BITS on every rank.  Skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.environ["KSN_ROOT"])
import torch, torch.distributed as dist
from kspace_neutrinos_b200 import capi, host
from tests import refs
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
L = capi.lib()
capi.check(L.ksn_init(local))
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
backend = os.environ["KSN_TEST_COMM"]
if backend == "p2p":
    host.init_p2p_from_torch(rank, world)          # no fall-back here: the test is about this backend
else:
    host.init_nccl_from_torch(rank, world)
assert L.ksn_comm_size() == world and L.ksn_comm_rank() == rank
# the stand-alone all-reduce of the active backend (host buffer in, host buffer out), longer than one mailbox slot
v = np.arange(40000, dtype=np.float64) * (rank + 1)
capi.check(L.ksn_comm_allreduce_host(refs.dptr(v), v.size))
assert np.array_equal(v, np.arange(40000, dtype=np.float64) * (world * (world + 1) // 2))
n = 64
g = refs.random_grid(n, seed=2024)
slab = host.slab_partition(n, world)[rank]
sub = np.ascontiguousarray(g[slab.start:slab.start + slab.count])
# K1 alone: total_powerspectrum on this rank's slab returns the GLOBAL spectrum on every rank
nret, P, Cn, K = refs.total_powerspectrum(L, sub, n // 2, startslab=slab.start, nslab=slab.count, fn="total_powerspectrum_f64")
o = refs.orc()
pw, kw = np.zeros(n // 2), np.zeros(n // 2)
cw = np.zeros(n // 2, dtype=np.int64)
nw = o.orc_total_powerspectrum(n, g.ctypes.data_as(C.c_void_p), 1, n // 2, 0, n, refs.dptr(pw), cw.ctypes.data_as(capi.c_longlong_p), refs.dptr(kw))
assert nret == nw and np.array_equal(Cn[:nret], cw[:nw])
np.testing.assert_allclose(P[:nret], pw[:nw], rtol=1e-10)
np.testing.assert_allclose(K[:nret], kw[:nw], rtol=1e-10)
if backend == "p2p":
    # contributions are added in rank order on every rank: identical bits everywhere, also when the reduce is not fused
    mine = torch.from_numpy(P.copy()).cuda()
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    assert all(torch.equal(e, mine) for e in everyone)
    # (the first sweep of a geometry runs the three-sum kernel; compare the cached-geometry sweeps with each other)
    nret1, P1, Cn1, K1 = refs.total_powerspectrum(L, sub, n // 2, startslab=slab.start, nslab=slab.count, fn="total_powerspectrum_f64")
    os.environ["KSN_P2P_UNFUSED"] = "1"
    nret2, P2, Cn2, K2 = refs.total_powerspectrum(L, sub, n // 2, startslab=slab.start, nslab=slab.count, fn="total_powerspectrum_f64")
    del os.environ["KSN_P2P_UNFUSED"]
    assert nret2 == nret1 == nret and np.array_equal(P2, P1)
    np.testing.assert_allclose(P1[:nret], P[:nret], rtol=1e-12)
# whole step, device-resident slab, several PM steps
sim = host.KspaceNeutrinos(host.Cosmology(transfer_file=refs.default_transfer_file(), mnu=(0.15, 0.15, 0.15), hybrid_neutrinos_on=1), n, rank=rank)
dev = refs.DeviceBuffer(L, sub)
m = refs.orc_module(n, masses=(0.15, 0.15, 0.15), hybrid=True)
want = g.copy()
for a in (0.01, 0.02, 0.0205, 0.2, 0.34, 0.35):
    sim.add_nu_power_to_rhogrid(a, dev.ptr, slab)
    assert o.orc_add_nu_power_to_rhogrid(C.byref(m), a, refs.BOX, want.ctypes.data_as(C.c_void_p), 1, n, 0, n) == 0
    assert (sim.state.ia, sim.state.nk) == (m.dtot.ia, m.dtot.nk)
    np.testing.assert_allclose(sim.delta_nu_last(), np.array([m.dtot.delta_nu_last[i] for i in range(m.dtot.nk)]), rtol=1e-10)
    np.testing.assert_allclose(dev.download(sub), want[slab.start:slab.start + slab.count], rtol=1e-10, atol=0)
dev.free()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}/{world} ok")
'''


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.parametrize("backend", ["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_sharded_step_matches_single_rank(world, backend, tmp_path):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KSN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", KSN_TEST_COMM=backend)
    port = 29700 + (os.getpid() % 1000) + world + (50 if backend == "p2p" else 0)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world
