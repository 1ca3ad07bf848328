"""world_size-2 (and 3) runs of the multi-rank host logic over torch.distributed/gloo on CPU: slab partition,
the host all-reduce backend of the C library (the path an MPI host uses) and the normalise/compact tail.
Per-rank bin sums come from the CPU oracle here (the checker stands in for K1, which needs a GPU); the
N-rank result must equal the 1-rank result."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.environ["KSN_ROOT"])
import torch.distributed as dist
from kspace_neutrinos_b200 import capi, host
from tests import refs
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, nrbins = 32, 16
L = capi.lib()
host.init_host_allreduce_from_torch(rank, world)
assert (L.ksn_comm_rank(), L.ksn_comm_size()) == (rank, world)
g = refs.random_grid(n, seed=99)
slab = host.slab_partition(n, world)[rank]
sub = np.ascontiguousarray(g[slab.start:slab.start + slab.count])
o = refs.orc()
p, k = np.zeros(nrbins), np.zeros(nrbins)
c = np.zeros(nrbins, dtype=np.int64)
m2 = C.c_double()
o.orc_powerspectrum_sums(n, sub.ctypes.data_as(C.c_void_p), 1, nrbins, slab.start, slab.count, refs.dptr(p), refs.dptr(k),
                         c.ctypes.data_as(capi.c_longlong_p), C.byref(m2))
# the reduce buffer layout of the library: power | mass2 | keff | count (counts exact as doubles)
buf = np.concatenate([p, [m2.value], k, c.astype(np.float64)])
capi.check(L.ksn_comm_allreduce_host(buf.ctypes.data_as(capi.c_double_p), len(buf)))
nret, P, Cn, K = host.finish_powerspectrum(buf[:nrbins], buf[nrbins + 1:2 * nrbins + 1], np.rint(buf[2 * nrbins + 1:]).astype(np.int64), buf[nrbins])
# single-rank truth
pw, kw = np.zeros(nrbins), np.zeros(nrbins)
cw = np.zeros(nrbins, dtype=np.int64)
nw = o.orc_total_powerspectrum(n, g.ctypes.data_as(C.c_void_p), 1, nrbins, 0, n, refs.dptr(pw), cw.ctypes.data_as(capi.c_longlong_p), refs.dptr(kw))
assert nret == nw and np.array_equal(Cn[:nret], cw[:nw]), (nret, nw)
np.testing.assert_allclose(P[:nret], pw[:nw], rtol=1e-13)
np.testing.assert_allclose(K[:nret], kw[:nw], rtol=1e-13)
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}/{world} ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_host_allreduce_backend_over_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KSN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 2000) + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world


FALLBACK_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KSN_ROOT"])
import torch.distributed as dist
from kspace_neutrinos_b200 import capi, host
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# No GPU here: exporting a mailbox fails on every rank.  The choice of backend must stay COLLECTIVE (every rank goes
# through the same gathers and agreements, nobody hangs) and end in the NCCL branch -- which, without a device, fails loudly.
calls = []
host.init_nccl_from_torch = lambda r, w: calls.append((r, w))
got = host.init_comm_from_torch(rank, world)
assert got == "nccl" and calls == [(rank, world)], (got, calls)
assert capi.lib().ksn_comm_size() == 1            # nothing half-initialised is left behind
# asked for NCCL outright: no peer-memory attempt at all
calls.clear()
assert host.init_comm_from_torch(rank, world, backend="nccl") == "nccl" and calls == [(rank, world)]
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}/{world} ok")
'''


def test_backend_choice_is_collective_and_falls_back(tmp_path):
    """host.init_comm_from_torch on a box without usable peer memory (here: no GPU at all): every rank must take the same
    path through the gathers/agreements and fall back to NCCL together."""
    script = tmp_path / "worker.py"
    script.write_text(FALLBACK_WORKER)
    env = dict(os.environ, KSN_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 2000) + 7
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 2
    assert "peer-memory backend not usable on every rank" in r.stderr
