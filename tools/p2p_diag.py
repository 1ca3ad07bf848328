#!/usr/bin/env python3
"""Per-phase timing of the step on every rank, peer-memory backend (fused / stand-alone round) against NCCL.
torchrun --nproc-per-node N tools/p2p_diag.py [pmgrid]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from kspace_neutrinos_b200 import capi, host
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.cuda.set_device(local)
L = capi.lib(); capi.check(L.ksn_init(local)); L.ksn_set_quiet(1)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
slab = host.slab_partition(n, world)[rank]
cosmo = host.Cosmology(transfer_file=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ics_transfer_99.dat"), mnu=(0.1, 0.1, 0.1), hybrid_neutrinos_on=1)
grid = host.DeviceGrid(n, slab); grid.fill_synthetic()
tm = capi.Timing()
for mode in ("p2p", "p2p-unfused", "nccl"):
    if mode == "nccl": host.init_nccl_from_torch(rank, world)
    else: host.init_p2p_from_torch(rank, world)
    if mode == "p2p-unfused": os.environ["KSN_P2P_UNFUSED"] = "1"
    else: os.environ.pop("KSN_P2P_UNFUSED", None)
    sim = host.KspaceNeutrinos(cosmo, n, rank=rank)
    sim.add_nu_power_to_rhogrid(cosmo.time_transfer, grid.ptr, slab)
    sim.seed_history(98)
    a = 0.98
    L.ksn_timing_enable(1)
    for i in range(6):
        a += 0.0002
        dist.barrier(); torch.cuda.synchronize()
        L.ksn_timing_reset()
        t0 = time.perf_counter()
        sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
        w = (time.perf_counter() - t0) * 1e3
        L.ksn_timing_get(C.byref(tm))
        print(f"[{mode} rank {rank}] step {i}: wall {w:7.3f} ms  k1 {tm.k1_ms:.3f}  k1_reduce {tm.k1_reduce_ms:.3f}  comm {tm.comm_ms:.3f}  k2 {tm.k2_ms:.3f}  k3 {tm.k3_ms:.3f}  launches {tm.launches}", flush=True)
    # back-to-back, no barrier between steps
    dist.barrier(); torch.cuda.synchronize()
    rows = []
    t00 = time.perf_counter()
    for i in range(12):
        a += 0.0002
        L.ksn_timing_reset()
        t0 = time.perf_counter()
        sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
        w = (time.perf_counter() - t0) * 1e3
        L.ksn_timing_get(C.byref(tm))
        rows.append(f"{w:.2f}(k1 {tm.k1_ms:.2f} red {tm.k1_reduce_ms:.3f} comm {tm.comm_ms:.3f} k2 {tm.k2_ms:.2f} k3 {tm.k3_ms:.2f})")
    print(f"[{mode} rank {rank}] 12 steps back to back: {(time.perf_counter() - t00) * 1e3 / 12:.3f} ms per step: " + " ".join(rows), flush=True)
    L.ksn_timing_enable(0)
    dist.barrier(); torch.cuda.synchronize()
    t00 = time.perf_counter()
    for i in range(12):
        a += 0.0002
        sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
    print(f"[{mode} rank {rank}] 12 steps back to back, library timing off: {(time.perf_counter() - t00) * 1e3 / 12:.3f} ms per step", flush=True)
    L.ksn_timing_enable(0)
dist.barrier()
dist.destroy_process_group()
