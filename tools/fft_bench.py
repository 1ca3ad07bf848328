#!/usr/bin/env python3
"""Developer benchmark of the device-resident slab FFT in front of the hook (ksn_fft_*) and of the PM step that starts
from a device-resident REAL-SPACE density: forward r2c FFT into the transposed-order y-slab -> add_nu_power_to_rhogrid
(K1, cross-rank sum, K2, K3) on it.  One process per GPU (torchrun), peer-memory backend.
   python tools/fft_bench.py PMGRID [reps]            or       torchrun --nproc-per-node N tools/fft_bench.py PMGRID [reps]
Check (size independent): inverse(forward(rho)) == N^3 rho on two sample planes per rank."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kspace_neutrinos_b200 import capi, host  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    L = capi.lib()
    capi.check(L.ksn_init(local))
    L.ksn_set_quiet(1)
    gather = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        host.init_p2p_from_torch(rank, world)

        def gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        def barrier():
            dist.barrier()
            capi.check(L.ksn_device_synchronize())

        def tmax(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
    else:
        def barrier():
            capi.check(L.ksn_device_synchronize())

        def tmax(x):
            return x
    fft = host.SlabFFT(n, rank, world, gather)
    xs, ys = fft.xslab, fft.yslab
    Lz = n // 2 + 1
    row = 2 * Lz * 8

    def fill():
        # the padded real grid seen as [nx][N][L] complex: the synthetic Gaussian field's values serve as a random density
        capi.check(L.ksn_fill_synthetic_grid(fft.real, 8, n, xs.start, xs.count, 7, 0.0))

    def sample(ptr, plane):
        out = np.empty((n, 2 * Lz))
        capi.check(L.ksn_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr.value + plane * n * row), out.nbytes))
        return out

    fill()
    planes = sorted({0, xs.count - 1})
    before = [sample(fft.real, p)[:, :n].copy() for p in planes]
    barrier()
    fft.forward()
    fft.inverse()
    barrier()
    worst = 0.0
    for p, b in zip(planes, before):
        a = sample(fft.real, p)[:, :n]
        worst = max(worst, float(np.max(np.abs(a / float(n) ** 3 - b)) / np.max(np.abs(b))))
    t_f, t_i = [], []
    stages = (C.c_float * 4)()
    for _ in range(reps):
        fill()
        barrier()
        t0 = time.perf_counter()
        fft.forward()
        barrier()
        t1 = time.perf_counter()
        L.ksn_fft_timing(stages)
        st_f = [float(x) for x in stages]
        fft.inverse()
        barrier()
        t2 = time.perf_counter()
        t_f.append(tmax(t1 - t0)); t_i.append(tmax(t2 - t1))
    # the PM step from a device-resident real-space density
    transfer = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ics_transfer_99.dat")
    cosmo = host.Cosmology(transfer_file=transfer, mnu=(0.1, 0.1, 0.1), hybrid_neutrinos_on=1)
    sim = host.KspaceNeutrinos(cosmo, n, rank=rank)
    fill()
    fft.forward()
    # (the field has no mean mode to normalise by: give it one, as a density has)
    if rank == 0:
        dc = np.array([float(n) ** 3, 0.0])
        capi.check(L.ksn_memcpy_h2d(fft.kspace, dc.ctypes.data_as(C.c_void_p), 16))
    sim.add_nu_power_to_rhogrid(cosmo.time_transfer, fft.kspace, ys)
    sim.seed_history(98)
    a = 0.98
    t_s = []
    for _ in range(reps + 2):
        a += 0.0005
        fill()
        barrier()
        t0 = time.perf_counter()
        fft.forward()
        sim.add_nu_power_to_rhogrid(a, fft.kspace, ys)
        barrier()
        t_s.append(tmax(time.perf_counter() - t0))
    if rank == 0:
        modes = n * n * Lz
        f_ms, i_ms, s_ms = np.median(t_f) * 1e3, np.median(t_i) * 1e3, np.median(t_s[2:]) * 1e3
        # bytes the transform must move per rank: 2-D pass r+w, exchange r+w, 1-D pass r+w of the 16 B/mode grid
        print(json.dumps({"pmgrid": n, "gpus": world, "forward_ms": f_ms, "inverse_ms": i_ms, "fft_plus_step_ms": s_ms,
                          "modes_per_s_fft_plus_step": modes / (s_ms * 1e-3), "forward_GBps_per_gpu_of_6_passes": 6 * 16 * modes / world / (f_ms * 1e-3) / 1e9,
                          "forward_stages_ms_rank0": {"wait_for_peers": st_f[0], "fft2d_with_exchange_behind": st_f[1], "closing_fence": st_f[2], "fft1d_along_x": st_f[3]},
                          "round_trip_max_rel_err": worst, "k1": L.ksn_last_k1_kernel().decode(), "k3": L.ksn_last_k3_kernel().decode()}), flush=True)
    barrier()
    fft.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
