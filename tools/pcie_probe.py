#!/usr/bin/env python3
"""Host <-> device copy rates with every rank copying at once (one process per GPU, torchrun): what the e2e leg of
bench.py (host-resident slabs) can get from this box.  Pinned buffers of 2 GiB per rank, H2D alone, D2H alone, both at once."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
nbytes = 2 << 30
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return 3 * nbytes / t.item() / 1e9


for _ in range(2):
    res = {"h2d_only_GBps_per_gpu": run(True, False), "d2h_only_GBps_per_gpu": run(False, True), "both_GBps_per_gpu_each_way": run(True, True)}
if rank == 0:
    res.update({"gpus": world, "aggregate_h2d_GBps": res["h2d_only_GBps_per_gpu"] * world, "aggregate_both_GBps": 2 * res["both_GBps_per_gpu_each_way"] * world})
    print(res, flush=True)
if world > 1:
    dist.destroy_process_group()
