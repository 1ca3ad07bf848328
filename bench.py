#!/usr/bin/env python3
"""bench.py -- PM k-modes/s of the per-PM-step k-space hot path (BASELINE.json metric).

A step = one add_nu_power_to_rhogrid call (P(k) binning K1 -> cross-rank sum -> linear-response
integral K2 at a ~100-row stored history -> mode scaling K3) on a synthetic Gaussian grid of
PMGRID^3 (default 2048^3, double), x-slab sharded over --gpus ranks (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                           the reference's CPU code on the host cores

Prints ONE JSON line on rank 0 (see the driver contract in the task statement).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PM k-modes/s (P(k) bin + nu rhogrid correction)"
UNIT = "modes/s"
TRANSFER = os.path.join(ROOT, "tests", "golden", "ics_transfer_99.dat")
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pmgrid", type=int, default=int(os.environ.get("KSN_BENCH_PMGRID", "2048")))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-greens", action="store_true", help="skip the fused Green's-function side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region (diagnostics)")
    ap.add_argument("--no-check", action="store_true", help="skip the end-of-run parity check against the CPU oracle (outside the timed region)")
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("KSN_REF_BUDGET_S", "420")),
                    help="--impl reference: wall-clock budget for the warm-up + timed steps; the full slab is timed when it fits "
                         "host memory and this budget, else fewer planes per rank (reported)")
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes per rank in the CPU sample (0 = auto)")
    # the other BASELINE.json configs (parity-test cases, measured beside the headline): e.g. --pmgrid 4096 --mnu 0.2,0.1,0.3
    # --no-hybrid for configs[4]; the CPU arm is fixed to the headline's neutrino set-up, so use --no-cpu-baseline with them
    ap.add_argument("--mnu", default="0.1,0.1,0.1", help="the three neutrino masses in eV")
    ap.add_argument("--no-hybrid", action="store_true", help="hybrid neutrinos off")
    args = ap.parse_args()
    args.mnu = tuple(float(x) for x in args.mnu.split(","))
    if len(args.mnu) != 3:
        ap.error("--mnu takes three masses")
    return args


def workload_config(n, gpus, extra=None, mnu=(0.1, 0.1, 0.1), hybrid=True):
    masses = "3x0.1 eV" if tuple(mnu) == (0.1, 0.1, 0.1) else "MNu = " + "/".join(f"{m:g}" for m in mnu) + " eV"
    cfg = {"workload": f"PMGRID={n}^3 double, x-slab sharded over {gpus} GPU(s), {masses}, hybrid neutrinos "
                       f"{'on (Vcrit=500, NuPartTime=0.333)' if hybrid else 'off'}, 99-row delta_tot history (a=0.98+)",
           "pmgrid": n, "stored_modes": n * n * (n // 2 + 1), "nrbins": n // 2,
           "parallelism": f"slab{gpus}", "l2": f"grid ({16 * n * n * (n // 2 + 1) / gpus / 1e9:.3g} GB per GPU) " + ("far exceeds" if 16 * n * n * (n // 2 + 1) / gpus > 1.26e9 else "vs") + " the 126 MB L2; no flush between steps"}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------- reference / CPU arm
def run_ref_bench(n, planes, ranks, steps, hybrid=1, timeout=3000, warmup=1, budget_s=0.0, mnu=(0.1, 0.1, 0.1)):
    """Time the reference's CPU path (oracle/_ref/ref_bench: reference sources + shims, forked ranks)."""
    out = subprocess.run([REF_BENCH, str(n), str(planes), str(ranks), str(steps), str(hybrid), TRANSFER, str(warmup), repr(float(budget_s))] +
                         [repr(float(m)) for m in mnu], capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_bench failed: " + out.stderr[-500:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_port_bench(n, planes, ranks, steps):
    """Fallback when oracle/_ref is absent: the oracle port (oracle/libksn_oracle.so), one forked worker per
    core, each on its own sub-slab (no cross-rank sum -- same arithmetic per mode)."""
    import multiprocessing as mp
    import numpy as np
    from tests import refs

    def work(r, q):
        o = refs.orc()
        m = refs.orc_module(n, masses=(0.1, 0.1, 0.1), hybrid=True)
        L = n // 2 + 1
        start = r * (n // ranks)
        rng = np.random.default_rng(r)
        g = rng.standard_normal((planes, n, L, 2))
        if r == 0:
            g[0, 0, 0] = (n ** 3, 0)
        o.orc_add_nu_power_to_rhogrid(C.byref(m), 0.01, 512000.0, g.ctypes.data_as(C.c_void_p), 1, n, start, planes)
        d = m.dtot
        nk, rows = d.nk, 98
        for k in range(nk):
            base = d.delta_tot[k * d.namax]
            for i in range(1, rows):
                d.delta_tot[k * d.namax + i] = base * (i + 1)
        for i in range(1, rows):
            d.scalefact[i] = np.log(0.01 * (i + 1))
        d.ia = rows
        ts = []
        for s in range(steps + 1):
            t0 = time.perf_counter()
            o.orc_add_nu_power_to_rhogrid(C.byref(m), 0.98 + 0.001 * (s + 1), 512000.0, g.ctypes.data_as(C.c_void_p), 1, n, start, planes)
            ts.append(time.perf_counter() - t0)
        q.put((r, ts[1:], int(d.n_evals)))

    q = mp.Queue()
    ps = [mp.Process(target=work, args=(r, q)) for r in range(ranks)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    per_step = [max(t[1][s] for t in res) for s in range(steps)]
    return {"N": n, "P": planes, "R": ranks, "steps": [{"total": t, "k1": None, "integral": None, "k3": None} for t in per_step]}


def mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def cpu_throughput(n, steps, planes=0, warmup=1, budget_s=0.0, full_slab=False, mnu=(0.1, 0.1, 0.1), hybrid=True):
    """modes/s of the CPU path on all host cores.  full_slab: every rank takes its whole slab of the PMGRID^3 grid when the
    grid fits host memory (and stays below 2^31 elements per rank, the reference's `int` indices) -- nothing extrapolated;
    ref_bench itself cuts the planes per rank after the first warm-up step if warmup+steps would not fit budget_s.
    Otherwise a bounded sample (P planes per rank, ~6 s of CPU work per step), grid passes scaled to the full slab."""
    ranks = min(os.cpu_count() or 1, 256)          # mini-MPI's limit (oracle/mini_mpi.c)
    while n % ranks:
        ranks -= 1
    whole = n // ranks
    plane_elems = n * (n // 2 + 1)
    if planes <= 0:
        grid_bytes = 16 * plane_elems * n
        if full_slab and whole * plane_elems < 2 ** 31 and mem_available_bytes() > 1.15 * grid_bytes + (8 << 30):
            planes = whole
        else:
            # ~0.2 us per mode per core for the two grid passes (measured 0.1-0.25) -> aim at ~6 s of CPU work per step:
            # long enough that the extrapolation to the full slab is not dominated by start-up noise
            planes = max(1, min(whole, int(6.0 / (plane_elems * 0.2e-6))))
    if os.path.exists(REF_BENCH):
        kind, r = "reference", run_ref_bench(n, planes, ranks, steps, hybrid=1 if hybrid else 0, warmup=warmup, budget_s=budget_s, mnu=mnu)
    else:
        kind, r = "port", run_port_bench(n, planes, ranks, steps)
    full, used = [], set()
    for s in r["steps"]:
        p_s = s.get("planes", planes)
        used.add(p_s)
        scale = whole / p_s
        if s.get("integral") is not None:
            full.append(s["total"] if p_s == whole else (s["k1"] + s["k3"]) * scale + s["integral"])
        else:
            full.append(s["total"] * scale)        # port: integral not separable; scaled with the grid (pessimistic for the CPU only by the integral share)
    t = statistics.median(full)
    modes = n * n * (n // 2 + 1)
    extrapolated = used != {whole}
    how = (f"whole slabs ({whole} planes per rank = the full {n}^3 grid, {16 * plane_elems * n / 1e9:.1f} GB): nothing extrapolated" if not extrapolated else
           f"{sorted(used)} planes per rank of {whole} (grid passes scaled to the full slab = extrapolated; integral timed in full)")
    sample = (f"{kind} CPU path, {ranks} forked ranks, {how}; nk={r.get('nk')}, Na={r.get('Na')}, {len(full)} timed step(s) after "
              f"{r.get('warmup', 1)} warm-up; grid = the GPU arm's counter-based synthetic field")
    return modes / t, t, {"kind": kind, "cores": ranks, "sample": sample, "raw": r, "extrapolated": extrapolated, "warmup": r.get("warmup", 1)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.pmgrid
    t0 = time.time()
    val, t_step, info = cpu_throughput(n, max(1, args.steps), args.cpu_planes, warmup=max(0, args.warmup), budget_s=args.ref_budget_s,
                                       full_slab=True, mnu=args.mnu, hybrid=not args.no_hybrid)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": info["warmup"], "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n, args.gpus, None, args.mnu, not args.no_hybrid),
            "collective": "MPI_Allreduce over the forked ranks (mini-MPI shim: shared memory)",
            "extrapolated": info["extrapolated"],
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import torch.distributed as dist
    from kspace_neutrinos_b200 import capi, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    L = capi.lib()
    capi.check(L.ksn_init(local_rank), "ksn_init")
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        comm_backend = host.init_comm_from_torch(rank, world)
    else:
        comm_backend = "none (one rank)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        capi.check(L.ksn_device_synchronize())

    n = args.pmgrid
    slab = host.slab_partition(n, world)[rank]
    modes_total = n * n * (n // 2 + 1)
    cosmo = host.Cosmology(transfer_file=TRANSFER, mnu=args.mnu, hybrid_neutrinos_on=0 if args.no_hybrid else 1)
    sim = host.KspaceNeutrinos(cosmo, n, rank=rank)
    grid = host.DeviceGrid(n, slab)
    grid.fill_synthetic()
    # first PM step at a = TimeTransfer initialises the integrator; then install a 98-row history
    sim.add_nu_power_to_rhogrid(cosmo.time_transfer, grid.ptr, slab)
    sim.seed_history(98)
    a = 0.98
    # < 0.009: the row is integrated every step but not kept -> steady state; and never past a = 1 (TimeMax), however
    # many steps the caller asks for
    n_calls = max(3, args.warmup) + args.steps + (0 if args.no_e2e else args.e2e_steps + 1) + 4
    da = min(0.001, (0.9995 - a) / n_calls)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
        time.sleep(1.0)                          # nvidia-smi needs ~1 s to deliver its first sample: wait BEFORE the
    barrier()                                    # warm-up, so that no GPU idles between the warm-up and the timed steps
    L.ksn_timing_enable(1)                       # (the phase timers' events are created on first use: in the warm-up)
    for _ in range(max(3, args.warmup)):         # (an idle second there cost the first timed step 15-40 ms)
        a += da
        sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
    if rank == 0:
        sampler.rows.clear()                     # keep only samples taken during the timed region
    stream = torch.cuda.ExternalStream(L.ksn_stream())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.ksn_timing_reset()
    barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    step_wall = []
    for _ in range(args.steps):
        a += da
        ts = time.perf_counter()
        sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)       # (returns after K3: the entry is synchronous, like the reference's)
        step_wall.append((time.perf_counter() - ts) * 1e3)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = ev0.elapsed_time(ev1)
    tm = capi.Timing()
    L.ksn_timing_get(C.byref(tm))
    L.ksn_timing_enable(0)
    t = torch.tensor([dev_ms, wall * 1e3, tm.k1_ms, tm.k2_ms, tm.k3_ms, tm.comm_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, k1_ms, k2_ms, k3_ms, comm_ms = t.tolist()
    step_ms = dev_ms / args.steps
    value = modes_total / (step_ms * 1e-3)
    launches = int(tm.launches)
    k3_name = L.ksn_last_k3_kernel().decode()           # the instantiation the timed steps launched
    k1_name = L.ksn_last_k1_kernel().decode()

    # ---- what fusing the PM Green's function into K3 saves (one launch each, N=1 only; not part of `value`)
    greens = None
    if world == 1 and not args.no_greens:
        try:
            import numpy as np
            thr = C.POINTER(C.c_uint)()
            iw = capi.c_double_p()
            L.ksn_bin_tables(n, n // 2, C.byref(thr), C.byref(iw))
            nk = n // 2                            # as many knots as the step's own table has bins
            logkk = np.log(np.geomspace(1.0, n * 0.86, nk) * 2 * math.pi / cosmo.box_size)
            ratio, zero = np.linspace(0.9, 0.1, nk), np.zeros(nk)
            dp = lambda x: x.ctypes.data_as(capi.c_double_p)    # noqa: E731
            asmth2 = (2 * math.pi * 1.25 / n) ** 2
            tg = capi.Timing()

            def one(call):
                ts = []
                for _ in range(3):
                    L.ksn_timing_enable(1)
                    L.ksn_timing_reset()
                    capi.check(call())
                    L.ksn_timing_get(C.byref(tg))
                    ts.append(tg.k3_ms)
                L.ksn_timing_enable(0)
                return statistics.median(ts)
            k3_plain = one(lambda: L.ksn_scale_modes(grid.ptr, 8, n, slab.start, slab.count, cosmo.box_size, dp(logkk), dp(ratio), nk, 0.01))
            k3_fused = one(lambda: L.ksn_scale_modes_greens(grid.ptr, 8, n, slab.start, slab.count, cosmo.box_size, dp(logkk), dp(ratio), nk, 0.01, iw, asmth2))
            g_alone = one(lambda: L.ksn_scale_modes_greens(grid.ptr, 8, n, slab.start, slab.count, cosmo.box_size, dp(logkk), dp(zero), nk, 0.0, iw, asmth2))
            greens = {"k3_ms": k3_plain, "k3_with_greens_fused_ms": k3_fused, "greens_as_its_own_pass_ms": g_alone,
                      "saved_ms_per_step": k3_plain + g_alone - k3_fused, "saved_bytes_per_mode": 32,
                      "note": "K3 fused with the Green's-function multiply of the surrounding Gadget PM step (pm_periodic.c); the host's separate pass disappears"}
            grid.fill_synthetic()                   # the Green's function was applied nine times: start again from a sane grid
        except capi.KsnError as exc:
            greens = {"error": str(exc)}

    # ---- end to end: the same call on HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e and args.e2e_steps > 0:
        pin, err = None, None
        try:
            pin = host.PinnedGrid(n, slab)
            capi.check(L.ksn_memcpy_d2h(pin.ptr, grid.ptr, grid.nbytes))
        except capi.KsnError as exc:
            err = str(exc)
        # page-locking tens of GB takes seconds and varies by rank; and whatever fails on one rank (allocation, a peer that
        # never arrives) must not leave the others waiting in a collective: every rank takes the same sequence of
        # torch.distributed calls, and the verdict after each step is shared
        def everyone_ok(e):
            ok = torch.tensor([0.0 if e else 1.0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            return ok.item() >= 1.0

        def host_step(a_now):
            try:
                sim.add_nu_power_to_rhogrid(a_now, pin.ptr, slab)
                return None
            except capi.KsnError as exc:
                return str(exc)

        if not everyone_ok(err):
            e2e = {"value": None, "unit": UNIT, "error": err or "another rank could not allocate its pinned host slab"}
        else:
            barrier()
            a += da
            err = host_step(a)                                         # warm-up (allocates the staging buffer)
            good = everyone_ok(err)
            e_ms = 0.0
            if good:
                barrier()
                te = time.perf_counter()
                for _ in range(args.e2e_steps):
                    a += da
                    err = host_step(a)
                    good = everyone_ok(err)                            # (a few microseconds beside a step of seconds)
                    if not good:
                        break
                if good:
                    barrier()
                    e_ms = (time.perf_counter() - te) * 1e3
            if good:
                te_t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
                e_step = te_t.item() / args.e2e_steps
                e2e = {"value": modes_total / (e_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": grid.nbytes * world,
                       "d2h_bytes_per_step": grid.nbytes * world, "ms_per_step": e_step, "steps": args.e2e_steps,
                       "note": "add_nu_power_to_rhogrid on pinned HOST slabs: upload, K1, K2, K3, download, every step"}
            else:
                e2e = {"value": None, "unit": UNIT, "error": err or "the step failed on another rank"}
        # what the host <-> device links of this box give when every rank copies at once (the ceiling of the leg above:
        # the upload must be complete before K2 and K3 can start, so a step cannot beat upload time + download time)
        if pin is not None and e2e and e2e.get("value"):
            try:
                probe = min(grid.nbytes, 4 << 30)
                rates = []
                for fn, dst, src in ((L.ksn_memcpy_h2d, grid.ptr, pin.ptr), (L.ksn_memcpy_d2h, pin.ptr, grid.ptr)):
                    capi.check(fn(dst, src, probe))                    # warm
                    barrier()
                    tp = time.perf_counter()
                    capi.check(fn(dst, src, probe))
                    dt_p = torch.tensor([time.perf_counter() - tp], dtype=torch.float64, device="cuda")
                    if world > 1:
                        dist.all_reduce(dt_p, op=dist.ReduceOp.MAX)
                    rates.append(probe / dt_p.item() / 1e9)
                bound_ms = (grid.nbytes / rates[0] + grid.nbytes / rates[1]) / 1e6
                e2e.update({"h2d_GBps_per_gpu_all_ranks_copying": rates[0], "d2h_GBps_per_gpu_all_ranks_copying": rates[1],
                            "transfer_bound_ms_per_step": bound_ms, "frac_of_transfer_bound": bound_ms / e2e["ms_per_step"],
                            "bound_note": "upload and download of a step cannot overlap (K2 needs the whole grid's P(k) before K3 scales the first mode): "
                                          "ceiling = slab bytes / H2D rate + slab bytes / D2H rate, rates measured with every rank copying at once"})
            except capi.KsnError as exc:
                e2e["transfer_probe_error"] = str(exc)
        if pin is not None:
            pin.free()

    # ---- parity check (outside every timed region): one more PM step whose K1 / K2 / K3 results are re-derived by the CPU
    # oracle from the same inputs (tests/bench_check.py); every rank takes the step, rank 0 compares
    check = None
    if not args.no_check:
        a += da
        if rank == 0:
            try:
                from tests import bench_check
                check = bench_check.check_step(L, sim, grid, slab, n, a, args.mnu, not args.no_hybrid, world)
            except Exception as exc:  # noqa: BLE001 - report, never lose the measured line
                check = {"ok": False, "error": repr(exc)}
        else:
            sim.add_nu_power_to_rhogrid(a, grid.ptr, slab)
        barrier()

    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    local_modes = host.modes_in_slab(n, slab)
    k3_launch_ms = k3_ms / args.steps
    k1_launch_ms = k1_ms / args.steps
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"k3_{n}")            # one launch over the whole grid (ncu, N=1)
        if traffic is not None:
            traffic = traffic * local_modes / modes_total   # a launch on this rank's slab
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": k3_name + " (read-modify-write, 32 B per stored mode)",
                "achieved": 32.0 * local_modes / (k3_launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": 32.0 * local_modes / (k3_launch_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "k1": {"kernel": k1_name + " (read-only, 16 B per stored mode)",
                       "achieved": 16.0 * local_modes / (k1_launch_ms * 1e-3) / 1e9,
                       "frac": 16.0 * local_modes / (k1_launch_ms * 1e-3) / 1e9 / peak},
                "step": {"achieved": 48.0 * local_modes / (step_ms * 1e-3) / 1e9,
                         "frac": 48.0 * local_modes / (step_ms * 1e-3) / 1e9 / peak,
                         "frac_of_nominal_8TBs": 48.0 * local_modes / (step_ms * 1e-3) / 1e9 / 8000.0},
                "k2_ms_per_step": k2_ms / args.steps, "comm_ms_per_step": comm_ms / args.steps,
                "k2_kernel": f"speculation width {L.ksn_k2_spec_width() or 'by regime (hybrid, one species: 3)'}, slowest bin {L.ksn_last_k2_max_trips()} passes through the integrand"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(n, world, None, args.mnu, not args.no_hybrid), "collective": comm_backend, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "wall_ms_per_step": wall_ms / args.steps,
            "step_wall_ms_rank0": {"min": min(step_wall), "median": statistics.median(step_wall), "max": max(step_wall),
                                   "slowest_step": step_wall.index(max(step_wall)), "all": [round(x, 3) for x in step_wall]}}
    if check is not None:
        line["parity_check"] = check
    if greens is not None:
        line["greens_fusion"] = greens
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, t_step, info = cpu_throughput(n, 1, args.cpu_planes, mnu=args.mnu, hybrid=not args.no_hybrid)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                                    "s_per_step": t_step}
        except Exception as exc:  # noqa: BLE001 - report, never fail the GPU number on the CPU leg
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": repr(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        pass


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
