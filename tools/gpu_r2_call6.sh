#!/bin/bash
# Round 2, one GPU, after the host-side table caching:  gpurun --timeout 900 -- 'bash tools/gpu_r2_call6.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_k3_gpu.py tests/test_step_gpu.py tests/test_promoted_kernels_gpu.py tests/test_fullsize_parity_gpu.py -q -x -k "not mpi_build" 2>&1 | tail -n 3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-greens > gpurun_out/r2v_bench_n1.json 2>gpurun_out/r2v_bench_n1.err
echo "bench exit $?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2v_bench_n1.json").read().strip().splitlines()[-1]); r = d["roofline"]
n = d["config"]["stored_modes"]
k1 = 16 * n / r["k1"]["achieved"] / 1e6; k3 = 32 * n / r["achieved"] / 1e6
print(d["ms_per_step"], "k1", k1, "k3", k3, "k2", r["k2_ms_per_step"], "gaps", d["ms_per_step"] - k1 - k3 - r["k2_ms_per_step"], d["parity_check"]["ok"], d["step_wall_ms_rank0"]["median"])
P
