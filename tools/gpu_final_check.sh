#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r1_bench_n1_final_device.log 2>&1
tail -n 1 $O/r1_bench_n1_final_device.log | cut -c1-2200
timeout 100 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_final.log 2>&1
echo "suite exit $?" | tee -a $O/pytest_gpu_final.log
tail -n 6 $O/pytest_gpu_final.log
