#!/bin/bash
# One gpurun call for the start of the next round: everything written after round 1's GPU minutes were spent.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_check.sh'
mkdir -p gpurun_out
O=gpurun_out
# 1. the reference's own cmocka programs against the product on the GPU (tests/test_zz_reference_programs_gpu.py)
timeout 200 python -m pytest tests/test_zz_reference_programs_gpu.py -q > $O/r2_reference_programs.log 2>&1
echo "reference programs exit $?" | tee -a $O/r2_reference_programs.log
# 2. K1 table cache (opt-in): whole GPU suite with it on, and the bench line with/without
KSN_K1_TABLE_CACHE=1 timeout 200 python -m pytest tests -q -m gpu -x > $O/r2_pytest_tabcache.log 2>&1
echo "suite (table cache) exit $?" | tee -a $O/r2_pytest_tabcache.log
timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_nocache.log 2>&1
KSN_K1_TABLE_CACHE=1 timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_tabcache.log 2>&1
tail -n 1 $O/r2_bench_nocache.log | cut -c1-400
tail -n 1 $O/r2_bench_tabcache.log | cut -c1-400
# 3. the whole suite as the driver runs it
timeout 200 python -m pytest tests -q -m gpu -x > $O/r2_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2_pytest_gpu.log
tail -n 4 $O/r2_pytest_gpu.log
