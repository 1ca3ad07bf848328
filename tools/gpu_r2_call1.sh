#!/bin/bash
# Round 2, first GPU call (one GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
# 1. the new full-width parity tests (reference sources on the same bytes at PMGRID 1024/2048/4096; GADGET-2 Green's loop)
# 2. everything tools/gpu_round2_check.sh lists (opt-in kernels' parity tests and timings, whole suite)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > $O/r2_gpu.txt; nproc >> $O/r2_gpu.txt; free -g | head -2 >> $O/r2_gpu.txt
timeout 600 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_k3_gpu.py -q -m gpu > $O/r2_fullsize.log 2>&1
echo "full-width parity exit $?" | tee -a $O/r2_fullsize.log
tail -n 30 $O/r2_fullsize.log | cut -c1-300
KSN_TEST_UNVERIFIED=1 timeout 500 python -m pytest tests/test_zz_optin_gpu.py -q > $O/r2_optin.log 2>&1
echo "opt-in kernels exit $?" | tee -a $O/r2_optin.log
tail -n 40 $O/r2_optin.log | cut -c1-300
timeout 300 python -m pytest tests -q -m gpu -x > $O/r2_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2_pytest_gpu.log
tail -n 6 $O/r2_pytest_gpu.log | cut -c1-300
# ---- timings (from tools/gpu_round2_check.sh)
timeout 120 python tools/quick_bench.py 2048 5 4 > $O/r2_f32_default.log 2>&1
KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 120 python tools/quick_bench.py 2048 5 4 > $O/r2_f32_optin.log 2>&1
grep -h "K1 fast\|K3" $O/r2_f32_default.log | tail -n 2 | cut -c1-200
grep -h "K1 fast\|K3" $O/r2_f32_optin.log | tail -n 2 | cut -c1-200
for shape in "8,17,2" "8,33,2" "8,33,3" "12,9,3"; do
  echo "float tile shape $shape:"; KSN_K1_TILE=$shape KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 120 python tools/quick_bench.py 2048 4 4 2>&1 | grep "K1 fast" | tail -n 1 | cut -c1-200
done 2>&1 | tee $O/r2_f32_shapes.txt
timeout 300 python tools/pm4096_probe.py 384 > $O/r2_pm4096_probe.log 2>&1
grep -h "^K1\|^K3" $O/r2_pm4096_probe.log | cut -c1-200
: > $O/r2_k2_cluster.txt
for cl in 0 2 3 4; do
  for cfg in "788 1" "788 1 nondegenerate" "788 0"; do
    echo "--- KSN_K2_CLUSTER=$cl  k2_bench.py $cfg" >> $O/r2_k2_cluster.txt
    KSN_K2_CLUSTER=$cl timeout 120 python tools/k2_bench.py $cfg 2>&1 | tail -n 3 >> $O/r2_k2_cluster.txt
  done
done
cut -c1-220 $O/r2_k2_cluster.txt
for n in 256 512 1024; do
  timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "K1 fast\|^K3" | tail -n 2 | sed "s/^/PMGRID $n: /" | cut -c1-200
done | tee $O/r2_small_grids.txt
for n in 512 1024; do
  KSN_K3_FLAT=1 timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "^K3" | tail -n 1 | sed "s/^/PMGRID $n KSN_K3_FLAT=1: /" | cut -c1-200
done | tee -a $O/r2_small_grids.txt
timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_default.txt
KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_optin.txt
KSN_K3_F32_TMA=2 KSN_K1_F32_TILE=1 timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_optin_short.txt
timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_nocache.log 2>&1
KSN_K1_TABLE_CACHE=1 timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_tabcache.log 2>&1
tail -n 1 $O/r2_bench_nocache.log | cut -c1-300
tail -n 1 $O/r2_bench_tabcache.log | cut -c1-300
