#!/bin/bash
# Round 2, third GPU call (one GPU):  gpurun --timeout 1200 -- 'bash tools/gpu_r2_call3.sh'
# Slab FFT (one rank, and ranks sharing cuda:0 through CUDA IPC), K3's branch-free fast path, whole suite, timings.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_fft_gpu.py -q -x > $O/r2c_fft.log 2>&1
echo "fft tests exit $?" | tee -a $O/r2c_fft.log
tail -n 25 $O/r2c_fft.log | cut -c1-250
timeout 600 python -m pytest tests -q -m gpu > $O/r2c_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2c_pytest_gpu.log
tail -n 30 $O/r2c_pytest_gpu.log | cut -c1-250
timeout 120 python tools/quick_bench.py 2048 6 8 512 greens 2>&1 | grep "K1 fast\|^K3" | tail -n 9 | cut -c1-230 | tee $O/r2c_quick_2048_f64.txt
KSN_K3_NOFAST=1 timeout 120 python tools/quick_bench.py 2048 6 8 512 greens 2>&1 | grep "^K3" | tail -n 9 | cut -c1-230 | tee $O/r2c_quick_2048_f64_nofast.txt
KSN_K3_EXACT=1 timeout 120 python tools/quick_bench.py 2048 6 8 512 greens 2>&1 | grep "^K3" | tail -n 9 | cut -c1-230 | tee $O/r2c_quick_2048_f64_exact.txt
timeout 120 python tools/quick_bench.py 2048 6 4 512 2>&1 | grep "K1 fast\|^K3" | tail -n 4 | cut -c1-230 | tee $O/r2c_quick_2048_f32.txt
for n in 256 512 1024; do
  timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "^K3" | tail -n 1 | sed "s/^/PMGRID $n f64: /" | cut -c1-230
  timeout 120 python tools/quick_bench.py $n 5 4 2>&1 | grep "^K3" | tail -n 1 | sed "s/^/PMGRID $n f32: /" | cut -c1-230
done | tee $O/r2c_small_grids.txt
timeout 300 python tools/pm4096_probe.py 384 2>&1 | grep -h "^K3" | cut -c1-230 | tee $O/r2c_pm4096_probe.txt
timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2c_step_f32.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2c_bench_n1.log 2>&1
echo "bench exit $?"; tail -n 1 $O/r2c_bench_n1.log | cut -c1-3500
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k3_scale_row_kernel -s 2 -c 1 -o $O/r2c_k3_row_2048 python tools/quick_bench.py 2048 4 8 256 > $O/r2c_ncu_k3.log 2>&1
timeout 400 $NCU -k regex:k3_scale_row_kernel -s 5 -c 1 -o $O/r2c_k3_row_greens_2048 python tools/quick_bench.py 2048 4 8 256 greens > $O/r2c_ncu_k3g.log 2>&1
timeout 400 $NCU -k regex:k3_scale_flat_kernel -s 2 -c 1 -o $O/r2c_k3_flat_f32_2048 python tools/quick_bench.py 2048 4 4 256 > $O/r2c_ncu_k3f.log 2>&1
ls -la $O/r2c*.ncu-rep
