#!/bin/bash
# Round 2, second GPU call (one GPU):  gpurun --timeout 1800 -- 'bash tools/gpu_r2_call2.sh'
# After the opt-in kernels were promoted / deleted and K3's factor modes were added: whole suite, bench line with the
# parity check, reference arm on full slabs, kernel timings, launch list and ncu --set full captures of the shipped kernels.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > $O/r2b_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2b_pytest_gpu.log
tail -n 30 $O/r2b_pytest_gpu.log | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2b_bench_n1.log 2>&1
echo "bench exit $?"; tail -n 1 $O/r2b_bench_n1.log | cut -c1-6000
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_ref.log 2>&1
echo "reference arm exit $?"; tail -n 1 $O/r2b_bench_ref.log | cut -c1-2500
# kernels alone, double and float grids (256-plane slabs of 2048: the kernels go by width), with the fused Green's function
timeout 120 python tools/quick_bench.py 2048 6 8 512 greens 2>&1 | grep "K1 fast\|^K3" | tail -n 9 | cut -c1-230 | tee $O/r2b_quick_2048_f64.txt
timeout 120 python tools/quick_bench.py 2048 6 4 512 greens 2>&1 | grep "K1 fast\|^K3" | tail -n 9 | cut -c1-230 | tee $O/r2b_quick_2048_f32.txt
KSN_K3_NOTMA=1 KSN_K1_PAIR=1 timeout 120 python tools/quick_bench.py 2048 4 4 512 2>&1 | grep "K1 fast\|^K3" | tail -n 4 | cut -c1-230 | tee $O/r2b_quick_2048_f32_plain.txt
KSN_K3_EXACT=1 timeout 120 python tools/quick_bench.py 2048 4 8 512 greens 2>&1 | grep "^K3" | tail -n 6 | cut -c1-230 | tee $O/r2b_quick_2048_f64_exact.txt
for n in 256 512 1024; do
  timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "K1 fast\|^K3" | tail -n 2 | sed "s/^/PMGRID $n f64: /" | cut -c1-230
  timeout 120 python tools/quick_bench.py $n 5 4 2>&1 | grep "K1 fast\|^K3" | tail -n 2 | sed "s/^/PMGRID $n f32: /" | cut -c1-230
done | tee $O/r2b_small_grids.txt
timeout 300 python tools/pm4096_probe.py 384 2>&1 | grep -h "^K1\|^K3" | cut -c1-230 | tee $O/r2b_pm4096_probe.txt
timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2b_step_f32.txt
# launch list of the bench command (serialised, cold caches: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_bench_2048.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-greens --no-check > $O/r2b_launches.log 2>&1
echo "launch list exit $?"
# ncu --set full of the shipped kernels (256-plane slabs of 2048^3: 8.6 GB double, the kernels go by width)
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k1_tile_kernel -s 3 -c 1 -o $O/r2b_k1_tile_2048 python tools/quick_bench.py 2048 4 8 256 > $O/r2b_ncu_k1.log 2>&1
timeout 400 $NCU -k regex:k3_scale_row_kernel -s 2 -c 1 -o $O/r2b_k3_row_2048 python tools/quick_bench.py 2048 4 8 256 > $O/r2b_ncu_k3.log 2>&1
timeout 400 $NCU -k regex:k3_scale_row_kernel -s 5 -c 1 -o $O/r2b_k3_row_greens_2048 python tools/quick_bench.py 2048 4 8 256 greens > $O/r2b_ncu_k3g.log 2>&1
timeout 400 $NCU -k regex:k1_tile_kernel -s 3 -c 1 -o $O/r2b_k1_tile_f32_2048 python tools/quick_bench.py 2048 4 4 256 > $O/r2b_ncu_k1f.log 2>&1
timeout 400 $NCU -k regex:k3_scale_flat_kernel -s 2 -c 1 -o $O/r2b_k3_flat_f32_2048 python tools/quick_bench.py 2048 4 4 256 > $O/r2b_ncu_k3f.log 2>&1
timeout 240 $NCU -k regex:k2_delta_nu -s 2 -c 1 -o $O/r2b_k2_spec3 python tools/k2_bench.py 788 1 > $O/r2b_ncu_k2.log 2>&1
timeout 240 $NCU -k regex:fslength_kernel -s 2 -c 1 -o $O/r2b_fslength python tools/k2_bench.py 788 1 > $O/r2b_ncu_fs.log 2>&1
ls -la $O/*.ncu-rep | tail -n 12
