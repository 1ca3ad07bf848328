#!/bin/bash
# Round 2, one GPU:  gpurun --timeout 900 -- 'bash tools/gpu_r2_call4.sh'
# PMGRID 4096 kernels (BASELINE configs[4]): the slab rank 0 of 8 owns (planes 0..511) under sustained load, and
# ncu --set full of the two 4096-only kernel variants (K1 with the bin window, K3 with rows cut into pieces).
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/pm4096_probe.py 512 2>&1 | grep -h "^K" | cut -c1-260 | tee $O/r2k_pm4096_probe_512.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k1_tile_kernel -s 3 -c 1 -o $O/r2k_k1_tile_win_4096 python tools/quick_bench.py 4096 4 8 64 > $O/r2k_ncu_k1.log 2>&1
timeout 400 $NCU -k regex:k3_scale_row_kernel -s 2 -c 1 -o $O/r2k_k3_row_split_4096 python tools/quick_bench.py 4096 4 8 64 > $O/r2k_ncu_k3.log 2>&1
grep -h "K1 fast\|^K3" $O/r2k_ncu_k1.log | tail -n 3 | cut -c1-200
ls -la $O/r2k*.ncu-rep
