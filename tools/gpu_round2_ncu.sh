#!/bin/bash
# ncu --set full captures of the kernels that were opt-in at the end of round 1 (one GPU; each capture replays its kernel
# ~40 times).  Run AFTER tools/gpu_round2_check.sh has shown them correct:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_ncu.sh'
# Reports land in gpurun_out/r2_*.ncu-rep; read them here with tools/ncu_summary.py and copy the summaries to profiles/.
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# float grid, 1024^3 (4.3 GB: the replays stay short): K1 tile kernel on float rows, K3 flat bulk-copy kernel, and the
# default float kernels beside them
KSN_K1_F32_TILE=1 KSN_K3_F32_TMA=1 timeout 300 $NCU -k regex:k1_tile_kernel -s 2 -c 1 -o $O/r2_k1_tile_f32_1024 python tools/quick_bench.py 1024 3 4 > $O/r2_ncu_k1_f32.log 2>&1
KSN_K1_F32_TILE=1 KSN_K3_F32_TMA=1 timeout 300 $NCU -k regex:k3_scale_tma_flat -s 1 -c 1 -o $O/r2_k3_flat_f32_1024 python tools/quick_bench.py 1024 3 4 > $O/r2_ncu_k3_f32.log 2>&1
timeout 300 $NCU -k regex:k1_pair_kernel -s 2 -c 1 -o $O/r2_k1_pair_f32_1024 python tools/quick_bench.py 1024 3 4 > $O/r2_ncu_k1_pair_f32.log 2>&1
timeout 300 $NCU -k regex:k3_scale_kernel -s 1 -c 1 -o $O/r2_k3_plain_f32_1024 python tools/quick_bench.py 1024 3 4 > $O/r2_ncu_k3_plain_f32.log 2>&1
# PMGRID 4096, 96-plane slab (12.9 GB): the bin window with a bin's home chosen per update / per tile
# (the probe launches the tile kernel 4 times to warm up, then 6 times each with KSN_K1_WIN = 0, 1, 3)
timeout 400 $NCU -k regex:k1_tile_kernel -s 10 -c 1 -o $O/r2_k1_win1_4096 python tools/pm4096_probe.py 96 > $O/r2_ncu_k1_win1.log 2>&1
timeout 400 $NCU -k regex:k1_tile_kernel -s 16 -c 1 -o $O/r2_k1_win3_4096 python tools/pm4096_probe.py 96 > $O/r2_ncu_k1_win3.log 2>&1
# K2 at the benchmark's shape: the width-3 one-CTA kernel and the cluster kernel
timeout 240 $NCU -k regex:k2_delta_nu -s 2 -c 1 -o $O/r2_k2_spec3 python tools/k2_bench.py 788 1 > $O/r2_ncu_k2_spec.log 2>&1
KSN_K2_CLUSTER=3 timeout 240 $NCU -k regex:k2_delta_nu_cluster -s 2 -c 1 -o $O/r2_k2_cluster3 python tools/k2_bench.py 788 1 > $O/r2_ncu_k2_cluster.log 2>&1
ls -la $O/*.ncu-rep
