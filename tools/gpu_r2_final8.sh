#!/bin/bash
# Round 2, last 8-GPU call:  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r2_final8.sh'
# The state at the end of the round (K2 prefetch beside K1, direct DMA of the history, pipelined slab FFT) on 8 GPUs.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu -k "[8-" > $O/r2g_multi_gpu_tests.log 2>&1
echo "8-rank parity exit $?"; tail -n 3 $O/r2g_multi_gpu_tests.log | cut -c1-200
timeout 400 $TR --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-greens > $O/r2g_bench_2048.log 2>&1
echo "bench 2048 exit $?"; tail -n 1 $O/r2g_bench_2048.log | cut -c1-2500
timeout 600 $TR --master-port 29613 bench.py --gpus 8 --pmgrid 4096 --mnu 0.2,0.1,0.3 --no-hybrid --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2g_bench_4096.log 2>&1
echo "bench 4096 exit $?"; tail -n 1 $O/r2g_bench_4096.log | cut -c1-1500
timeout 300 $TR --master-port 29611 bench.py --gpus 8 --pmgrid 1024 --no-hybrid --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-greens > $O/r2g_bench_1024.log 2>&1
echo "bench 1024 exit $?"; tail -n 1 $O/r2g_bench_1024.log | cut -c1-1200
timeout 400 $TR --master-port 29614 tools/fft_bench.py 2048 5 > $O/r2g_fft_2048.log 2>&1
echo "fft bench 2048 exit $?"; tail -n 1 $O/r2g_fft_2048.log | cut -c1-900
