#!/bin/bash
# compute-sanitizer over the kernels added or rewritten in round 2 (small cases only: the tools slow kernels down 10-50x):
# K3 row / flat kernels with every factor mode, the float K1 tile kernel, the K2 prefetch path, the FFT exchange kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_sanitize.sh'
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck python -m pytest tests/test_promoted_kernels_gpu.py -q -k "not 2048 and not 4096 and not 1024 and not mpi_build and not 256" > $O/r2_memcheck_promoted.log 2>&1
timeout 600 $CS --tool memcheck python -m pytest tests/test_k3_gpu.py -q -k "not 2048 and not 4096 and not 1024 and not 256" > $O/r2_memcheck_k3.log 2>&1
timeout 600 $CS --tool memcheck python -m pytest tests/test_fft_gpu.py -q -k "one_rank and not 96" > $O/r2_memcheck_fft.log 2>&1
timeout 600 $CS --tool memcheck python -m pytest tests/test_k2_gpu.py -q -k "prefetched and masses0" > $O/r2_memcheck_k2prefetch.log 2>&1
timeout 600 $CS --tool racecheck python -m pytest tests/test_promoted_kernels_gpu.py -q -k "64 and not 2048 and not 4096 and not 1024 and not mpi_build and not 256" > $O/r2_racecheck_promoted.log 2>&1
timeout 600 $CS --tool racecheck python -m pytest tests/test_k3_gpu.py -q -k "gadget2 and 64" > $O/r2_racecheck_k3.log 2>&1
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2_memcheck_*.log $O/r2_racecheck_*.log | tee $O/r2_sanitizer.txt
