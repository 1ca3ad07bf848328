#!/bin/bash
# compute-sanitizer over the K2 kernels after the critical-path work (CTA-wide spline solve in shared scratch, redux argmax,
# first pass with the halves):  gpurun --timeout 900 -- 'bash tools/gpu_r2_sanitize2.sh'
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 400 $CS --tool memcheck python -m pytest tests/test_k2_gpu.py -q -k "update_sequence_matches_reference or speculative_bisection" > $O/r2s_memcheck_k2.log 2>&1
timeout 400 $CS --tool racecheck python -m pytest tests/test_k2_gpu.py -q -k "speculative_bisection and masses0" > $O/r2s_racecheck_k2.log 2>&1
timeout 300 $CS --tool synccheck python -m pytest tests/test_k2_gpu.py -q -k "speculative_bisection and masses0" > $O/r2s_synccheck_k2.log 2>&1
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2s_memcheck_k2.log $O/r2s_racecheck_k2.log $O/r2s_synccheck_k2.log | tee $O/r2s_sanitizer.txt
