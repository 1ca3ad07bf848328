#!/bin/bash
# compute-sanitizer over the kernels added after profiles/r1_sanitizer.txt was taken (small cases only: the tools slow
# kernels down 10-50x).   gpurun --timeout 1500 -- 'bash tools/gpu_round2_sanitize.sh'
mkdir -p gpurun_out
O=gpurun_out
# K2 with bisections ahead of time, K1 bin window, K3 row pieces; then the opt-in kernels of tests/test_zz_optin_gpu.py
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $CS --tool memcheck python -m pytest tests/test_k2_gpu.py -q -k "speculative" > $O/r2_memcheck_k2spec.log 2>&1
timeout 600 $CS --tool memcheck python -m pytest tests/test_k1_gpu.py -q -k "bin_window and 256" > $O/r2_memcheck_k1win.log 2>&1
timeout 600 $CS --tool racecheck python -m pytest tests/test_k2_gpu.py -q -k "speculative and not True" > $O/r2_racecheck_k2spec.log 2>&1
timeout 600 $CS --tool racecheck python -m pytest tests/test_k1_gpu.py -q -k "bin_window and 256-100" > $O/r2_racecheck_k1win.log 2>&1
KSN_TEST_UNVERIFIED=1 timeout 600 $CS --tool memcheck python -m pytest tests/test_zz_optin_gpu.py -q -k "not 2048 and not 4096 and not k2 and not 512" > $O/r2_memcheck_optin.log 2>&1
KSN_TEST_UNVERIFIED=1 timeout 600 $CS --tool racecheck python -m pytest tests/test_zz_optin_gpu.py -q -k "64 and not 2048 and not 4096 and not k2" > $O/r2_racecheck_optin.log 2>&1
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2_memcheck_*.log $O/r2_racecheck_*.log
