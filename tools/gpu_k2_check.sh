#!/bin/bash
# K2 only: parity tests, then the K2 phase per speculation width (hybrid one species / three species / no hybrid cut).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest -q -m gpu tests/test_k2_gpu.py tests/test_step_gpu.py > $O/k2_tests.log 2>&1
echo "k2 tests exit $?" | tee -a $O/k2_tests.log
tail -4 $O/k2_tests.log
: > $O/r1_k2_widths.txt
for m in 1 2 3 4 0; do
  for cfg in "788 1" "788 1 nondegenerate" "788 0"; do
    echo "--- KSN_K2_SPEC=$m  k2_bench.py $cfg" >> $O/r1_k2_widths.txt
    KSN_K2_SPEC=$m timeout 120 python tools/k2_bench.py $cfg 2>&1 | tail -n 3 >> $O/r1_k2_widths.txt
  done
done
cat $O/r1_k2_widths.txt
