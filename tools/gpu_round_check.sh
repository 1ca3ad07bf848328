#!/bin/bash
# One gpurun call: the new opt-in kernels (K2 speculative bisection, K1 bin window, K3 row pieces) -- parity tests, timings,
# then the bench with the fastest K2 width, the PMGRID-4096 probe and the whole GPU suite.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
echo "== new tests" ; date
timeout 900 python -m pytest -q -m gpu \
  tests/test_k2_gpu.py::test_speculative_bisection_is_bit_identical \
  tests/test_k3_gpu.py::test_long_rows_split_between_ctas_at_pmgrid_4096 \
  tests/test_k1_gpu.py::test_bin_window_kernel_agrees_small \
  tests/test_k1_gpu.py::test_bin_window_kernel_agrees_at_pmgrid_4096 > $O/new_tests.log 2>&1
echo "new tests exit $?" | tee -a $O/new_tests.log
tail -5 $O/new_tests.log
echo "== K2 widths" ; date
for m in 1 2 3 4; do
  KSN_K2_SPEC=$m timeout 120 python tools/k2_bench.py 788 1 > $O/k2_spec_$m.log 2>&1
  KSN_K2_SPEC=$m timeout 120 python tools/k2_bench.py 788 1 nondegenerate > $O/k2_spec_${m}_3masses.log 2>&1
  KSN_K2_SPEC=$m timeout 120 python tools/k2_bench.py 788 0 > $O/k2_spec_${m}_nohybrid.log 2>&1
done
KSN_K2_SPEC=4 KSN_K2_SPEC4_ONE_PER_SM=1 timeout 120 python tools/k2_bench.py 788 1 > $O/k2_spec_4one.log 2>&1
tail -n 2 $O/k2_spec_*.log
BEST=$(python - <<'PY'
import re, glob
best, bt = 1, 1e9
for m in (1, 2, 3, 4):
    ts = [float(x) for x in re.findall(r"K2 phase: ([0-9.]+) ms", open(f"gpurun_out/k2_spec_{m}.log").read())]
    if len(ts) >= 4:
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        if t < bt: best, bt = m, t
print(best)
PY
)
echo "best K2 width: $BEST" | tee $O/k2_best.txt
echo "== bench (device-resident), sequential K2 then the best width" ; date
KSN_K2_SPEC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/bench_k2seq.log 2>&1
KSN_K2_SPEC=$BEST timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/bench_k2best.log 2>&1
tail -n 1 $O/bench_k2seq.log | cut -c1-400 ; tail -n 1 $O/bench_k2best.log | cut -c1-400
echo "== ncu: K2 kernels" ; date
KSN_K2_SPEC=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:k2_delta_nu -s 2 -c 1 -f -o $O/r1_k2_seq python tools/k2_bench.py 788 1 > $O/ncu_k2_seq.log 2>&1
KSN_K2_SPEC=$BEST timeout 240 ncu --set full --clock-control none --import-source on -k regex:k2_delta_nu -s 2 -c 1 -f -o $O/r1_k2_spec python tools/k2_bench.py 788 1 > $O/ncu_k2_spec.log 2>&1
ls -la $O/*.ncu-rep
echo "== PMGRID 4096 probe" ; date
timeout 300 python tools/pm4096_probe.py 384 > $O/pm4096_probe.log 2>&1
cat $O/pm4096_probe.log
echo "== whole GPU suite" ; date
timeout 1200 python -m pytest tests -q -m gpu --durations=12 > $O/pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/pytest_gpu.log
tail -n 25 $O/pytest_gpu.log
date
