#!/bin/bash
# Round 2, one GPU, after the K2 critical-path work:  gpurun --timeout 1500 -- 'bash tools/gpu_r2_call5.sh'
# Whole GPU suite, smoke(), the bench with the driver's flags, K2 alone in its three regimes, ncu --set full of the K2 kernel.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r2p_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2p_pytest_gpu.log
tail -n 4 $O/r2p_pytest_gpu.log | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2p_bench_n1.json 2> $O/r2p_bench_n1.err
echo "bench exit $?"; tail -n 1 $O/r2p_bench_n1.json | cut -c1-1500
for args in "788 1" "788 1 nondegenerate" "788 0"; do echo "--- k2_bench.py $args"; timeout 120 python tools/k2_bench.py $args 2>&1 | grep "K2 phase" | tail -n 3 | cut -c1-200; done | tee $O/r2p_k2_phase.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 240 $NCU -k regex:k2_delta_nu -s 2 -c 1 -o $O/r2p_k2_spec3 python tools/k2_bench.py 788 1 > $O/r2p_ncu_k2.log 2>&1
ls -la $O/r2p*.ncu-rep
