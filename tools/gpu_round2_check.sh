#!/bin/bash
# One gpurun call for the start of the next round: everything written after round 1's GPU minutes were spent.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_check.sh'      (then tools/gpu_round2_sanitize.sh in a call of its own)
# and, on two GPUs, the collective bootstrap of the -DKSN_HAVE_MPI build plus the e2e leg with the peer-memory backend:
#   gpurun --gpus 2 --timeout 600 -- 'KSN_TEST_UNVERIFIED=1 python -m pytest tests/test_zz_optin_gpu.py -q -k mpi_build; python -m pytest tests/test_multi_gpu.py -q -m gpu; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 | tail -n 1 > gpurun_out/r2_bench_2gpu_e2e.json'
mkdir -p gpurun_out
O=gpurun_out
# 1. the reference's own cmocka programs against the product on the GPU (tests/test_zz_reference_programs_gpu.py)
timeout 200 python -m pytest tests/test_zz_reference_programs_gpu.py -q > $O/r2_reference_programs.log 2>&1
echo "reference programs exit $?" | tee -a $O/r2_reference_programs.log
# 2. K1 table cache (opt-in): whole GPU suite with it on, and the bench line with/without
KSN_K1_TABLE_CACHE=1 timeout 200 python -m pytest tests -q -m gpu -x > $O/r2_pytest_tabcache.log 2>&1
echo "suite (table cache) exit $?" | tee -a $O/r2_pytest_tabcache.log
timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_nocache.log 2>&1
KSN_K1_TABLE_CACHE=1 timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2_bench_tabcache.log 2>&1
tail -n 1 $O/r2_bench_nocache.log | cut -c1-400
tail -n 1 $O/r2_bench_tabcache.log | cut -c1-400
# 3. the whole suite as the driver runs it
timeout 200 python -m pytest tests -q -m gpu -x > $O/r2_pytest_gpu.log 2>&1
echo "suite exit $?" | tee -a $O/r2_pytest_gpu.log
tail -n 4 $O/r2_pytest_gpu.log
# 3b. float-grid kernels written after round 1's GPU minutes were spent (opt-in): parity tests, then timings at 2048^3
#     (float grid = 34.4 GB): default float path, K3 through bulk copies, K1 through the tile kernel
KSN_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_zz_optin_gpu.py -q > $O/r2_optin.log 2>&1
echo "opt-in kernels exit $?" | tee -a $O/r2_optin.log
tail -n 3 $O/r2_optin.log
timeout 120 python tools/quick_bench.py 2048 5 4 > $O/r2_f32_default.log 2>&1
KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 120 python tools/quick_bench.py 2048 5 4 > $O/r2_f32_optin.log 2>&1
grep -h "K1 fast\|K3" $O/r2_f32_default.log | tail -n 2
grep -h "K1 fast\|K3" $O/r2_f32_optin.log | tail -n 2
# tile shapes for float rows at 2048^3: the model's choice (15 warps x 9), the double grid's shape, a whole row as one tile
for shape in "8,17,2" "8,33,2" "8,33,3" "12,9,3"; do
  KSN_K1_TILE=$shape KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 120 python tools/quick_bench.py 2048 4 4 2>&1 | grep "K1 fast" | tail -n 1 | cut -c1-200
done
# 3c. K1 bin window at PMGRID 4096 (384-plane slab, 51.6 GB): home of a bin chosen per update (default) / per tile (opt-in)
timeout 300 python tools/pm4096_probe.py 384 > $O/r2_pm4096_probe.log 2>&1
grep -h "^K1" $O/r2_pm4096_probe.log | cut -c1-200
# 3d. K2 with one k bin per cluster of M CTAs (opt-in): K2 phase per cluster size next to the one-CTA kernels
: > $O/r2_k2_cluster.txt
for cl in 0 2 3 4; do
  for cfg in "788 1" "788 1 nondegenerate" "788 0"; do
    echo "--- KSN_K2_CLUSTER=$cl  k2_bench.py $cfg" >> $O/r2_k2_cluster.txt
    KSN_K2_CLUSTER=$cl timeout 120 python tools/k2_bench.py $cfg 2>&1 | tail -n 3 >> $O/r2_k2_cluster.txt
  done
done
cat $O/r2_k2_cluster.txt
# 3e. K1 / K3 alone at the smaller BASELINE grids (configs[1], [2]): GB/s per kernel, double grids
for n in 256 512 1024; do
  timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "K1 fast\|^K3" | tail -n 2 | sed "s/^/PMGRID $n: /" | cut -c1-200
done | tee $O/r2_small_grids.txt
# 3f. K3 with several short rows per CTA (PMGRID <= 1150): the default kernel against the flat-chunk one (opt-in)
for n in 512 1024; do
  KSN_K3_FLAT=1 timeout 120 python tools/quick_bench.py $n 5 8 2>&1 | grep "^K3" | tail -n 1 | sed "s/^/PMGRID $n KSN_K3_FLAT=1: /" | cut -c1-200
done | tee -a $O/r2_small_grids.txt
# 3g. the whole step on a float grid (Gadget-2's default build) at 2048^3: default float kernels / the opt-in bulk-copy ones
timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_default.txt
KSN_K3_F32_TMA=1 KSN_K1_F32_TILE=1 timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_optin.txt
KSN_K3_F32_TMA=2 KSN_K1_F32_TILE=1 timeout 200 python tools/step_bench.py 2048 4 10 2>&1 | tail -n 1 | cut -c1-400 | tee $O/r2_step_f32_optin_short.txt
