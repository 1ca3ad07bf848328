#!/bin/bash
# Two GPUs: the multi-GPU parity test with both collective backends, then the bench with each (2048^3 and a small grid
# where the collective's latency shows).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/p2p_gpus.txt 2>&1; nvidia-smi topo -m >> $O/p2p_gpus.txt 2>&1
timeout 600 python -m pytest -q -m gpu tests/test_multi_gpu.py -k "2-" > $O/p2p_tests.log 2>&1
echo "multi-gpu tests exit $?" | tee -a $O/p2p_tests.log
tail -n 30 $O/p2p_tests.log
run() {  # backend pmgrid
  KSN_COMM=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens --pmgrid $2 > $O/p2p_bench_$1_$2.log 2>&1
  tail -n 1 $O/p2p_bench_$1_$2.log | cut -c1-2500
}
for g in 512 2048; do for b in nccl p2p; do run $b $g; done; done
