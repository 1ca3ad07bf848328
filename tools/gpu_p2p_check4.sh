#!/bin/bash
# Four GPUs: the multi-GPU parity test with the peer-memory backend, then the 2048^3 bench with each backend.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest -q -m gpu tests/test_multi_gpu.py -k "4-p2p or 2-p2p" > $O/p2p_tests4.log 2>&1
echo "multi-gpu tests exit $?" | tee -a $O/p2p_tests4.log
tail -n 4 $O/p2p_tests4.log
run() {  # backend pmgrid
  KSN_COMM=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-greens --pmgrid $2 > $O/p2p4_bench_$1_$2.log 2>&1
  tail -n 1 $O/p2p4_bench_$1_$2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['collective'], d['config']['pmgrid'], 'ms_per_step', round(d['ms_per_step'], 4), 'k2', round(d['roofline']['k2_ms_per_step'], 4), 'comm', round(d['roofline']['comm_ms_per_step'], 4), 'value', d['value'])
"
}
for b in p2p nccl p2p nccl; do run $b 2048; done
