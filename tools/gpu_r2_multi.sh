#!/bin/bash
# Round 2, multi-GPU calls:  gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_r2_multi.sh N'     (N = 2, 4, 8)
# The BASELINE.json configs that had never run (configs[2]: 1024^3; configs[4]: 4096^3 non-degenerate masses on 4 / 8 GPUs),
# the N-rank parity tests with both collective backends, the slab FFT across GPUs, and what the host links give.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > $O/r2m${N}_gpus.txt; nproc >> $O/r2m${N}_gpus.txt; free -g | head -2 >> $O/r2m${N}_gpus.txt
nvidia-smi topo -m > $O/r2m${N}_topo.txt 2>&1
# parity: slab-sharded steps on N ranks == one rank, NCCL and peer memory (bit-identical across ranks for the latter)
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -k "[$N-" > $O/r2m${N}_multi_gpu_tests.log 2>&1
echo "multi-GPU parity tests (world $N) exit $?" | tee -a $O/r2m${N}_multi_gpu_tests.log
tail -n 5 $O/r2m${N}_multi_gpu_tests.log | cut -c1-250
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_promoted_kernels_gpu.py -q -k mpi_build > $O/r2m${N}_mpi_build.log 2>&1
  echo "MPI-build bootstrap on 2 GPUs exit $?" | tee -a $O/r2m${N}_mpi_build.log
  tail -n 5 $O/r2m${N}_mpi_build.log | cut -c1-250
fi
timeout 600 python -m pytest tests/test_fft_gpu.py -q -k "slab_ranks" > $O/r2m${N}_fft_tests.log 2>&1
echo "slab FFT tests exit $?" | tee -a $O/r2m${N}_fft_tests.log
tail -n 5 $O/r2m${N}_fft_tests.log | cut -c1-250
# configs[2]: PMGRID 1024, 0.3 eV total, long history
timeout 300 $TR --master-port 29611 bench.py --gpus $N --pmgrid 1024 --no-hybrid --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-greens > $O/r2m${N}_bench_1024.log 2>&1
echo "bench 1024 exit $?"; tail -n 1 $O/r2m${N}_bench_1024.log | cut -c1-1200
# configs[3]: the headline
timeout 400 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-greens > $O/r2m${N}_bench_2048.log 2>&1
echo "bench 2048 exit $?"; tail -n 1 $O/r2m${N}_bench_2048.log | cut -c1-3000
# configs[4]: PMGRID 4096 (550 GB: 4 GPUs or more), non-degenerate masses
if [ "$N" -ge 4 ]; then
  timeout 600 $TR --master-port 29613 bench.py --gpus $N --pmgrid 4096 --mnu 0.2,0.1,0.3 --no-hybrid --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-greens > $O/r2m${N}_bench_4096.log 2>&1
  echo "bench 4096 exit $?"; tail -n 1 $O/r2m${N}_bench_4096.log | cut -c1-3000
fi
# the slab FFT in front of the hook, and the PM step that starts from a device-resident real-space density
timeout 400 $TR --master-port 29614 tools/fft_bench.py 2048 5 > $O/r2m${N}_fft_2048.log 2>&1
echo "fft bench 2048 exit $?"; tail -n 1 $O/r2m${N}_fft_2048.log | cut -c1-800
timeout 300 $TR --master-port 29615 tools/fft_bench.py 1024 5 > $O/r2m${N}_fft_1024.log 2>&1
echo "fft bench 1024 exit $?"; tail -n 1 $O/r2m${N}_fft_1024.log | cut -c1-800
# host <-> device links with every rank copying at once
timeout 200 $TR --master-port 29616 tools/pcie_probe.py > $O/r2m${N}_pcie.log 2>&1
echo "pcie probe exit $?"; tail -n 1 $O/r2m${N}_pcie.log | cut -c1-600
