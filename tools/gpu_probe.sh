#!/bin/bash
# what does the GPU box look like?
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
nvidia-smi topo -m 2>/dev/null | head -12
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)"
cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/cpu.max 2>/dev/null
df -h /dev/shm | tail -1
which gcc nvcc ncu; ls /root/reference 2>&1 | head -2
} > gpurun_out/probe.txt 2>&1
