#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/p2p_diag.py 2048 > $O/p2p_diag_2048.log 2>&1
grep "back to back" $O/p2p_diag_2048.log | sort
tail -n 5 $O/p2p_diag_2048.log
