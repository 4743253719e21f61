#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29575 scripts/r2_gather_ab.py > $O/gather_ab_n8.txt 2> $O/gather_ab_n8.err; echo "rc=$?"; grep gather_ab $O/gather_ab_n8.txt; tail -5 $O/gather_ab_n8.err
