#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 scripts/r2_gather_diag.py 2>gpurun_out/gather_diag.err | grep "diag\]" | tee gpurun_out/gather_diag.txt
