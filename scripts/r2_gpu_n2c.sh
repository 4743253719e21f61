#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29575 scripts/r2_gather_ab.py 2>/dev/null | grep gather_ab
