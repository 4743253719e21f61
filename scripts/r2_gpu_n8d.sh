#!/bin/bash
set -u
show() { grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['gather_config4'])"; }
echo "== gather_ab"; MB_AB_F64_ONLY=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 scripts/r2_gather_ab.py 2>/dev/null | grep "gather_ab\] f64"
echo "== bench as is"; timeout 300 python bench.py --gpus 8 --steps 3 --no-e2e --no-k1 2>/dev/null | show
echo "== bench fresh batch for the peer leg"; MB_BENCH_FRESH_PEER=1 timeout 300 python bench.py --gpus 8 --steps 3 --no-e2e --no-k1 2>/dev/null | show
