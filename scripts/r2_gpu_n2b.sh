#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_round2.py -m gpu -x -q -k "sharded or gather" > $O/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_n2.log
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-k1 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; grep "^{" $O/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['gather_config4'])"
timeout 600 python bench.py --gpus 2 --config 4 --steps 20 > $O/bench_c4_n2.json 2>> $O/bench_n2.err; grep "^{" $O/bench_c4_n2.json | cut -c1-200
