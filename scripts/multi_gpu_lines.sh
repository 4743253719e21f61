#!/bin/bash
# 8-GPU run of the bench lines kept under profiles/ (builder-run; the driver runs its own SCALE at round end)
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --gpus 8 --steps 30 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
timeout 600 python bench.py --gpus 4 --steps 30 --warmup 3 > $O/bench_n4.json 2>> $O/bench_n8.err; echo "bench n4 rc=$?"
timeout 600 python bench.py --gpus 2 --steps 30 --warmup 3 > $O/bench_n2.json 2>> $O/bench_n8.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --gpus 8 --config 4 --steps 20 > $O/bench_c4_n8.json 2>> $O/bench_n8.err; echo "bench c4 n8 rc=$?"
timeout 600 python bench.py --gpus 8 --config 3 --steps 50 > $O/bench_c3_n8.json 2>> $O/bench_n8.err
timeout 600 python bench.py --gpus 8 --config 5 --steps 50 > $O/bench_c5_n8.json 2>> $O/bench_n8.err
timeout 600 python bench.py --gpus 8 --strong --muscles 524288 --steps 20 --no-e2e-variants --no-gather > $O/bench_strong_n8.json 2>> $O/bench_n8.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_n8.log 2>&1; tail -3 $O/pytest_n8.log
for f in bench_n8 bench_n4 bench_n2 bench_c4_n8 bench_c3_n8 bench_c5_n8 bench_strong_n8; do grep "^{" $O/$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
e = d.get('e2e') or {}
print('$f', 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %s' % ('%.4g' % e['value'] if e else None), d['roofline'].get('gather_config4'))
"; done
tail -3 $O/bench_n8.err
