#!/bin/bash
# final 8-GPU headline line (gather trio timed from rested batches)
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
grep "^{" gpurun_out/bench_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g'%(d['value'], d['e2e']['value']), d['roofline']['gather_config4'])"
