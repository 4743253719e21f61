#!/bin/bash
# 8-GPU lines of the final code: headline (with the gather trio) and config 4 with the fused gather (gpurun --gpus 8)
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
grep "^{" gpurun_out/bench_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g'%(d['value'], d['e2e']['value']), d['roofline']['gather_config4'])"
timeout 120 python bench.py --gpus 8 --config 4 > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8.err
grep "^{" gpurun_out/bench_c4_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 value %.4g ms %.3f'%(d['value'], d['ms_per_step']))"
