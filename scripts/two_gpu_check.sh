#!/bin/bash
set -u
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-k1 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g'%(d['value'], d['e2e']['value']), d['roofline']['gather_config4'])"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
