#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for ns in 2 3; do for bps in 0; do echo "== MB_K1_STAGES=$ns"; MB_K1_STAGES=$ns timeout 300 python scripts/r2_ragged.py 2>&1 | grep "f32" ; done; done
echo "== config 5 single"; timeout 100 python scripts/quick_perf.py single
MB_K1_STAGES=3 timeout 100 python scripts/quick_perf.py single
