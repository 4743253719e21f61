#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python scripts/r2_ragged.py 2>&1 | grep "f32"
timeout 100 python scripts/quick_perf.py single
