#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
MB_FUSED_PERSIST=1 timeout 300 python scripts/r2_shortwin.py 2>&1 | grep potvin
MB_FUSED_PERSIST=0 timeout 300 python scripts/r2_shortwin.py 2>&1 | grep "potvin   K= 1000\|potvin   K=   32"
