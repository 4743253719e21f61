#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
MB_FUSED_PERSIST=1 timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_persist.log 2>&1; echo "pytest(persist=1) rc=$?"; tail -12 $O/pytest_persist.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
MB_FUSED_PERSIST=0 timeout 300 python scripts/r2_shortwin.py 2>&1 | tee $O/shortwin0.txt
MB_FUSED_PERSIST=1 timeout 300 python scripts/r2_shortwin.py 2>&1 | tee $O/shortwin1.txt
