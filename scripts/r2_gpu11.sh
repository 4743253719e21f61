#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python scripts/r2_wide.py 2>&1 | grep -i "hopper\|arm"
MB_FUSED_UNITS=1 timeout 300 python scripts/r2_mid.py
