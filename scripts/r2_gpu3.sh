#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 300 python scripts/r2_wide.py > $O/wide.txt 2>&1; echo "wide rc=$?"; cat $O/wide.txt
