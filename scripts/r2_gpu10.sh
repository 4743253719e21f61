#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 100 python scripts/quick_perf.py single
timeout 300 python scripts/r2_mid.py
MB_FUSED_WIDE=1 timeout 300 python scripts/r2_mid.py
