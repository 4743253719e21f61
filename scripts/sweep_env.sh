#!/bin/bash
# Config #3 (arm env step) tuning sweep of the streaming kernel (development aid)
for W in 256 128; do for NS in 2 3 4; do
  echo "W=$W NS=$NS: $(MB_K1_W=$W MB_K1_STAGES=$NS timeout 200 python scripts/perf_env.py 2>&1 | grep 'arm x' | awk '{print $5, $(NF-8), $(NF-7), $(NF-5), $(NF-4)}' | tr '\n' '|')"
done; done
