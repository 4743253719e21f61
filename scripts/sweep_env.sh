#!/bin/bash
# Config #3 (arm env step) tuning sweep of the streaming kernel (development aid)
for W in 256 192 160 128; do for NS in 2 3; do
  echo "W=$W NS=$NS: $(MB_K1_W=$W MB_K1_STAGES=$NS timeout 200 python scripts/perf_single_f64.py f32 2>&1 | tail -1)"
done; done
