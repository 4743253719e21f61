#!/bin/bash
# K1 streaming kernel tuning sweep (development aid)
for NS in 2 3 4 5 6 8; do for BPS in 0 1 2; do
  echo "NS=$NS BPS=$BPS: $(MB_K1_STAGES=$NS MB_K1_BPS=$BPS timeout 120 python scripts/quick_perf.py single 2>&1 | awk '{printf "%s %s %s %s | ", $2,$3,$4,$8}')"
done; done
