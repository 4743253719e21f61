#!/bin/bash
# A/B of two builds of the library on the single-step shapes (development aid): MB_LIB picks the build
for L in pymuscle_b200/csrc/libmuscle_b200_prev.so pymuscle_b200/csrc/libmuscle_b200.so; do
  echo "== $L"
  MB_LIB=$PWD/$L timeout 200 python scripts/quick_perf.py single 2>&1 | tail -4
  MB_LIB=$PWD/$L timeout 200 python scripts/perf_env_probe.py f32 2>&1 | grep -v "^$"
done
