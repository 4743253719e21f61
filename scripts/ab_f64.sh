#!/bin/bash
# A/B of library builds on the single-step shapes (development aid)
for L in libmuscle_b200.so libmuscle_b200_f2g2.so libmuscle_b200_f2g4.so libmuscle_b200_f2g8.so; do
  echo "$L: $(MB_LIB=$PWD/pymuscle_b200/csrc/$L timeout 200 python scripts/perf_single_f64.py f64 2>&1 | tail -1)"
done
