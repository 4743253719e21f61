#!/bin/bash
# A/B of library builds on the single-step shapes (development aid)
for L in libmuscle_b200_prev.so libmuscle_b200.so; do for P in f32 f64; do
  echo "$L: $(MB_LIB=$PWD/pymuscle_b200/csrc/$L timeout 200 python scripts/perf_single_f64.py $P 2>&1 | tail -1)"
done; done
