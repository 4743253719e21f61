#!/bin/bash
# A/B of library builds on the single-step shapes, both precision modes (development aid)
for L in libmuscle_b200_prev.so libmuscle_b200.so libmuscle_b200_fc3.so; do
  for P in f32 f64; do echo "$L: $(MB_LIB=$PWD/pymuscle_b200/csrc/$L timeout 200 python scripts/perf_single_f64.py $P 2>&1 | tail -1)"; done
done
for NS in 2 4; do echo "new NS=$NS: $(MB_K1_STAGES=$NS timeout 200 python scripts/perf_single_f64.py f64 2>&1 | tail -1)"; done
echo "fc3 NS=3: $(MB_LIB=$PWD/pymuscle_b200/csrc/libmuscle_b200_fc3.so MB_K1_STAGES=3 timeout 200 python scripts/perf_single_f64.py f32 2>&1 | tail -1)"
