#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
timeout 200 python scripts/prof_hopper.py f32 > $O/plain.log 2>&1 && \
  timeout 600 $NCU --set full --import-source on -k regex:k1_flat -s 2 -c 1 -f -o $O/prof_hopper python scripts/prof_hopper.py f32 > $O/ncu_hopper.log 2>&1; echo "ncu hopper rc=$?"
python scripts/ncu_summary.py $O/prof_hopper.ncu-rep > $O/prof_hopper_summary.txt 2>&1
ncu -i $O/prof_hopper.ncu-rep --page source --csv --print-source sass > $O/prof_hopper_src.csv 2>/dev/null
rm -f $O/prof_hopper.ncu-rep
grep -v "launch__\|lts__\|l1tex__t" $O/prof_hopper_summary.txt
