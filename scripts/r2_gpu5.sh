#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log
timeout 200 python scripts/r2_f64ab.py 2>&1 | tee $O/f64ab.txt
for v in s0u1 s1u1 s1u2 s2u2; do MB_LIB=$PWD/pymuscle_b200/csrc/variants/libmb_$v.so timeout 200 python scripts/r2_f64ab.py 2>&1 | tee -a $O/f64ab.txt; done
timeout 300 python scripts/r2_ragged.py > $O/ragged_auto.txt 2>&1; cat $O/ragged_auto.txt
MB_K1_FLAT=1 timeout 300 python scripts/r2_ragged.py > $O/ragged_flat.txt 2>&1; grep "arm\|1137\|256" $O/ragged_flat.txt
NCU="ncu --clock-control none"
timeout 200 python scripts/prof_fused.py f64 32768 > $O/plain.log 2>&1 && \
  timeout 600 $NCU --set full --import-source on -k regex:k2w_f64 -s 2 -c 1 -f -o $O/prof_k2w_f64 python scripts/prof_fused.py f64 32768 > $O/ncu_k2w_f64.log 2>&1; echo "ncu k2w_f64 rc=$?"
MB_K1_FLAT=1 timeout 200 python scripts/prof_env.py f32 > $O/plain.log 2>&1 && \
  MB_K1_FLAT=1 timeout 600 $NCU --set full --import-source on -k regex:k1_flat -s 2 -c 1 -f -o $O/prof_flat_arm python scripts/prof_env.py f32 > $O/ncu_flat.log 2>&1; echo "ncu flat rc=$?"
for r in prof_k2w_f64 prof_flat_arm; do
  if [ -f $O/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
    ncu -i $O/$r.ncu-rep --page source --csv --print-source sass > $O/${r}_src.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
cat $O/prof_k2w_f64_summary.txt $O/prof_flat_arm_summary.txt
