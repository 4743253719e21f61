#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log
timeout 200 python scripts/r2_f64ab.py 2>&1 | tee $O/f64ab.txt
timeout 300 python scripts/quick_perf.py single > $O/quick.txt 2>&1; cat $O/quick.txt
timeout 300 python scripts/r2_ragged.py > $O/ragged_auto.txt 2>&1; cat $O/ragged_auto.txt
MB_K1_FLAT=1 timeout 300 python scripts/r2_ragged.py > $O/ragged_flat.txt 2>&1; grep "arm\|1137\|n=256" $O/ragged_flat.txt
MB_K1_FLAT=0 timeout 300 python scripts/r2_ragged.py > $O/ragged_stream.txt 2>&1; grep "arm\|1137\|n=256" $O/ragged_stream.txt
NCU="ncu --clock-control none"
timeout 200 python scripts/prof_fused.py f64 32768 > $O/plain.log 2>&1 && \
  timeout 600 $NCU --set full --import-source on -k regex:k2w_f64 -s 2 -c 1 -f -o $O/prof_k2w_f64 python scripts/prof_fused.py f64 32768 > $O/ncu_k2w_f64.log 2>&1; echo "ncu k2w_f64 rc=$?"
timeout 200 python scripts/prof_fused.py f32 65536 32 > $O/plain.log 2>&1 && \
  timeout 600 $NCU --set full --import-source on -k regex:k2w_f32 -s 2 -c 1 -f -o $O/prof_k2w_k32 python scripts/prof_fused.py f32 65536 32 > $O/ncu_k2w_k32.log 2>&1; echo "ncu k2w K=32 rc=$?"
for r in prof_k2w_f64 prof_k2w_k32; do
  if [ -f $O/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
    ncu -i $O/$r.ncu-rep --page source --csv --print-source sass > $O/${r}_src.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
cat $O/prof_k2w_f64_summary.txt $O/prof_k2w_k32_summary.txt
