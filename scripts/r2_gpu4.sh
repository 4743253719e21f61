#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 300 python scripts/r2_ragged.py > $O/ragged_flat.txt 2>&1; cat $O/ragged_flat.txt
MB_K1_FLAT=0 timeout 300 python scripts/r2_ragged.py > $O/ragged_stream.txt 2>&1; cat $O/ragged_stream.txt
MB_K1_FLAT_GROUP=2048 timeout 300 python scripts/r2_ragged.py > $O/ragged_flat2048.txt 2>&1; grep arm $O/ragged_flat2048.txt
MB_K1_FLAT_GROUP=8192 timeout 300 python scripts/r2_ragged.py > $O/ragged_flat8192.txt 2>&1; grep arm $O/ragged_flat8192.txt
NCU="ncu --clock-control none"
timeout 200 python scripts/prof_fused.py f64 32768 > $O/plain.log 2>&1 && \
  timeout 600 $NCU --set full --import-source on -k regex:k2w_f64 -s 2 -c 1 -f -o $O/prof_k2w_f64 python scripts/prof_fused.py f64 32768 > $O/ncu_k2w_f64.log 2>&1; echo "ncu k2w_f64 rc=$?"
for r in prof_k2w_f64; do
  if [ -f $O/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
    ncu -i $O/$r.ncu-rep --page source --csv --print-source sass > $O/${r}_src.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
cat $O/prof_k2w_f64_summary.txt
