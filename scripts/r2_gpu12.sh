#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
for p in 0 1; do
  MB_FUSED_PERSIST=$p timeout 200 python scripts/prof_fused.py f32 65536 256 > $O/plain.log 2>&1 && \
  MB_FUSED_PERSIST=$p timeout 600 $NCU --set full --import-source on -k regex:k2w_f32 -s 2 -c 1 -f -o $O/prof_persist$p python scripts/prof_fused.py f32 65536 256 > $O/ncu_persist$p.log 2>&1; echo "ncu persist=$p rc=$?"
  python scripts/ncu_summary.py $O/prof_persist$p.ncu-rep > $O/prof_persist${p}_summary.txt 2>&1
  ncu -i $O/prof_persist$p.ncu-rep --page source --csv --print-source sass > $O/prof_persist${p}_src.csv 2>/dev/null
  rm -f $O/prof_persist$p.ncu-rep
  grep -v "stalled_\|lts__\|l1tex__t" $O/prof_persist${p}_summary.txt
done
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-k1 --no-e2e-variants | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('value %.4g e2e %.4g'%(d['value'], d['e2e']['value']))"
