#!/bin/bash
# Round-end validation on one B200 (run through gpurun): GPU tests, smoke, both bench arms, the ncu launch
# list of the bench command and full ncu captures of the single-step kernel (configs #5, #3, FP64 #5).
# Every ncu run follows the same command having exited 0 without ncu.  Outputs land in gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2>> $O/bench.err; echo "bench ref rc=$?"
NCU="ncu --clock-control none"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
if [ "${1:-}" != "--no-full" ]; then
  timeout 200 python scripts/prof_single.py f32 > $O/plain.log 2>&1 && \
    timeout 500 $NCU --set full --import-source on -k regex:k1_stream -s 2 -c 1 -f -o $O/prof_k1 python scripts/prof_single.py f32 > $O/ncu_k1.log 2>&1; echo "k1 f32 rc=$?"
  timeout 200 python scripts/prof_env.py f32 > $O/plain.log 2>&1 && \
    timeout 500 $NCU --set full --import-source on -k regex:k1_stream -s 2 -c 1 -f -o $O/prof_env python scripts/prof_env.py f32 > $O/ncu_env.log 2>&1; echo "env f32 rc=$?"
  timeout 200 python scripts/prof_single.py f64 65536 > $O/plain.log 2>&1 && \
    timeout 500 $NCU --set full --import-source on -k regex:k1_stream -s 2 -c 1 -f -o $O/prof_k1_f64 python scripts/prof_single.py f64 65536 > $O/ncu_k1_f64.log 2>&1; echo "k1 f64 rc=$?"
fi
# gpurun brings back at most 64 MiB: summarise the captures here and keep only the text
for r in prof_k1 prof_env prof_k1_f64; do
  if [ -f $O/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
    ncu -i $O/$r.ncu-rep --page source --csv --print-source sass > $O/${r}_src.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
cat $O/bench.json
