#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err; cat $O/bench.json
for c in 3 4 5; do timeout 300 python bench.py --config $c --steps 50 --no-cpu-baseline > $O/bench_c$c.json 2>> $O/bench.err; echo "bench config $c rc=$?"; cat $O/bench_c$c.json; done
timeout 300 python bench.py --config 5 --precision f64 --steps 50 --no-cpu-baseline > $O/bench_c5_f64.json 2>> $O/bench.err; cat $O/bench_c5_f64.json
timeout 300 python bench.py --precision f64 --muscles 32768 --steps 20 --no-cpu-baseline --no-k1 > $O/bench_c2_f64.json 2>> $O/bench.err; cat $O/bench_c2_f64.json
timeout 200 python scripts/perf_dropin.py > $O/dropin.txt 2>&1; cat $O/dropin.txt
