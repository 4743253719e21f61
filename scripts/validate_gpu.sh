#!/bin/bash
# Round-end validation on one B200 (run through gpurun): GPU tests, smoke, both bench arms, every BASELINE config, the ncu
# launch list of the bench command and full ncu captures of the headline kernels.  Every ncu run follows the same command
# having exited 0 without ncu.  Outputs land in gpurun_out/; scripts/refresh_profiles.sh copies the summaries to profiles/.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2>> $O/bench.err; echo "bench ref rc=$?"
for c in 3 4 5; do timeout 300 python bench.py --config $c --steps 50 --no-cpu-baseline > $O/bench_c$c.json 2>> $O/bench.err; echo "config $c rc=$?"; done
timeout 300 python bench.py --config 5 --precision f64 --steps 50 --no-cpu-baseline > $O/bench_c5_f64.json 2>> $O/bench.err
timeout 300 python bench.py --config 3 --precision f64 --steps 50 --no-cpu-baseline > $O/bench_c3_f64.json 2>> $O/bench.err
timeout 300 python bench.py --precision f64 --muscles 32768 --steps 20 --no-cpu-baseline --no-k1 > $O/bench_c2_f64.json 2>> $O/bench.err
timeout 300 python scripts/sweep_sizes.py > $O/sweeps.txt 2>&1
timeout 300 python scripts/r2_shortwin.py >> $O/sweeps.txt 2>&1
timeout 300 python scripts/r2_wide.py > $O/wide.txt 2>&1
timeout 300 python scripts/r2_ragged.py > $O/ragged.txt 2>&1
timeout 200 python scripts/perf_env_small.py > $O/env_small.txt 2>&1
timeout 200 python scripts/perf_dropin.py > $O/dropin.txt 2>&1
timeout 200 python scripts/pcie_probe.py > $O/pcie_n1.json 2>/dev/null
NCU="ncu --clock-control none"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
cap() {  # name, kernel regex, script args...
  local name=$1 rx=$2; shift 2
  timeout 300 python "$@" > $O/plain.log 2>&1 && \
    timeout 700 $NCU --set full --import-source on -k regex:$rx -s 2 -c 1 -f -o $O/$name python "$@" > $O/ncu_$name.log 2>&1
  echo "$name rc=$?"
  if [ -f $O/$name.ncu-rep ]; then
    python scripts/ncu_summary.py $O/$name.ncu-rep > $O/${name}_summary.txt 2>&1
    rm -f $O/$name.ncu-rep
  fi
}
cap prof_k2w_f32 k2w_f32 scripts/prof_fused.py f32 65536
cap prof_k2w_f64 k2w_f64 scripts/prof_fused.py f64 32768
cap prof_k1 k1_stream scripts/prof_single.py f32
cap prof_k1_f64 k1_stream scripts/prof_single.py f64 65536
cap prof_env k1_flat scripts/prof_env.py f32
cap prof_hopper k1_flat scripts/prof_hopper.py f32
cap prof_wide k2x_f32 scripts/prof_wide.py
cat $O/bench.json
