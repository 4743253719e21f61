#!/bin/bash
# Copy the outputs of scripts/validate_gpu.sh from gpurun_out/ into profiles/ (run here, after gpurun returns).
set -e
O=gpurun_out; P=profiles
cp $O/bench.json $P/r1_bench_n1.json
cp $O/bench_ref.json $P/r1_bench_reference_n1.json
cp $O/launches.csv $P/r1_launches_bench.csv
cp $O/prof_k1_summary.txt $P/r1_k1_stream_f32_config5_ncu_full_summary.txt
cp $O/prof_env_summary.txt $P/r1_k1_stream_f32_config3_env_ncu_full_summary.txt
cp $O/prof_k1_f64_summary.txt $P/r1_k1_stream_f64_n2000_ncu_full_summary.txt
python scripts/launch_list_summary.py $P/r1_launches_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 3 --warmup 3 --no-cpu-baseline" > $P/r1_launches_bench_summary.txt
python - <<'PY'
import json, re
t = json.load(open('profiles/traffic.json'))
def traffic(f):
    s = open(f).read()
    r = float(re.search(r'dram__bytes_read.sum,Gbyte,([\d.]+)', s).group(1))
    w = float(re.search(r'dram__bytes_write.sum,Gbyte,([\d.]+)', s).group(1))
    return int((r + w) * 1e9)
t['k1_stream_f32_config5'] = traffic('gpurun_out/prof_k1_summary.txt')
t['k1_stream_f32_config3_env'] = traffic('gpurun_out/prof_env_summary.txt')
t['k1_stream_f64_n2000x65536'] = traffic('gpurun_out/prof_k1_f64_summary.txt')
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
b = json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g K2 frac %.3f K1 frac %.3f (%.0f GB/s) cpu %.3g ref %.3g' % (
    b['value'], b['e2e']['value'], b['roofline']['frac'], b['roofline_single_step']['frac'],
    b['roofline_single_step']['achieved'], b['cpu_baseline']['value'], json.load(open('gpurun_out/bench_ref.json'))['value']))
PY
