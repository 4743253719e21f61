#!/bin/bash
# Copy the outputs of scripts/validate_gpu.sh from gpurun_out/ into profiles/ (run here, after gpurun returns).
set -e
O=gpurun_out; P=profiles; R=r2
cp $O/bench.json $P/${R}_bench_n1.json
cp $O/bench_ref.json $P/${R}_bench_reference_n1.json
for c in c3 c4 c5 c5_f64 c3_f64 c2_f64; do [ -f $O/bench_$c.json ] && cp $O/bench_$c.json $P/${R}_bench_config_${c#c}.json; done
cp $O/launches.csv $P/${R}_launches_bench.csv
cp $O/prof_k2w_f32_summary.txt $P/${R}_k2w_f32_config2_ncu_full_summary.txt
cp $O/prof_k2w_f64_summary.txt $P/${R}_k2w_f64_config2_ncu_full_summary.txt
cp $O/prof_k1_summary.txt $P/${R}_k1_stream_f32_config5_ncu_full_summary.txt
cp $O/prof_k1_f64_summary.txt $P/${R}_k1_stream_f64_config5_ncu_full_summary.txt
cp $O/prof_env_summary.txt $P/${R}_k1_flat_f32_config3_env_ncu_full_summary.txt
cp $O/prof_hopper_summary.txt $P/${R}_k1_flat_f32_hopper_env_ncu_full_summary.txt
rm -f $P/${R}_k1_stream_f32_config3_env_ncu_full_summary.txt
cp $O/prof_wide_summary.txt $P/${R}_k2x_f32_n2000_ncu_full_summary.txt
cat $O/sweeps.txt > $P/${R}_sweeps_one_gpu.txt
{ echo "== scripts/r2_wide.py =="; cat $O/wide.txt; echo; echo "== scripts/r2_ragged.py =="; cat $O/ragged.txt; echo; echo "== scripts/perf_env_small.py =="; cat $O/env_small.txt; echo; echo "== scripts/perf_dropin.py =="; cat $O/dropin.txt; } > $P/${R}_kernels_one_gpu.txt
python scripts/launch_list_summary.py $P/${R}_launches_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 3 --warmup 3 --no-cpu-baseline" > $P/${R}_launches_bench_summary.txt
python scripts/sass_counts.py > /dev/null
python - <<'PY'
import json, re
def grab(f):
    s = open(f).read()
    def val(name, unit_scale=None):
        m = re.search(r'^' + re.escape(name) + r',([^,]*),([-\d.eE+]+)', s, re.M)
        return (float(m.group(2)), m.group(1)) if m else (None, None)
    r, ru = val('dram__bytes_read.sum'); w, wu = val('dram__bytes_write.sum')
    sc = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    traffic = int(r * sc[ru] + w * sc[wu]) if r is not None else None
    pipes = {}
    for key, name in (("duration", "gpu__time_duration.sum"), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                      ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                      ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                      ("pipe_fp64_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                      ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("registers", "launch__registers_per_thread")):
        v, u = val(name)
        if v is not None:
            pipes[key] = v if key != "duration" else f"{v} {u}"
    pipes["source"] = f
    return traffic, pipes
traffic, pipes = {}, {}
for key, f in (("k2w_f32_config2", "profiles/r2_k2w_f32_config2_ncu_full_summary.txt"),
               ("k2w_f64_config2", "profiles/r2_k2w_f64_config2_ncu_full_summary.txt"),
               ("k2w_f64_config4", "profiles/r2_k2w_f64_config2_ncu_full_summary.txt"),
               ("k1_stream_f32_config5", "profiles/r2_k1_stream_f32_config5_ncu_full_summary.txt"),
               ("k1_stream_f64_config5", "profiles/r2_k1_stream_f64_config5_ncu_full_summary.txt"),
               ("k1_flat_f32_config3_env", "profiles/r2_k1_flat_f32_config3_env_ncu_full_summary.txt"),
               ("k1_flat_f32_hopper_env", "profiles/r2_k1_flat_f32_hopper_env_ncu_full_summary.txt"),
               ("k2x_f32_n2000", "profiles/r2_k2x_f32_n2000_ncu_full_summary.txt")):
    try:
        traffic[key], pipes[key] = grab(f)
    except Exception as e:
        print("skip", key, e)
json.dump(traffic, open('profiles/traffic.json', 'w'), indent=1)
json.dump(pipes, open('profiles/pipes.json', 'w'), indent=1)
b = json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g K2 frac %.3f K1 frac %.3f (%.0f GB/s) cpu %.3g' % (
    b['value'], b['e2e']['value'], b['roofline']['frac'], b['roofline']['k1']['frac'], b['roofline']['k1']['achieved'], b['cpu_baseline']['value']))
PY
