#!/bin/bash
# round 2, GPU call 1: the whole GPU suite after the translation-unit split + first probes
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
MB_K1_SMALL=0 timeout 300 python scripts/r2_small.py > $O/small_stream.txt 2>&1; echo "small(stream) rc=$?"
MB_K1_SMALL=1000000000 timeout 300 python scripts/r2_small.py > $O/small_lowlat.txt 2>&1; echo "small(lowlat) rc=$?"
timeout 300 python scripts/quick_perf.py fused single > $O/quick.txt 2>&1; echo "quick rc=$?"; cat $O/quick.txt
timeout 300 python scripts/sweep_sizes.py > $O/sweeps.txt 2>&1; echo "sweeps rc=$?"
timeout 200 python scripts/perf_env.py > $O/env.txt 2>&1; cat $O/env.txt
timeout 200 python scripts/perf_env_small.py > $O/env_small.txt 2>&1; cat $O/env_small.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
