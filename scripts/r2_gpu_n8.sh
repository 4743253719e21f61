#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
timeout 600 python bench.py --gpus 8 --steps 30 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; grep "^{" $O/bench_n8.json | cut -c1-300
timeout 600 python bench.py --gpus 8 --config 4 --steps 20 > $O/bench_c4_n8.json 2>> $O/bench_n8.err; echo "bench c4 n8 rc=$?"; grep "^{" $O/bench_c4_n8.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29570 scripts/pcie_probe.py > $O/pcie_n8.json 2>> $O/bench_n8.err; grep "^{" $O/pcie_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 scripts/pcie_probe.py > $O/pcie_n4.json 2>> $O/bench_n8.err; grep "^{" $O/pcie_n4.json
timeout 600 python bench.py --gpus 8 --config 3 --steps 50 > $O/bench_c3_n8.json 2>> $O/bench_n8.err; grep "^{" $O/bench_c3_n8.json | cut -c1-200
timeout 600 python bench.py --gpus 8 --config 5 --steps 50 > $O/bench_c5_n8.json 2>> $O/bench_n8.err; grep "^{" $O/bench_c5_n8.json | cut -c1-200
timeout 600 python bench.py --gpus 8 --strong --muscles 524288 --steps 20 --no-e2e-variants --no-gather > $O/bench_strong_n8.json 2>> $O/bench_n8.err; grep "^{" $O/bench_strong_n8.json | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_n8.log 2>&1; tail -3 $O/pytest_n8.log
tail -5 $O/bench_n8.err
