#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_round2.py -m gpu -x -q -k "sharded or gather" > $O/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_n2.log
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_n2.err; cat $O/bench_n2.json
timeout 600 python bench.py --gpus 2 --config 4 --steps 20 > $O/bench_c4_n2.json 2>> $O/bench_n2.err; echo "bench c4 n2 rc=$?"; cat $O/bench_c4_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29570 scripts/pcie_probe.py > $O/pcie_n2.json 2>> $O/bench_n2.err; cat $O/pcie_n2.json
timeout 300 python scripts/pcie_probe.py > $O/pcie_n1.json 2>> $O/bench_n2.err; cat $O/pcie_n1.json
