"""Host <-> device copy ceiling of this box, all ranks at once (the limit of bench.py's end-to-end legs).

    python scripts/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29570 scripts/pcie_probe.py                  # N GPUs copying concurrently

Per rank: pinned 262 MB buffers (one fused window of config #2's excitations / totals), H2D only, D2H only and
both directions at once; the time of a phase is the MAX over ranks, GB/s are per rank and aggregate."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=250)
a = ap.parse_args()
h = bench.Harness(a)
res = bench.host_copy_probe(h, a.mb * 1024 * 1024 // 4 * 4)
if h.rank == 0:
    res["aggregate_gbs"] = {k: round(v * h.world, 1) for k, v in res.items() if k.endswith("per_rank") or k.endswith("direction")}
    print(json.dumps(res), flush=True)
if h.world > 1:
    h.barrier()
    h.dist.destroy_process_group()
