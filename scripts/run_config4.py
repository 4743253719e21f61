"""BASELINE config #4 under torchrun: 262,144 PotvinFuglevandMuscle(120) in FP64 sharded over the
ranks (32,768 per GPU on 8), ramp-hold-rest excitation over 10,000 steps in fused windows of 1,000,
per-muscle totals gathered on every rank each window.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29540 scripts/run_config4.py
"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleSpec, ShardedMuscleBatch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per_rank, K, W, dt = 32768, 10000, 1000, 0.02
M = per_rank * world
sb = ShardedMuscleBatch(MuscleSpec.potvin(120), M, dev, "f64", fresh_outputs=False)
sb.enable_peer_gather(W)
g = torch.Generator(device=dev).manual_seed(4 + rank)
emax = 67.0 * (0.5 + 0.5 * torch.rand(per_rank, device=dev, generator=g, dtype=torch.float64))
k = torch.arange(K, device=dev, dtype=torch.float64)
shape = torch.where(k < 2000, k / 2000.0, torch.where(k < 6000, torch.ones_like(k), torch.zeros_like(k)))

def protocol(gather):
    sb.local.reset()
    acc = torch.zeros((), dtype=torch.float64, device=dev)
    for w in range(0, K, W):
        exc = shape[w:w + W, None] * emax[None, :]
        if gather == "peer":
            views = sb.step_many_peer_gather(exc, dt, dense=False)
            acc += sum(v[-1].sum() for v in views)          # consume every rank's totals (last step of the window)
        elif gather == "nccl":
            acc += sb.step_many(exc, dt, gather=True)[-1].sum()
        else:
            acc += sb.step_many(exc, dt)[-1].sum()
    return acc

res = {}
for mode in ("none", "nccl", "peer"):
    protocol(mode)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    acc = protocol(mode)
    b.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) * 1e-3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = (float(t.item()), float(acc.item()))
if rank == 0:
    us = M * 120 * K
    print(f"[config4] {M} muscles x 120 units x {K} steps, FP64, world {world}:")
    for mode, (t, acc) in res.items():
        print(f"[config4]   gather={mode:5s} {t * 1e3:8.1f} ms  {us / t:.3e} unit-steps/s   (checksum {acc:.6e})")
    assert abs(res["nccl"][1] - res["peer"][1]) <= 1e-9 * abs(res["nccl"][1]), "NCCL and fused gather disagree"
dist.destroy_process_group()
