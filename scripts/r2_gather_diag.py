"""Why does the fused gather cost more inside bench.py than in r2_gather_ab.py?  Per-rank, per-window device times of
the none / peer legs, cold (right after start-up) and after ~10 s of FP32 load, with each GPU's SM clock and power."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec, ShardedMuscleBatch
import pynvml

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pynvml.nvmlInit()
uuid = str(torch.cuda.get_device_properties(local).uuid)
nvh = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
K, dt, per_rank = 1000, 0.02, 32768
sink = torch.zeros((), dtype=torch.float64, device=dev)
g = torch.Generator(device=dev).manual_seed(4 + rank)
e = (67.0 * (0.5 + 0.5 * torch.rand(per_rank, device=dev, generator=g, dtype=torch.float64))).repeat(K, 1).contiguous()


def leg(tag, mode, it=6):
    sb = ShardedMuscleBatch(MuscleSpec.potvin(120), per_rank * world, dev, "f64", fresh_outputs=False)
    if mode == "peer":
        sb.enable_peer_gather(K)
    def fn():
        if mode == "peer":
            sink.add_(sb.step_many_peer_gather(e, dt, dense=False)[(rank + 1) % world][-1].sum())
        else:
            sink.add_(sb.step_many(e, dt)[-1].sum())
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(it + 1)]
    ev[0].record()
    for i in range(it):
        fn()
        ev[i + 1].record()
    mhz = pynvml.nvmlDeviceGetClockInfo(nvh, pynvml.NVML_CLOCK_SM)     # while the queue is still running
    watts = pynvml.nvmlDeviceGetPowerUsage(nvh) / 1e3
    torch.cuda.synchronize(); dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(it)]
    row = torch.tensor(per + [mhz, watts], dtype=torch.float64, device=dev)
    rows = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(rows, row)
    if rank == 0:
        tot = max(sum(r[:it].tolist()) for r in rows) / it
        print(f"[diag] {tag} {mode}: max-over-ranks mean {tot:.3f} ms/window", flush=True)
        for r, x in enumerate(rows):
            x = x.tolist()
            print(f"[diag]   rank {r}: " + " ".join(f"{v:.2f}" for v in x[:it]) + f" | {x[it]:.0f} MHz {x[it + 1]:.0f} W", flush=True)
    del sb
    torch.cuda.empty_cache()


leg("cold", "none"); leg("cold", "peer"); leg("cold", "none"); leg("cold", "peer")
# ~10 s of the FP32 headline window
b = MuscleBatch(MuscleSpec.potvin(120), 65536, dev, "f32", fresh_outputs=False)
e32 = (67.0 * torch.rand(65536, device=dev)).repeat(K, 1).contiguous()
for _ in range(1700):
    b.step_many(e32, dt)
torch.cuda.synchronize()
leg("hot", "none"); leg("hot", "peer"); leg("hot", "none"); leg("hot", "peer")
dist.destroy_process_group()
