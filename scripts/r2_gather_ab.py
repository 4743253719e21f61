"""Cost of gathering every rank's totals on every rank, per fused window (torchrun, one process per GPU):
none / NCCL all_gather / kernel-fused peer stores (pair-major, muscle-major) / kernel-fused NVLink multicast stores.
FP64 mode: config #4's shape (32,768 muscles per rank); FP32 mode: config #2's (65,536 per rank)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleSpec, ShardedMuscleBatch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K, dt = 1000, 0.02

def timed(fn, it=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / it], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for prec, per_rank in ((("f64", 32768),) if os.environ.get("MB_AB_F64_ONLY") else (("f64", 32768), ("f32", 65536))):
    sink = torch.zeros((), dtype=torch.float64, device=dev)
    def make():
        return ShardedMuscleBatch(MuscleSpec.potvin(120), per_rank * world, dev, prec, fresh_outputs=False)
    sb = make()
    g = torch.Generator(device=dev).manual_seed(4 + rank)
    e = (67.0 * (0.5 + 0.5 * torch.rand(per_rank, device=dev, generator=g))).to(sb.local.dtype).repeat(K, 1).contiguous()
    res = {}
    res["none"] = timed(lambda: sink.add_(sb.step_many(e, dt)[-1].sum()))
    res["nccl"] = timed(lambda: sink.add_(sb.step_many(e, dt, gather=True)[-1].sum()))
    ref = None
    for layout in ("pair", "muscle"):
        for mc in (False, True):
            name = f"peer_{layout}" + ("_multicast" if mc else "")
            try:
                s2 = make()
                s2.enable_peer_gather(K, multicast=mc, layout=layout)
                res[name] = timed(lambda: sink.add_(sum(v[-1].sum() for v in s2.step_many_peer_gather(e, dt, dense=False))))
                # correctness: the gathered totals of one window equal the NCCL gather of a twin batch
                s3, s4 = make(), make()
                s3.enable_peer_gather(K, multicast=mc, layout=layout)
                got = s3.step_many_peer_gather(e, dt, dense=True).clone()
                want = s4.step_many(e, dt, gather=True)
                torch.cuda.synchronize()
                assert torch.equal(got, want), f"{name}: gathered totals differ from the NCCL gather"
                del s2, s3, s4
            except Exception as exc:
                res[name] = f"{type(exc).__name__}: {str(exc)[:120]}"
            torch.cuda.empty_cache()
    if rank == 0:
        print(f"[gather_ab] {prec} world {world} {per_rank} muscles/rank x {K} steps (ms per window): " +
              ", ".join(f"{k} {v:.3f}" if isinstance(v, float) else f"{k} <{v}>" for k, v in res.items()), flush=True)
    del sb
    torch.cuda.empty_cache()
dist.destroy_process_group()
