"""torchrun check of ShardedMuscleBatch on real GPUs (NCCL): sharded == unsharded.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/dist_check.py [--timing]

tests/test_gpu_multi.py launches this (without --timing) when the box has at least two GPUs.
"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec, ShardedMuscleBatch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
M, K, dt = 1000 + 3, 40, 0.02                                   # uneven shards on purpose
g = torch.Generator(device="cpu").manual_seed(7)
exc = (67.0 * torch.rand(K + 1, M, generator=g)).to(dev)        # identical on every rank
for precision in ("f32", "f64"):
    sb = ShardedMuscleBatch(MuscleSpec.potvin(120), M, dev, precision)
    one = sb.step(exc[0].to(sb.local.dtype), dt, gather=True, global_input=True)
    win = sb.step_many(exc[1:].to(sb.local.dtype), dt, gather=True, global_input=True)
    h = sb.gather_totals(sb.step(exc[0, sb.row_lo:sb.row_hi].to(sb.local.dtype), dt), async_op=True)
    again = h.wait()
    ref = MuscleBatch(MuscleSpec.potvin(120), M, dev, precision)
    want_one = ref.step(exc[0].to(ref.dtype), dt)
    want_win = ref.step_many(exc[1:].to(ref.dtype), dt)
    want_again = ref.step(exc[0].to(ref.dtype), dt)
    torch.cuda.synchronize()
    assert torch.equal(one, want_one), "single step differs"
    assert torch.equal(win, want_win), "fused window differs"
    assert torch.equal(again, want_again), "async gather differs"
    assert torch.equal(sb.local.cpf, ref.cpf[sb.row_lo:sb.row_hi])
    # the gather fused into the kernel: totals stored straight into every rank's buffer over NVLink
    sp = ShardedMuscleBatch(MuscleSpec.potvin(120), M, dev, precision)
    rp = MuscleBatch(MuscleSpec.potvin(120), M, dev, precision)
    sp.enable_peer_gather(K)
    for w in range(3):
        got = sp.step_many(exc[1:].to(sp.local.dtype), dt, gather="peer", global_input=True).clone()
        want = rp.step_many(exc[1:].to(rp.dtype), dt)
        torch.cuda.synchronize()
        assert torch.equal(got, want), f"peer gather differs in window {w}"
    if rank == 0:
        print(f"[dist_check] {precision}: world {world}, {M} muscles sharded {sb.sizes}: sharded == unsharded (bitwise), "
              f"NCCL gather and kernel-fused peer gather", flush=True)

if "--timing" not in sys.argv:                               # tests/test_gpu_multi.py runs the correctness part only
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)

# ---- cost of the gather on the headline shape (per rank 65,536 muscles x 1,000 steps, FP32 mode) ----
Mr, K2 = 65536, 1000
big = ShardedMuscleBatch(MuscleSpec.potvin(120), Mr * world, dev, "f32", fresh_outputs=False)
big.enable_peer_gather(K2)
e = (67.0 * torch.rand(Mr, device=dev)).repeat(K2, 1).contiguous()

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

t_none = timed(lambda: big.step_many(e, dt))
t_nccl = timed(lambda: big.step_many(e, dt, gather=True))
t_peer = timed(lambda: big.step_many(e, dt, gather="peer"))
t_peer_v = timed(lambda: big.step_many_peer_gather(e, dt, dense=False))
if rank == 0:
    us = Mr * world * 120 * K2
    print(f"[dist_check] window of {K2} steps, {Mr} muscles/rank, world {world}: no gather {t_none:.2f} ms "
          f"({us / t_none / 1e6:.0f} G unit-steps/s), NCCL all_gather after the kernel {t_nccl:.2f} ms, "
          f"gather fused into the kernel (peer stores) {t_peer:.2f} ms dense / {t_peer_v:.2f} ms as views", flush=True)
dist.destroy_process_group()
