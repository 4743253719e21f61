"""torchrun check of ShardedMuscleBatch on real GPUs (NCCL): sharded == unsharded.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/dist_check.py
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
    if rank == 0:
        print(f"[dist_check] {precision}: world {world}, {M} muscles sharded {sb.sizes}: sharded == unsharded (bitwise)", flush=True)
dist.destroy_process_group()
