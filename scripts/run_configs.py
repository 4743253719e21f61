"""BASELINE.json configs #2-#5 at full (per-GPU) size on one B200: throughput table.
Config #1 is the reference's own CPU case (see bench.py --impl reference).  Development aid; the
contract line is bench.py's."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
HBM = 6545.9
rows = []

# 2: PotvinMuscle 120 x 65,536, fused 1,000-step window, FP32 and FP64
for prec, M in (("f32", 65536), ("f64", 65536)):
    b = MuscleBatch(MuscleSpec.potvin(120), M, dev, prec, fresh_outputs=False)
    e = (67.0 * torch.rand(M, device=dev)).to(b.dtype).repeat(1000, 1).contiguous()
    out = torch.empty(1000, M, dtype=b.dtype, device=dev)
    best, _ = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=5, warm=2)
    rows.append((f"#2 Potvin(120) x {M}, fused 1,000 steps, {prec}", M * 120 * 1000 / best, f"{best * 1e3:.2f} ms / window", ""))
    del b, e, out

# 3: arm (30 muscles, 35,032 units) x 4,096 agents, per-step random excitation, single-step launches + fatigue
forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]
env = MuscleGroupEnv([MuscleSpec.standard(f) for f in forces], 4096, dev, "f32")
acts = [torch.rand(4096, 30, device=dev) for _ in range(4)]
i = [0]
def frame():
    env.step(acts[i[0] & 3], 0.02); i[0] += 1
best, _ = time_cuda(frame, iters=20, warm=3)
u = env.batch.n_units
rows.append(("#3 arm 30 muscles / 35,032 units x 4,096 agents, env step (outputs + fatigue), f32", u / best,
             f"{best * 1e6:.0f} us / env step", f"{u * 20 / best / 1e9:.0f} GB/s = {u * 20 / best / 1e9 / HBM:.0%} of HBM peak"))
del env

# 4: one GPU's shard of 262,144 Potvin(120) in FP64: ramp-hold-rest over 10,000 steps, windows of 1,000
M = 32768
b = MuscleBatch(MuscleSpec.potvin(120), M, dev, "f64", fresh_outputs=False)
k = torch.arange(10000, device=dev, dtype=torch.float64)
shape = torch.where(k < 2000, k / 2000.0, torch.where(k < 6000, torch.ones_like(k), torch.zeros_like(k)))
emax = 67.0 * (0.5 + 0.5 * torch.rand(M, device=dev, dtype=torch.float64))
out = torch.empty(1000, M, dtype=torch.float64, device=dev)
def protocol():
    b.reset()
    for w in range(10):
        b.step_many(shape[w * 1000:(w + 1) * 1000, None] * emax[None, :], 0.02, out=out)
best, _ = time_cuda(protocol, iters=3, warm=1)
rows.append((f"#4 shard: Potvin(120) x {M}, f64, ramp-hold-rest 10,000 steps (10 windows)", M * 120 * 10000 / best, f"{best * 1e3:.0f} ms / protocol", ""))
del b, out

# 5: StandardMuscle 2,000 units x 131,072, one launch per frame
for prec, M, bpu in (("f32", 131072, 20), ("f64", 65536, 24)):
    b = MuscleBatch(MuscleSpec.standard(527.1), M, dev, prec, fresh_outputs=False)
    xs = [torch.rand(M, device=dev).to(b.dtype) for _ in range(4)]
    def frame5():
        b.step(xs[i[0] & 3], 0.02); i[0] += 1
    best, _ = time_cuda(frame5, iters=20, warm=3)
    rows.append((f"#5 StandardMuscle(2,000 units) x {M}, single-step launches, {prec}", b.n_units / best, f"{best * 1e6:.0f} us / frame",
                 f"{b.n_units * bpu / best / 1e9:.0f} GB/s = {b.n_units * bpu / best / 1e9 / HBM:.0%} of HBM peak"))
    del b, xs

print(f"{'config':92s} {'unit-steps/s':>13s}  time                 roofline")
for name, us, t, r in rows:
    print(f"{name:92s} {us:13.3e}  {t:20s} {r}")
