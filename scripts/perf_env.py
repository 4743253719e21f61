"""Config #3 shape: the 30-muscle arm x 4,096 agents, one env step per launch (development aid)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]
specs = [MuscleSpec.standard(f) for f in forces]
for prec in ("f32", "f64"):
    env = MuscleGroupEnv(specs, 4096, dev, prec)
    acts = [torch.rand(4096, 30, device=dev) for _ in range(4)]
    i = [0]
    def frame():
        env.step(acts[i[0] & 3], 0.02); i[0] += 1
    best, mean = time_cuda(frame, iters=20, warm=3)
    units = env.batch.n_units
    bpu = 20 if prec == "f32" else 24
    print(f"arm x 4096 agents {prec}: {units} units, {units / best / 1e9:.1f} G unit-steps/s, {units * bpu / best / 1e9:.0f} GB/s algorithmic, {best * 1e6:.0f} us per env step ({4096 / best / 1e6:.2f} M agent-steps/s)")
    def frame2():
        env.batch.step(acts[i[0] & 3], 0.02); env.batch.get_peripheral_fatigue(); i[0] += 1
    best2, _ = time_cuda(frame2, iters=20, warm=3)
    print(f"   two-kernel path (step + fatigue pass): {best2 * 1e6:.0f} us")
