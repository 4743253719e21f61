"""Minimal driver for ncu: hopper-like rows (4 short muscles) x 65,536 agents, env steps (k1_flat + fatigue epilogue)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
env = MuscleGroupEnv([MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)], 65536, dev, sys.argv[1] if len(sys.argv) > 1 else "f32")
a = torch.rand(65536, 4, device=dev)
for _ in range(4):
    env.step(a, 0.02)
torch.cuda.synchronize()
print("ok")
