"""Minimal driver for ncu: config #3 env steps (ragged streaming kernel + fatigue epilogue)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]
env = MuscleGroupEnv([MuscleSpec.standard(f) for f in forces], 4096, dev, sys.argv[1] if len(sys.argv) > 1 else "f32")
a = torch.rand(4096, 30, device=dev)
for _ in range(4):
    env.step(a, 0.02)
torch.cuda.synchronize()
print("ok")
