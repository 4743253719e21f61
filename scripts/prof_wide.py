"""Minimal driver for ncu: fused windows of StandardMuscle(527.1 N) = 2,000 units (k2x_f32)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
M, K = 8192, 250
b = MuscleBatch(MuscleSpec.standard(527.1), M, dev, "f32", fresh_outputs=False)
e = torch.rand(K, M, device=dev)
out = torch.empty(K, M, device=dev)
for _ in range(4):
    b.step_many(e, 0.02, out=out)
torch.cuda.synchronize()
print("ok")
