"""Minimal driver for ncu: a few single-step launches (config #5 shape by default)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
kind = sys.argv[3] if len(sys.argv) > 3 else "standard2000"
spec = {"standard2000": lambda: MuscleSpec.standard(527.1), "standard": MuscleSpec.standard,
        "potvin": lambda: MuscleSpec.potvin(120)}[kind]()
b = MuscleBatch(spec, M, dev, prec, fresh_outputs=False)
x = torch.rand(b.n_muscles, device=dev).to(b.dtype) * (67.0 if kind == "potvin" else 1.0)
for _ in range(4):
    b.step(x, 0.02)
torch.cuda.synchronize()
print("ok", b.n_units)
