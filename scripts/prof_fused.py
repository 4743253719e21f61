"""Minimal driver for ncu: a few fused windows of the headline config."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
kind = sys.argv[4] if len(sys.argv) > 4 else "potvin"
spec = MuscleSpec.potvin(120) if kind == "potvin" else MuscleSpec.standard()
b = MuscleBatch(spec, M, dev, prec, fresh_outputs=False)
e = ((67.0 if kind == "potvin" else 1.0) * torch.rand(M, device=dev)).to(b.dtype).repeat(K, 1).contiguous()
out = torch.empty(K, M, dtype=b.dtype, device=dev)
for _ in range(4):
    b.step_many(e, 0.02, out=out)
torch.cuda.synchronize()
print("ok", b.fused_geometry(K))
