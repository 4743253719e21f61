"""Headline fused-window timing (development aid): 65,536 x 120 units x 1,000 steps, FP32 mode."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
for kind, spec, scale in (("potvin", MuscleSpec.potvin(120), 67.0), ("standard", MuscleSpec.standard(), 1.0)):
    M, K = (65536, 1000) if prec == "f32" else (16384, 1000)
    b = MuscleBatch(spec, M, dev, prec, fresh_outputs=False)
    exc = (scale * torch.rand(K, M, device=dev)).to(b.dtype)
    out = torch.empty(K, M, device=dev, dtype=b.dtype)
    best, mean = time_cuda(lambda: b.step_many(exc, 0.02, out=out), iters=5, warm=2)
    print(f"{kind} {prec} auto {b.fused_geometry(K)}: {M * 120 * K / best / 1e9:.1f} G unit-steps/s ({best * 1e3:.2f} ms)", flush=True)
