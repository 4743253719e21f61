"""Fused windows of mid-sized muscles (129..480 units): block kernel vs block-per-muscle kernel (MB_FUSED_WIDE=1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
tag = "wide=" + os.environ.get("MB_FUSED_WIDE", "0")
K = 1000
for label, spec, scale in (("Standard n=340", MuscleSpec.standard(90.0), 1.0), ("Potvin n=200", MuscleSpec.potvin(200), 67.0),
                           ("Potvin n=340", MuscleSpec.potvin(340), 67.0), ("Potvin n=480", MuscleSpec.potvin(480), 67.0),
                           ("Standard n=169", MuscleSpec.standard(45.0), 1.0), ("Standard n=256", MuscleSpec.standard(68.0), 1.0)):
    M = 8_000_000 // spec.n
    b = MuscleBatch(spec, M, dev, "f32", fresh_outputs=False)
    e = scale * torch.rand(K, M, device=dev)
    out = torch.empty(K, M, device=dev)
    best, _ = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=3, warm=1)
    print(f"[{tag}] {label:16s} x {M}: {b.n_units * K / best:.3e} unit-steps/s ({best * 1e3:.2f} ms)", flush=True)
    del b, e, out
