"""Per-unit excitation [M, n] through the single-step kernel (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
for spec, M, label, bpu in ((MuscleSpec.standard(527.1), 131072, "standard n=2000", 24), (MuscleSpec.potvin(120), 1 << 20, "potvin n=120", 40)):
    b = MuscleBatch(spec, M, dev, "f32", fresh_outputs=False)
    x = torch.rand(M, b.L, device=dev) * (1.0 if spec.name == "standard" else 67.0)
    best, _ = time_cuda(lambda: b.step(x, 0.02), iters=20, warm=3)
    print(f"per-unit input, {label} f32: {b.n_units / best / 1e9:.1f} G unit-steps/s, {b.n_units * bpu / best / 1e9:.0f} GB/s algorithmic ({bpu} B/unit-step)")
