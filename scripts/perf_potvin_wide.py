"""Potvin muscles with many units, single step (development aid): A/B of library builds via MB_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
out = []
for n, rows in ((120, 1048576), (480, 262144), (2000, 65536)):
    b = MuscleBatch(MuscleSpec.potvin(n), rows, dev, "f32", fresh_outputs=False)
    x = torch.rand(rows, device=dev) * 67.0
    best, _ = time_cuda(lambda: b.step(x, 0.02), iters=10, warm=3)
    out.append(f"potvin{n} {best * 1e6:.0f}us {b.n_units * 36 / best / 1e9:.0f}GB/s")
    del b; torch.cuda.empty_cache()
print(" | ".join(out))
