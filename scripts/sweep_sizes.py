"""Throughput against batch size and window length (one B200, FP32 mode; development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
print("fused window (Potvin 120 units, K = 1000): muscles -> unit-steps/s")
for M in (256, 1024, 4096, 16384, 65536, 262144):
    b = MuscleBatch(MuscleSpec.potvin(120), M, dev, "f32", fresh_outputs=False)
    e = (67.0 * torch.rand(1000, M, device=dev)); out = torch.empty(1000, M, device=dev)
    best, _ = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=3, warm=1)
    print(f"  M={M:7d}: {M * 120 * 1000 / best:.3e}  ({best * 1e3:.2f} ms)", flush=True)
    del b, e, out
print("fused window (Potvin 120 units, M = 65536): steps per launch -> unit-steps/s")
b = MuscleBatch(MuscleSpec.potvin(120), 65536, dev, "f32", fresh_outputs=False)
for K in (8, 32, 128, 512, 2048):
    e = (67.0 * torch.rand(K, 65536, device=dev)); out = torch.empty(K, 65536, device=dev)
    best, _ = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=3, warm=1)
    print(f"  K={K:5d}: {65536 * 120 * K / best:.3e}  ({best * 1e3:.3f} ms)", flush=True)
del b
print("single step (StandardMuscle 2000 units): muscles -> GB/s algorithmic (20 B/unit-step)")
for M in (128, 1024, 8192, 32768, 131072):
    b = MuscleBatch(MuscleSpec.standard(527.1), M, dev, "f32", fresh_outputs=False)
    x = torch.rand(M, device=dev)
    best, _ = time_cuda(lambda: b.step(x, 0.02), iters=20, warm=3)
    print(f"  M={M:7d}: {M * 2000 * 20 / best / 1e9:7.0f} GB/s  ({best * 1e6:.1f} us)", flush=True)
    del b
