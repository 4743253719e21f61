"""Short fused windows (RL control intervals): Potvin(120) x 65,536, K steps per launch (development aid).
Run with MB_FUSED_PERSIST=0 / 1 to compare the one-pair-per-warp launch with the persistent prefetching variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
tag = "persist=" + os.environ.get("MB_FUSED_PERSIST", "auto")
for kind, spec, scale in (("potvin", MuscleSpec.potvin(120), 67.0), ("standard", MuscleSpec.standard(), 1.0)):
    b = MuscleBatch(spec, 65536, dev, "f32", fresh_outputs=False)
    for K in (8, 16, 32, 64, 128, 256, 512, 1000):
        e = scale * torch.rand(K, 65536, device=dev)
        out = torch.empty(K, 65536, device=dev)
        # back-to-back launches without a sync in between: the GPU-side cost of a window, not the Python call overhead
        def burst():
            for _ in range(10):
                b.step_many(e, 0.02, out=out)
        best, _ = time_cuda(burst, iters=3, warm=1)
        best /= 10
        print(f"[{tag}] {kind:8s} K={K:5d}: {65536 * 120 * K / best:.3e} unit-steps/s ({best * 1e6:8.1f} us per window)", flush=True)
