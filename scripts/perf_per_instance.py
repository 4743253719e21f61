"""Per-instance parameter tables on the headline shape (development aid)."""
import os, sys, numpy as np, torch
from dataclasses import replace
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
M, K = 16384, 1000
base = MuscleSpec.potvin(120)
rng = np.random.default_rng(0)
specs = [replace(base, fibers=replace(base.fibers, max_fatigue_rate=float(rng.uniform(0.015, 0.03))),
                 pool=replace(base.pool, adaptation_magnitude=float(rng.uniform(0.4, 0.8)))) for _ in range(M)]
for prec in ("f32", "f64"):
    b = MuscleBatch(base, M, dev, prec, fresh_outputs=False, row_specs=specs)
    e = (67.0 * torch.rand(K, M, device=dev)).to(b.dtype)
    out = torch.empty(K, M, dtype=b.dtype, device=dev)
    best, _ = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=3, warm=1)
    print(f"fused, per-instance tables, {prec}: {M * 120 * K / best / 1e9:.0f} G unit-steps/s ({b.fused_geometry(K)})")
    x = e[0].contiguous()
    best, _ = time_cuda(lambda: b.step(x, 0.02), iters=10, warm=2)
    bpu = (36 + 48) if prec == "f32" else (40 + 80)
    print(f"single step, per-instance tables, {prec}: {M * 120 / best / 1e9:.0f} G unit-steps/s, {M * 120 * bpu / best / 1e9:.0f} GB/s incl. the table rows")
