"""Ragged rows, single step: flat-tiled kernel (default) vs one item per muscle (MB_K1_FLAT=0) -- development aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
tag = "flat=" + os.environ.get("MB_K1_FLAT", "auto") + " group=" + os.environ.get("MB_K1_FLAT_GROUP", "dflt")
arm = [MuscleSpec.standard(round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1)) for i in range(30)]
hopper = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)]
for label, spec, rows in (("arm x 4096", arm, 4096), ("hopper x 65536", hopper, 65536), ("hopper x 4096", hopper, 4096),
                          ("8 x n=256 x 65536", [MuscleSpec.standard(68.0)] * 8, 65536),
                          ("30 x n=1137 x 4096", [MuscleSpec.standard(300.0)] * 30, 4096)):
    for prec in ("f32", "f64"):
        b = MuscleBatch(spec, rows, dev, prec, fresh_outputs=False)
        bpu = 20 if prec == "f32" else 24
        x = torch.rand(rows, b.S, device=dev, dtype=b.dtype)
        for what, fn in (("step", lambda: b.step(x, 0.02)), ("step_env", lambda: b.step_env(x, 0.02))):
            best, _ = time_cuda(fn, iters=15, warm=3)
            print(f"[{tag}] {label:20s} {prec} {what:8s}: {best * 1e6:7.0f} us  {b.n_units * bpu / best / 1e9:6.0f} GB/s ({b.n_units * bpu / best / 6545.9e9 * 100:4.1f} % of HBM peak)", flush=True)
        del b
        torch.cuda.empty_cache()
