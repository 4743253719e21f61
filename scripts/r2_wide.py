"""Fused windows of wide muscles and ragged rows (K2x) and of FP64 muscles (K2w64) -- development aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
K = 1000
def run(label, spec, rows, prec, K=K, scale=1.0):
    b = MuscleBatch(spec, rows, dev, prec, fresh_outputs=False)
    e = (scale * torch.rand(K, b.n_muscles, device=dev)).to(b.dtype)
    out = torch.empty(K, b.n_muscles, dtype=b.dtype, device=dev)
    l0 = b.launches
    best, mean = time_cuda(lambda: b.step_many(e, 0.02, out=out), iters=3, warm=1)
    print(f"{label:46s} {prec}: {b.n_units * K / best:.3e} unit-steps/s ({best * 1e3:8.2f} ms per window, {(b.launches - l0) // 4} launch(es))", flush=True)
    del b, e, out
    torch.cuda.empty_cache()
run("StandardMuscle(527.1) n=2000 x 32768", MuscleSpec.standard(527.1), 32768, "f32")
run("StandardMuscle(527.1) n=2000 x 131072", MuscleSpec.standard(527.1), 131072, "f32", K=250)
run("StandardMuscle(90) n=340 x 24576 (block kernel)", MuscleSpec.standard(90.0), 24576, "f32")
arm = [MuscleSpec.standard(round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1)) for i in range(30)]
run("arm 30 muscles x 1024 agents", arm, 1024, "f32", K=250)
hopper = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)]
run("hopper 4 muscles x 65536 agents", hopper, 65536, "f32", K=250)
run("Potvin(700) x 16384", MuscleSpec.potvin(700), 16384, "f32", scale=67.0)
run("Potvin(120) x 65536", MuscleSpec.potvin(120), 65536, "f64", scale=67.0)
run("Potvin(120) x 32768", MuscleSpec.potvin(120), 32768, "f64", scale=67.0)
run("Standard(120) x 32768", MuscleSpec.standard(), 32768, "f64")
run("Potvin(64) x 65536", MuscleSpec.potvin(64), 65536, "f64", scale=67.0)
