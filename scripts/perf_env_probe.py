"""Which property of config #3 costs bandwidth?  Separates ragged segments, the per-item reduction
and the fatigue epilogue (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
bpu = 20 if prec == "f32" else 24
dt = torch.float32 if prec == "f32" else torch.float64

def run(name, spec, rows):
    b = MuscleBatch(spec, rows, dev, prec, fresh_outputs=False)
    x = torch.rand(rows, b.S, device=dev, dtype=dt)
    for label, fn in (("step", lambda: b.step(x, 0.02)), ("step_env", lambda: b.step_env(x, 0.02))):
        best, _ = time_cuda(fn, iters=12, warm=3)
        print(f"{name:46s} {label:8s}: {b.n_units:>10d} units {best * 1e6:7.0f} us  {b.n_units * bpu / best / 1e9:6.0f} GB/s", flush=True)
    del b
    torch.cuda.empty_cache()

forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]
run("arm (30 ragged muscles) x 4096", [MuscleSpec.standard(f) for f in forces], 4096)
run("30 equal muscles n=1137 x 4096", [MuscleSpec.standard(300.0)] * 30, 4096)
run("30 equal muscles n=1024(ish) x 4096", [MuscleSpec.standard(270.0)] * 30, 4096)
run("uniform n=1137 x 122880", MuscleSpec.standard(300.0), 122880)
run("uniform n=2000 x 71680", MuscleSpec.standard(527.0), 71680)
run("8 muscles n=256 x 65536", [MuscleSpec.standard(68.0)] * 8, 65536)
run("uniform n=256 x 524288", MuscleSpec.standard(68.0), 524288)
