"""FP64-mode single-step shapes (development aid): used for A/B runs of library builds via MB_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
dt = torch.float64 if prec == "f64" else torch.float32
forces = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]
out = []
for name, spec, rows, bpu in (("std2000", MuscleSpec.standard(527.1), 65536, 24), ("potvin120", MuscleSpec.potvin(120), 524288, 40),
                              ("arm-env", [MuscleSpec.standard(f) for f in forces], 4096, 24)):
    if prec == "f32":
        bpu -= 4
    b = MuscleBatch(spec, rows, dev, prec, fresh_outputs=False)
    x = torch.rand(rows, b.S, device=dev, dtype=dt) * (67.0 if name == "potvin120" else 1.0)
    fn = (lambda: b.step_env(x, 0.02)) if name == "arm-env" else (lambda: b.step(x, 0.02))
    best, _ = time_cuda(fn, iters=10, warm=3)
    out.append(f"{name} {best * 1e6:.0f}us {b.n_units * bpu / best / 1e9:.0f}GB/s")
    del b
    torch.cuda.empty_cache()
print(prec, " | ".join(out))
