"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
torch.manual_seed(0)
for prec in ("f32", "f64"):
    # K2: warp kernel (n = 120, 33), block kernel (explicit U, n = 340, n = 2), constant + unaligned batches
    for spec, M, K, upt in ((MuscleSpec.potvin(120), 37, 70, 0), (MuscleSpec.potvin(33), 5, 40, 0),
                            (MuscleSpec.standard(), 19, 45, 4), (MuscleSpec.standard(90.0), 3, 20, 0),
                            (MuscleSpec.potvin(2), 9, 17, 0)):
        b = MuscleBatch(spec, M, dev, prec)
        scale = 1.0 if spec.name == "standard" else 67.0
        e = (scale * torch.rand(K, M, device=dev)).to(b.dtype)
        b.step_many(e, 0.02, units_per_thread=upt)
        b.step_many(e[0], 0.02, steps=K)
        b.step_many(e, 0.02, forces_every=7, return_masks=True, units_per_thread=upt)
    # K1: streaming (even L, uniform and ragged), generic (odd L, per-unit input, stages), env epilogue, K3
    arm = [MuscleSpec.standard(f) for f in (25.0, 31.0, 40.0)]
    for spec, rows in ((MuscleSpec.standard(), 21), (MuscleSpec.potvin(120), 9), (arm, 11), (MuscleSpec.potvin(33), 7)):
        b = MuscleBatch(spec, rows, dev, prec)
        scale = 67.0 if (not isinstance(spec, list) and spec.name == "potvin") else 1.0
        x = (scale * torch.rand(b.rows, b.S, device=dev)).to(b.dtype)
        for _ in range(3):
            b.step(x, 0.02)
        b.step((scale * torch.rand(b.rows, b.L, device=dev)).to(b.dtype), 0.02)
        b.step_env(x, 0.02, gain=torch.rand(b.rows, b.S, device=dev))
        b.get_peripheral_fatigue()
torch.cuda.synchronize()
print("sanitize_small ok")
