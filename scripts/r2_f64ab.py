"""FP64 fused window: config #2 shape (uniform random excitation per muscle, constant) and config #4's protocol
(ramp-hold-rest).  Run under different MB_LIB builds to compare (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev = torch.device("cuda:0")
tag = os.path.basename(os.environ.get("MB_LIB", "default"))
M, K = 32768, 1000
b = MuscleBatch(MuscleSpec.potvin(120), M, dev, "f64", fresh_outputs=False)
out = torch.empty(K, M, dtype=torch.float64, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
e2 = (67.0 * torch.rand(M, device=dev, generator=g, dtype=torch.float64)).repeat(K, 1).contiguous()
best, _ = time_cuda(lambda: b.step_many(e2, 0.02, out=out), iters=4, warm=1)
print(f"[{tag}] config2 f64 (e ~ U[0,67) constant): {M * 120 * K / best:.3e} unit-steps/s ({best * 1e3:.2f} ms)", flush=True)
emax = 67.0 * (0.5 + 0.5 * torch.rand(M, device=dev, generator=g, dtype=torch.float64))
k = torch.arange(10000, device=dev, dtype=torch.float64)
shape = torch.where(k < 2000, k / 2000.0, torch.where(k < 6000, torch.ones_like(k), torch.zeros_like(k)))
excs = [(shape[w:w + K, None] * emax[None, :]).contiguous() for w in range(0, 10000, K)]
def protocol():
    b.reset()
    for e in excs:
        b.step_many(e, 0.02, out=out)
best, _ = time_cuda(protocol, iters=3, warm=1)
print(f"[{tag}] config4 protocol (ramp 2k / hold 4k / rest 4k): {M * 120 * 10000 / best:.3e} unit-steps/s ({best * 1e3:.1f} ms per protocol)", flush=True)
for name, idx in (("ramp", 1), ("hold", 3), ("rest", 7)):
    best, _ = time_cuda(lambda: b.step_many(excs[idx], 0.02, out=out), iters=3, warm=1)
    print(f"[{tag}]    window in {name}: {best * 1e3:.2f} ms", flush=True)
