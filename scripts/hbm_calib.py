"""HBM reference points with torch elementwise kernels (development aid): what a plain
streaming kernel reaches for read-only / write-only / in-place read+write traffic."""
import torch
dev = torch.device("cuda:0")
N = 262_144_000
a = torch.rand(N, dtype=torch.float64, device=dev)
b = torch.empty_like(a)
f = torch.empty(N, dtype=torch.float32, device=dev)
def t(fn, nbytes, label, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    print(f"{label:40s} {nbytes / best / 1e9:8.1f} GB/s ({best * 1e6:.0f} us)")
t(lambda: b.copy_(a), 16 * N, "copy f64 (R8 W8)")
t(lambda: a.add_(1e-9), 16 * N, "in-place add f64 (R8 W8 same array)")
t(lambda: a.sum(), 8 * N, "sum f64 (R8)")
t(lambda: b.fill_(1.0), 8 * N, "fill f64 (W8)")
t(lambda: torch.add(a, 1.0, out=b), 16 * N, "add out-of-place f64 (R8 W8)")
def mix():
    a.add_(1e-9); f.fill_(0.5)
t(mix, 20 * N, "in-place add f64 + fill f32 (R8 W12), 2 kernels")
