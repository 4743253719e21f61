"""PCIe reference points (development aid): pinned H2D, D2H, and both at once."""
import torch, time
dev = torch.device("cuda:0")
n = 262_144_000 // 4
h1 = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d1 = torch.empty(n, dtype=torch.float32, device=dev); d2 = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, label, nbytes):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{label:28s} {dt * 1e3:6.2f} ms  {nbytes / dt / 1e9:6.1f} GB/s per direction")
t(lambda: d1.copy_(h1, non_blocking=True), "H2D 262 MB", 4 * n)
t(lambda: h2.copy_(d2, non_blocking=True), "D2H 262 MB", 4 * n)
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t(both, "H2D + D2H concurrently", 4 * n)
