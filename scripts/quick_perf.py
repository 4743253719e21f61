"""Ad-hoc timing of the kernels (development aid; bench.py is the contract)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleBatch, MuscleSpec
from pymuscle_b200 import _native as nat


def time_cuda(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sum(ts) / len(ts)


def probes(dev):
    lib = nat.lib()
    sink = torch.empty(148 * 8 * 1024, dtype=torch.float64, device=dev)
    for kind, name in ((0, "FFMA"), (1, "DFMA"), (2, "MUFU.EX2"), (3, "FFMA2(instr)"), (4, "FMNMX"), (5, "mix5F2A1M"), (6, "FMUL2(instr)"), (7, "FADD2(instr)")):
        ops = C.c_int64()
        iters = 4096 if kind != 1 else 1024
        def run():
            nat.check(lib.mb_peak_probe(kind, 148 * 8, 1024, iters, C.c_void_p(sink.data_ptr()), C.byref(ops),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        best, mean = time_cuda(run)
        print(f"probe {name:9s}: {ops.value / best / 1e12:8.2f} Tlane-op/s best, {ops.value / mean / 1e12:8.2f} mean", flush=True)


def fused(dev, precision, M=65536, n=120, K=1000, kind="potvin"):
    spec = MuscleSpec.potvin(n) if kind == "potvin" else MuscleSpec.standard()
    scale = 67.0 if kind == "potvin" else 1.0
    exc = (scale * torch.rand(K, M, device=dev)).to(torch.float32 if precision == "f32" else torch.float64)
    for upt in ((1, 2, 4) if precision == "f32" else (1, 2)):
        b = MuscleBatch(spec, M, dev, precision, fresh_outputs=False)
        out = torch.empty(K, M, dtype=b.dtype, device=dev)
        best, mean = time_cuda(lambda: b.step_many(exc, 0.02, out=out, units_per_thread=upt), iters=3, warm=1)
        print(f"fused {kind} {precision} U={upt} {b.fused_geometry(K, upt)}: {M * n * K / best / 1e9:9.1f} G unit-steps/s best "
              f"({best * 1e3:.2f} ms), mean {M * n * K / mean / 1e9:9.1f}", flush=True)


def single(dev, precision, M, spec, label, bytes_per_unit):
    b = MuscleBatch(spec, M, dev, precision, fresh_outputs=False)
    x = torch.rand(b.n_muscles, device=dev).to(b.dtype)
    best, mean = time_cuda(lambda: b.step(x, 0.02), iters=20, warm=3)
    print(f"single {label} {precision}: {b.n_units / best / 1e9:8.1f} G unit-steps/s, "
          f"{b.n_units * bytes_per_unit / best / 1e9:8.1f} GB/s algorithmic ({best * 1e6:.1f} us)", flush=True)


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    what = sys.argv[1:] or ["probes", "fused", "single"]
    if "probes" in what:
        probes(dev)
    if "fused" in what:
        fused(dev, "f32")
        fused(dev, "f32", kind="standard")
        fused(dev, "f64", M=16384)
    if "single" in what:
        single(dev, "f32", 131072, MuscleSpec.standard(527.1), "standard n=2000", 20)
        single(dev, "f64", 65536, MuscleSpec.standard(527.1), "standard n=2000", 24)
        single(dev, "f32", 1 << 20, MuscleSpec.potvin(120), "potvin n=120", 36)
        single(dev, "f32", 1 << 20, MuscleSpec.standard(), "standard n=120", 20)
