#!/usr/bin/env python
"""
bench.py -- motor-unit-steps/s of the PyMuscle hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): 65,536 independent PotvinFuglevandMuscle(120)
per GPU, FP32 mode, one "step" = one fused window of 1,000 simulation steps
(dt = 0.02) = 7.86e9 motor-unit-steps.  Excitations are a [1000, 65536] float32
tensor (constant per muscle, e_m = 67*u_m, but the kernel reads it per step, as
it would a time-varying signal).  Weak scaling: every rank owns its own 65,536
muscles; muscles are independent so there is no collective on the data path.

One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "motor-unit-steps/s"
N_UNITS = 120
OPS_PER_UNIT_STEP = {"potvin": 38, "standard": 31}        # SURVEY.md Appendix A.2 canonical counts
K1_BYTES_PER_UNIT_STEP = {"standard_f32": 20, "potvin_f32": 36}


# ======================================================================================
# CPU arm: the reference's algorithm, one Python object per muscle, looped over muscles
# (examples/envs/muscled_hopper.py:33-34), one worker process per host core.
# ======================================================================================
_W = {}


def _cpu_init(n, muscles, seed):
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from oracle.pymuscle_oracle import OracleMuscle
    rng = np.random.default_rng(seed + os.getpid())
    _W["m"] = [OracleMuscle(n, "potvin") for _ in range(muscles)]
    _W["e"] = [float(np.float32(67.0 * rng.random())) for _ in range(muscles)]
    _W["n"] = n


def _cpu_run(sim_steps):
    t0 = time.perf_counter()
    ms, es = _W["m"], _W["e"]
    for _ in range(sim_steps):
        for m, e in zip(ms, es):
            m.step(e, 0.02)
    return len(ms) * _W["n"] * sim_steps, time.perf_counter() - t0


class CpuArm:
    def __init__(self, n=N_UNITS, muscles_per_core=16, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.n, self.mpc = n, muscles_per_core
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(n, muscles_per_core, 2017))

    def step(self, sim_steps):
        """All cores advance their muscles `sim_steps` steps; returns (unit-steps, wall seconds)."""
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_run, [sim_steps] * self.cores)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(seconds_target=12.0):
    """Bounded sample: every core loops 16 reference-shaped muscles; ~12 s of wall time."""
    arm = CpuArm()
    arm.step(20)                                              # warm-up
    units, wall = arm.step(100)
    sim_steps = max(100, int(100 * seconds_target / max(wall, 1e-3)))
    units, wall = arm.step(sim_steps)
    arm.close()
    return {"value": units / wall, "unit": METRIC, "cores": arm.cores, "kind": "port",
            "sample": f"{arm.cores} processes x {arm.mpc} PotvinFuglevandMuscle(120) x {sim_steps} steps, "
                      f"per-muscle Python loop (oracle port of the reference), {wall:.1f} s wall"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = CpuArm()
    sim_steps = 100                                           # one bench "step" = a bounded sample of the window
    for _ in range(max(args.warmup, 1)):
        arm.step(sim_steps)
    units = wall = 0.0
    for _ in range(args.steps):
        u, w = arm.step(sim_steps)
        units, wall = units + u, wall + w
    arm.close()
    value = units / wall
    sample = (f"per step: {arm.cores} processes x {arm.mpc} muscles x {sim_steps} simulation steps of the "
              f"65,536 x 120 x 1,000 workload, per-muscle Python loop")
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": arm.cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    return {"workload": f"PotvinFuglevandMuscle({N_UNITS}) x {args.muscles} muscles per GPU, fused window of "
                        f"{args.window} steps, dt=0.02, per-step excitation tensor (constant per muscle)",
            "units_per_muscle": N_UNITS, "muscles_per_gpu": args.muscles, "window_steps": args.window,
            "precision_mode": args.precision, "parallelism": "muscle batch sharded across ranks, no collective",
            "l2_policy": "inputs larger than L2 (excitations 262 MB + state 126 MB per window vs 126 MB L2)"}


# ======================================================================================
# clocks
# ======================================================================================
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU every 50 ms while running."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


BAD_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


# ======================================================================================
# GPU arm
# ======================================================================================
def ncu_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs closest to its GPU (NVML's ideal affinity) so that the pinned
    staging buffers of the end-to-end leg are first-touched on the GPU's own NUMA node.  Several
    ranks streaming through one socket's memory controllers otherwise throttle each other."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()                                  # before CUDA is initialised (fork)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from pymuscle_b200 import MuscleBatch, MuscleSpec
    from pymuscle_b200 import _native as nat

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)                     # pinned host buffers land next to this GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    M, n, K = args.muscles, N_UNITS, args.window
    dt = 0.02
    gen = torch.Generator(device=dev).manual_seed(2017 + rank)
    batch = MuscleBatch(MuscleSpec.potvin(n), M, dev, args.precision, fresh_outputs=False)
    e_m = 67.0 * torch.rand(M, device=dev, generator=gen, dtype=batch.dtype)
    exc = e_m.repeat(K, 1).contiguous()                      # [K, M]: read per step by the kernel
    totals = torch.empty(K, M, dtype=batch.dtype, device=dev)
    unit_steps = M * n * K                                    # per window per rank

    def window():
        batch.step_many(exc, dt, out=totals, units_per_thread=args.units_per_thread)

    # ---- FP32 / FP64 lane-op peak, measured live (not in MEASURED_PEAKS.json) -------------
    lib = nat.lib()
    sink = torch.empty(148 * 8 * 1024, dtype=torch.float64, device=dev)
    probe_kind = 0 if args.precision == "f32" else 1

    def probe(kind, iters):
        ops = C.c_int64()
        best = 1e9
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(lib.mb_peak_probe(kind, 148 * 8, 1024, iters, C.c_void_p(sink.data_ptr()), C.byref(ops),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, a.elapsed_time(b) * 1e-3)
        return ops.value / best

    peak_lane_ops = probe(probe_kind, 4096 if probe_kind == 0 else 1024)

    # ---- timed region ------------------------------------------------------------------------
    def timed(fn, steps, warmup, counter=lambda: 0, join=lambda: None, collective=True):
        """`collective`: every rank runs this region (barriers + max over ranks); False = this rank only."""
        sync = barrier if collective else (lambda: None)
        for _ in range(warmup):
            fn()
        join()
        torch.cuda.synchronize()
        sync()
        sampler = ClockSampler(local_rank)
        sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = counter()
        a.record()
        for _ in range(steps):
            fn()
        join()                                                # side streams (downloads) rejoin the timed stream
        b.record()
        torch.cuda.synchronize()
        sync()
        clocks = sampler.stop()
        clocks["launches"] = counter() - c0
        sec = a.elapsed_time(b) * 1e-3
        if world > 1 and collective:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, clocks

    sec, clocks = timed(window, args.steps, max(args.warmup, 3), lambda: batch.launches)
    if BAD_REASONS & set(clocks["reasons"]):                  # thermal / hw slowdown seen: measure once more
        time.sleep(5)
        sec, clocks = timed(window, args.steps, 3, lambda: batch.launches)
        clocks["remeasured"] = True
    gpu_launches = clocks.pop("launches")
    value = world * unit_steps * args.steps / sec
    per_launch = sec / args.steps

    # ---- end to end: host buffers in, host buffers out, through the public API ---------------
    e2e = None
    if not args.no_e2e:
        exc_h = torch.empty(K, M, dtype=batch.dtype).pin_memory()
        exc_h.copy_(exc)
        out_h = torch.empty(K, M, dtype=batch.dtype).pin_memory()
        e_steps = max(3, min(args.steps, 20))

        def window_host():
            batch.step_many_host(exc_h, dt, out_h, chunk=args.chunk, overlap_calls=True)

        sec_e, _ = timed(window_host, e_steps, 3, join=batch.host_pipeline_join)
        e2e = {"value": world * unit_steps * e_steps / sec_e, "unit": METRIC,
               "h2d_bytes_per_step": exc_h.numel() * exc_h.element_size(),
               "d2h_bytes_per_step": out_h.numel() * out_h.element_size(),
               "steps": e_steps, "ms_per_step": 1e3 * sec_e / e_steps,
               "api": f"MuscleBatch.step_many_host(pinned exc[{K},{M}], dt, pinned totals[{K},{M}], chunk={args.chunk}, overlap_calls=True): "
                      "H2D / fused kernel / D2H of sub-windows on three streams; every window's excitations are "
                      "uploaded and its totals downloaded inside the timed region"}
        del exc_h, out_h

    # ---- the single-step kernel against the HBM roofline (BASELINE config #5 shape) -----------
    k1 = None
    if not args.no_k1 and rank == 0:
        hbm_peak, hbm_src = measured_hbm_peak()
        del exc, totals
        torch.cuda.empty_cache()
        m5 = 131072
        b5 = MuscleBatch(MuscleSpec.standard(527.1), m5, dev, "f32", fresh_outputs=False)
        x5 = [torch.rand(m5, device=dev, generator=gen) for _ in range(4)]
        it = [0]

        def frame():
            b5.step(x5[it[0] & 3], dt)
            it[0] += 1

        sec5, _ = timed(frame, 50, 5, collective=False)     # rank 0 only: no barrier, no all-reduce
        us5 = b5.n_units * 50 / sec5
        bytes_per = K1_BYTES_PER_UNIT_STEP["standard_f32"]
        k1 = {"kernel": "k1_stream<float>", "workload": f"StandardMuscle n=2000 x {m5}, one launch per frame",
              "bound": "hbm", "achieved": us5 * bytes_per / 1e9, "peak": hbm_peak, "unit": "GB/s",
              "frac": us5 * bytes_per / 1e9 / hbm_peak, "peak_source": hbm_src,
              "traffic": ncu_traffic("k1_stream_f32_config5"),
              "algorithmic_bytes_per_launch": b5.n_units * bytes_per,
              "bytes_per_unit_step": bytes_per, "value": us5, "value_unit": METRIC}

    if rank != 0:
        barrier()                                             # wait for rank 0's single-rank extras
        dist.destroy_process_group()
        return 0

    ops = OPS_PER_UNIT_STEP["potvin"]
    achieved = unit_steps * ops / per_launch / 1e12           # canonical FLOP/s: every add / mul / compare / exp = 1
    peak_flops = 2.0 * peak_lane_ops / 1e12                   # an FMA lane-op is 2 FLOP
    pipe = "fp32_pipe" if args.precision == "f32" else "fp64_pipe"
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * per_launch, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args), "clocks": clocks, "gpu_launches": gpu_launches, "e2e": e2e,
        "roofline": {
            "kernel": ("k2w_f32" if args.precision == "f32" else "k2_fused_f64"), "bound": pipe,
            "achieved": achieved, "peak": peak_flops, "unit": "TFLOP/s", "frac": achieved / peak_flops,
            "traffic": ncu_traffic("k2_fused_f32_config2") if args.precision == "f32" else None,
            "ops_per_unit_step": ops,
            "measured_fma_lane_ops_per_s": peak_lane_ops,
            "frac_vs_lane_op_peak": achieved / (peak_lane_ops / 1e12),
            "peak_source": "mb_peak_probe dependent-FMA chains (8 per thread), best of 3, measured in this run, x 2 FLOP "
                           "per FMA; MEASURED_PEAKS.json has no FP32/FP64 pipe figure",
            "note": "achieved = 38 canonical FP operations per unit-step (SURVEY.md A.2 / 8d: the reference's own "
                    "operation list; each add, multiply, compare, select and exp counts 1, no FMA contraction) x "
                    "unit-steps per launch / mean launch time.  The operations are FLOPs, so the peak is the measured "
                    "FMA lane rate x 2 FLOP.  SURVEY 8d divides by the lane rate instead (frac_vs_lane_op_peak); that "
                    "figure exceeds 1 because the kernel contracts FMAs, issues packed FFMA2/FMUL2/FADD2 and carries "
                    "exp(-dur/tau) by recurrence.  profiles/ holds ncu's issue-slot and per-pipe utilisation of the "
                    "same launch"},
        "roofline_single_step": k1,
        "cpu_baseline": cpu,
        "fused_geometry": batch.fused_geometry(K, args.units_per_thread),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--muscles", type=int, default=65536)
    ap.add_argument("--window", type=int, default=1000)
    ap.add_argument("--units-per-thread", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=256, help="sub-window length of the host-buffer pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-k1", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
