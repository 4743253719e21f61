#!/usr/bin/env python
"""
bench.py -- motor-unit-steps/s of the PyMuscle hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path, BASELINE config #2
    python bench.py --config {2,3,4,5} [--strong]                   # the other BASELINE configs / strong scaling
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Default workload (BASELINE.json configs[1]): 65,536 independent PotvinFuglevandMuscle(120) per GPU,
FP32 mode, one "step" = one fused window of 1,000 simulation steps (dt = 0.02) = 7.86e9
motor-unit-steps.  `value` reads a [1000, 65536] device-resident excitation tensor per step (as a
time-varying signal would be); `e2e` is config #2 literally -- a CONSTANT excitation per muscle -- through
the public host-buffer API (MuscleBatch.step_many_host): the [65536] excitation vector is uploaded from
pinned memory and totals are downloaded to pinned memory inside the timed region of every window.
Weak scaling: every rank owns its own muscles; muscles are independent, so there is no collective on
the data path.  At N > 1 the line also carries the cost of the optional gather of totals (config #4
shape: FP64, 32,768 muscles per rank) done three ways: none / NCCL all_gather / fused into the kernel.

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
K1_BYTES_PER_UNIT_STEP = {"standard_f32": 20, "potvin_f32": 36, "standard_f64": 24, "potvin_f64": 40}
ARM_FORCES = [round(68.0 * (824.0 / 68.0) ** (i / 29.0), 1) for i in range(30)]     # config #3


# ======================================================================================
# CPU arm: the reference's algorithm, one Python object per muscle, looped over muscles
# (examples/envs/muscled_hopper.py:33-34), one worker process per host core.
# ======================================================================================
_W = {}


def find_reference():
    """The unmodified reference package if it is importable here: baseline/_ref (an install that travels
    with the repo) or /root/reference (the build container only).  None on the GPU box."""
    for base in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(base, "pymuscle", "__init__.py")):
            return base
    return None


def _cpu_init(n, muscles, seed, ref_path):
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    rng = np.random.default_rng(seed + os.getpid())
    if ref_path:
        sys.path.insert(0, ref_path)
        from pymuscle import PotvinFuglevandMuscle
        _W["m"] = [PotvinFuglevandMuscle(n) for _ in range(muscles)]
    else:
        from oracle.pymuscle_oracle import OracleMuscle
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
        self.ref_path = find_reference()
        self.kind = "reference" if self.ref_path else "port"
        self.what = (f"the unmodified reference imported from {self.ref_path}" if self.ref_path else
                     "oracle port of the reference (oracle/pymuscle_oracle.py: bit-identical to it, "
                     "tests/golden/make_golden.py --check; the Python reference does not travel to the GPU box)")
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init,
                                                initargs=(n, muscles_per_core, 2017, self.ref_path))

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
    return {"value": units / wall, "unit": METRIC, "cores": arm.cores, "kind": arm.kind,
            "sample": f"{arm.cores} processes x {arm.mpc} PotvinFuglevandMuscle(120) x {sim_steps} steps, "
                      f"per-muscle Python loop, {arm.what}, {wall:.1f} s wall"}


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
              f"65,536 x 120 x 1,000 workload, per-muscle Python loop, {arm.what}")
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    """Shared by both arms (the arithmetic type of each arm is its line's `dtype`)."""
    if args.config == 2:
        per = "in total (strong scaling)" if args.strong else "per GPU"
        return {"workload": f"BASELINE configs[1]: PotvinFuglevandMuscle({N_UNITS}) x {args.muscles} muscles {per}, fused window "
                            f"of {args.window} steps, dt=0.02, constant excitation per muscle (device-timed value reads it "
                            f"from a [{args.window}, muscles] tensor per step; e2e uploads the [muscles] vector per window)",
                "units_per_muscle": N_UNITS, "muscles_per_gpu": args.muscles, "window_steps": args.window,
                "parallelism": "muscle batch sharded across ranks, no collective on the data path",
                "l2_policy": "inputs larger than L2 (excitations 262 MB + state 126 MB per window vs 126 MB L2)"}
    if args.config == 3:
        return {"workload": "BASELINE configs[2]: arm of 30 StandardMuscles (35,032 units) x 4,096 agents per GPU, one env "
                            "step per launch (outputs + peripheral fatigue), per-step random excitation",
                "units_per_agent": 35032, "agents_per_gpu": 4096, "parallelism": "agents sharded across ranks, no collective",
                "l2_policy": "state 1.15 GB per step, larger than L2"}
    if args.config == 4:
        return {"workload": "BASELINE configs[3]: PotvinFuglevandMuscle(120) x 32,768 per GPU in FP64, ramp-hold-rest "
                            "excitation, fused windows of 1,000 steps, totals of every window gathered on every rank "
                            "(fused into the kernel's stores when N > 1)",
                "units_per_muscle": N_UNITS, "muscles_per_gpu": 32768, "window_steps": 1000,
                "parallelism": "muscle batch sharded across ranks; gather of totals = peer stores over NVLink",
                "l2_policy": "excitations 262 MB per window, larger than L2"}
    return {"workload": "BASELINE configs[4]: StandardMuscle(527.1 N) = 2,000 units x 131,072 instances per GPU, one launch "
                        "per frame, per-step random excitation",
            "units_per_muscle": 2000, "muscles_per_gpu": 131072, "parallelism": "muscle batch sharded across ranks, no collective",
            "l2_policy": "state 2.1 GB per step, larger than L2"}


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
def profile_facts(key):
    """Figures read from the committed ncu captures (profiles/traffic.json: DRAM bytes per launch;
    profiles/pipes.json: issue-slot and per-pipe utilisation), or None."""
    out = {}
    for fname in ("traffic.json", "pipes.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fname)) as f:
                out[fname[:-5]] = json.load(f).get(key)
        except Exception:
            out[fname[:-5]] = None
    return out


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


class Harness:
    """Process-group plumbing and the timed region shared by all configs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            if not os.environ.get("MB_BENCH_NO_BIND"):
                bind_to_gpu_numa_node(self.local_rank)        # pinned host buffers land next to this GPU
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def timed(self, fn, steps, warmup, counter=lambda: 0, join=lambda: None, collective=True):
        """`collective`: every rank runs this region (barriers + max over ranks); False = this rank only."""
        torch = self.torch
        sync = self.barrier if collective else (lambda: None)
        for _ in range(warmup):
            fn()
        join()
        # the sampler's NVML start-up takes milliseconds and differs between ranks: do it BEFORE the barrier, or a
        # region with a cross-rank dependency (a gather) is charged the skew between the ranks' start times
        sampler = ClockSampler(self.local_rank)
        if os.environ.get("MB_BENCH_NO_SAMPLER"):
            sampler.h = None
        sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        sync()
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
        if self.world > 1 and collective:
            t = torch.tensor([sec], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, clocks

    def timed_checked(self, fn, steps, warmup, counter=lambda: 0, join=lambda: None):
        sec, clocks = self.timed(fn, steps, max(warmup, 3), counter, join)
        if BAD_REASONS & set(clocks["reasons"]):              # thermal / hw slowdown seen: measure once more
            time.sleep(5)
            sec, clocks = self.timed(fn, steps, 3, counter, join)
            clocks["remeasured"] = True
        return sec, clocks

    def finish(self, line):
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.barrier()
            self.dist.destroy_process_group()
        return 0


def lane_op_peak(h, precision):
    """FP32 / FP64 FMA lane rate measured live with dependent-chain microkernels (MEASURED_PEAKS.json has none)."""
    import ctypes as C
    from pymuscle_b200 import _native as nat
    torch = h.torch
    lib = nat.lib()
    sink = torch.empty(148 * 8 * 1024, dtype=torch.float64, device=h.dev)
    kind = 0 if precision == "f32" else 1
    iters = 4096 if kind == 0 else 1024
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


def host_copy_probe(h, nbytes=262_144_000):
    """Pinned-memory copy rates of THIS box, all ranks at once: H2D only, D2H only, both directions.
    GB/s per rank (max time over ranks).  States the host limit of the end-to-end legs."""
    torch = h.torch
    n = nbytes // 4
    hp1, hp2 = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
    d1, d2 = torch.empty(n, dtype=torch.float32, device=h.dev), torch.empty(n, dtype=torch.float32, device=h.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(hp1, non_blocking=True)
        with torch.cuda.stream(s2):
            hp2.copy_(d2, non_blocking=True)

    def join():
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    out = {"bytes_per_copy": n * 4, "ranks_at_once": h.world}
    for name, fn, j in (("h2d_gbs_per_rank", lambda: d1.copy_(hp1, non_blocking=True), lambda: None),
                        ("d2h_gbs_per_rank", lambda: hp2.copy_(d2, non_blocking=True), lambda: None),
                        ("both_gbs_per_rank_per_direction", both, join)):
        sec, _ = h.timed(fn, 4, 1, join=j)
        out[name] = round(n * 4 * 4 / sec / 1e9, 2)
    return out


def k1_roofline(h, gen):
    """The single-step kernel against the HBM roofline (BASELINE config #5 shape), rank 0 only."""
    from pymuscle_b200 import MuscleBatch, MuscleSpec
    torch = h.torch
    hbm_peak, hbm_src = measured_hbm_peak()
    m5 = 131072
    b5 = MuscleBatch(MuscleSpec.standard(527.1), m5, h.dev, "f32", fresh_outputs=False)
    x5 = [torch.rand(m5, device=h.dev, generator=gen) for _ in range(4)]
    it = [0]

    def frame():
        b5.step(x5[it[0] & 3], 0.02)
        it[0] += 1

    sec5, _ = h.timed(frame, 50, 5, collective=False)        # rank 0 only: no barrier, no all-reduce
    us5 = b5.n_units * 50 / sec5
    bytes_per = K1_BYTES_PER_UNIT_STEP["standard_f32"]
    facts = profile_facts("k1_stream_f32_config5")
    return {"kernel": "k1_stream<float>", "workload": f"StandardMuscle n=2000 x {m5}, one launch per frame",
            "bound": "hbm", "achieved": us5 * bytes_per / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": us5 * bytes_per / 1e9 / hbm_peak, "peak_source": hbm_src,
            "traffic": facts["traffic"], "ncu": facts["pipes"],
            "algorithmic_bytes_per_launch": b5.n_units * bytes_per,
            "bytes_per_unit_step": bytes_per, "value": us5, "value_unit": METRIC}


def gather_cost(h):
    """Config #4's exchange step at N > 1: one fused FP64 window of 1,000 steps x 32,768 muscles per rank with the
    totals of all ranks needed on every rank -- not gathered / NCCL all_gather after the kernel / gather fused
    into the kernel's epilogue (peer stores over NVLink through symmetric memory)."""
    from pymuscle_b200 import MuscleSpec, ShardedMuscleBatch
    torch = h.torch
    per_rank, K = 32768, 1000
    res = {"workload": f"PotvinFuglevandMuscle(120) x {per_rank} per rank x {K} steps, FP64 mode, totals of every rank on every rank",
           "unit": "ms per window (max over ranks)"}
    try:
        def fresh():                                          # every leg starts from rested muscles: the kernel's time
            return ShardedMuscleBatch(MuscleSpec.potvin(N_UNITS), per_rank * h.world, h.dev, "f64",   # drifts with the state
                                      fresh_outputs=False)
        gen = torch.Generator(device=h.dev).manual_seed(4 + h.rank)
        e = (67.0 * (0.5 + 0.5 * torch.rand(per_rank, device=h.dev, generator=gen, dtype=torch.float64))).repeat(K, 1).contiguous()
        sink = torch.zeros((), dtype=torch.float64, device=h.dev)
        peer_of = (h.rank + 1) % h.world                      # consume the totals that came from the next rank

        def run(sb, mode):
            if mode == "peer":
                sink.add_(sb.step_many_peer_gather(e, 0.02, dense=False)[peer_of][-1].sum())
            elif mode == "nccl":
                sink.add_(sb.step_many(e, 0.02, gather=True)[-1].sum())
            else:
                sink.add_(sb.step_many(e, 0.02)[-1].sum())

        for mode in ("none", "nccl"):
            sb = fresh()
            sec, clk = h.timed(lambda: run(sb, mode), 5, 2)
            res[mode + "_ms"] = round(1e3 * sec / 5, 3)
            res[mode + "_sm_mhz"] = clk.get("sm_mhz")
            del sb
        try:
            sb = fresh()
            sb.enable_peer_gather(K)
            res["peer_stores"] = "nvlink multicast (multimem.st)" if sb._pg.get("multicast") else "one store per peer"
            sec, clk = h.timed(lambda: run(sb, "peer"), 5, 2)
            res["peer_ms"] = round(1e3 * sec / 5, 3)
            res["peer_sm_mhz"] = clk.get("sm_mhz")
            res["peer_over_none"] = round(res["peer_ms"] / res["none_ms"], 4)
        except Exception as exc:                               # symmetric memory unavailable on this box
            res["peer_error"] = f"{type(exc).__name__}: {exc}"[:300]
        res["unit_steps_per_s_none"] = per_rank * h.world * N_UNITS * K / (res["none_ms"] * 1e-3)
    except Exception as exc:
        res["error"] = f"{type(exc).__name__}: {exc}"[:300]
    return res


def run_config2(h, args, cpu):
    torch = h.torch
    from pymuscle_b200 import MuscleBatch, MuscleSpec
    world, rank, dev = h.world, h.rank, h.dev
    M_total = args.muscles if args.strong else args.muscles * world
    M = M_total // world if args.strong else args.muscles
    n, K, dt = N_UNITS, args.window, 0.02
    gen = torch.Generator(device=dev).manual_seed(2017 + rank)
    batch = MuscleBatch(MuscleSpec.potvin(n), M, dev, args.precision, fresh_outputs=False)
    e_m = 67.0 * torch.rand(M, device=dev, generator=gen, dtype=batch.dtype)
    exc = e_m.repeat(K, 1).contiguous()                      # [K, M]: read per step by the kernel
    totals = torch.empty(K, M, dtype=batch.dtype, device=dev)
    unit_steps = M * n * K                                    # per window per rank

    def window():
        batch.step_many(exc, dt, out=totals, units_per_thread=args.units_per_thread)

    peak_lane_ops = lane_op_peak(h, args.precision)
    sec, clocks = h.timed_checked(window, args.steps, args.warmup, lambda: batch.launches)
    gpu_launches = clocks.pop("launches")
    value = world * unit_steps * args.steps / sec
    per_launch = sec / args.steps

    # ---- end to end: host buffers in, host buffers out, through the public API ---------------
    e2e = None
    if not args.no_e2e:
        esz = exc.element_size()
        e_steps = max(3, min(args.steps, 20))
        const_h = torch.empty(M, dtype=batch.dtype).pin_memory()
        const_h.copy_(e_m)
        tev = args.totals_every
        out_dec = torch.empty(K // tev, M, dtype=batch.dtype).pin_memory()

        def leg(fn, h2d, d2h, api):
            s, _ = h.timed(fn, e_steps, 3, join=batch.host_pipeline_join)
            return {"value": world * unit_steps * e_steps / s, "unit": METRIC, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e_steps, "ms_per_step": 1e3 * s / e_steps, "api": api}

        # headline: config #2 as stated -- constant excitation per muscle, totals sampled every `tev` steps
        # (one sub-window per window: with 256 KB up and 26 MB down there is nothing to pipeline inside a window; the
        #  download of window i overlaps the kernel of window i + 1)
        e2e = leg(lambda: batch.step_many_host(const_h, dt, out_dec, chunk=K, steps=K, totals_every=tev,
                                               overlap_calls=True),
                  M * esz, out_dec.numel() * esz,
                  f"MuscleBatch.step_many_host(pinned exc[{M}] (constant over the window), dt, pinned totals[{K // tev},{M}], "
                  f"steps={K}, totals_every={tev}, chunk={K}, overlap_calls=True): the excitation vector is uploaded and "
                  f"every {tev}-th step's totals are downloaded inside the timed region of every window")
        del out_dec
        variants = {}
        if not args.no_e2e_variants:
            out_h = torch.empty(K, M, dtype=batch.dtype).pin_memory()
            variants["constant_exc_all_totals"] = leg(
                lambda: batch.step_many_host(const_h, dt, out_h, chunk=args.chunk, steps=K, overlap_calls=True),
                M * esz, out_h.numel() * esz, f"constant excitation [{M}] in, ALL per-step totals [{K},{M}] out")
            exc_h = torch.empty(K, M, dtype=batch.dtype).pin_memory()
            exc_h.copy_(exc)
            variants["per_step_exc_all_totals"] = leg(
                lambda: batch.step_many_host(exc_h, dt, out_h, chunk=args.chunk, overlap_calls=True),
                exc_h.numel() * esz, out_h.numel() * esz,
                f"per-step excitations [{K},{M}] in, ALL per-step totals [{K},{M}] out (the round-1 e2e form: bound by "
                f"the host's copy rate when several GPUs share it, see host_copy_probe)")
            del exc_h, out_h
            variants["host_copy_probe"] = host_copy_probe(h)
        e2e["variants"] = variants

    # ---- the optional exchange step, three ways (N > 1) ----------------------------------------------
    gather = gather_cost(h) if (world > 1 and not args.no_gather) else None

    # ---- the single-step kernel against the HBM roofline (BASELINE config #5 shape) -----------
    k1 = None
    if not args.no_k1 and rank == 0:
        del exc, totals
        torch.cuda.empty_cache()
        k1 = k1_roofline(h, gen)

    ops = OPS_PER_UNIT_STEP["potvin"]
    achieved = unit_steps * ops / per_launch / 1e12           # canonical FLOP/s: every add / mul / compare / exp = 1
    peak_flops = 2.0 * peak_lane_ops / 1e12                   # an FMA lane-op is 2 FLOP
    pipe = "fp32_pipe" if args.precision == "f32" else "fp64_pipe"
    kname = "k2w_f32" if args.precision == "f32" else "k2w_f64"
    facts = profile_facts(kname + "_config2")
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * per_launch, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args), "clocks": clocks, "gpu_launches": gpu_launches, "e2e": e2e,
        "roofline": {
            "kernel": kname, "bound": pipe,
            "achieved": achieved, "peak": peak_flops, "unit": "TFLOP/s", "frac": achieved / peak_flops,
            "traffic": facts["traffic"], "ncu": facts["pipes"],
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
                    "exp(-dur/tau) by recurrence.  `ncu` = issue-slot and per-pipe utilisation of the same launch from "
                    "the committed capture in profiles/",
            "k1": k1,
            "gather_config4": gather,
        },
        "cpu_baseline": cpu,
        "fused_geometry": batch.fused_geometry(K, args.units_per_thread),
    }
    return h.finish(line)


def run_config3(h, args, cpu):
    """Arm x agents: one env step (outputs + fatigue readout) per launch; HBM-bound."""
    torch = h.torch
    from pymuscle_b200 import MuscleGroupEnv, MuscleSpec
    agents = 4096 // h.world if args.strong else 4096
    env = MuscleGroupEnv([MuscleSpec.standard(f) for f in ARM_FORCES], agents, h.dev, args.precision)
    gen = torch.Generator(device=h.dev).manual_seed(2017 + h.rank)
    acts = [torch.rand(agents, 30, device=h.dev, generator=gen, dtype=env.dtype) for _ in range(4)]
    it = [0]

    def frame():
        env.step(acts[it[0] & 3], 0.02)
        it[0] += 1

    sec, clocks = h.timed_checked(frame, args.steps, args.warmup, lambda: env.batch.launches)
    launches = clocks.pop("launches")
    units = env.batch.n_units
    value = h.world * units * args.steps / sec
    # end to end: actions from pinned host memory, outputs and fatigue back to pinned host memory, every step
    a_h = [a.cpu().pin_memory() for a in acts]
    o_h = torch.empty(agents, 30, dtype=env.dtype).pin_memory()
    f_h = torch.empty(agents, 30, dtype=torch.float64).pin_memory()

    def frame_host():
        o, f = env.step(a_h[it[0] & 3].to(h.dev, non_blocking=True), 0.02)
        o_h.copy_(o, non_blocking=True)
        f_h.copy_(f, non_blocking=True)
        it[0] += 1

    sec_e, _ = h.timed(frame_host, args.steps, 3)
    bpu = K1_BYTES_PER_UNIT_STEP["standard_" + args.precision]
    peak, src = measured_hbm_peak()
    ach = units * args.steps / sec * bpu / 1e9
    kkey = "k1_flat_f32_config3_env" if args.precision == "f32" else "k1_stream_f64_config3_env"
    facts = profile_facts(kkey)
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": h.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": h.world * units * args.steps / sec_e, "unit": METRIC,
                    "h2d_bytes_per_step": a_h[0].numel() * a_h[0].element_size(),
                    "d2h_bytes_per_step": o_h.numel() * o_h.element_size() + f_h.numel() * 8,
                    "api": "MuscleGroupEnv.step(actions from pinned host memory) -> outputs, fatigue copied to pinned host memory"},
            "roofline": {"kernel": ("k1_flat<float>" if args.precision == "f32" else "k1_stream<double>") + " + fatigue epilogue", "bound": "hbm", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "peak_source": src, "bytes_per_unit_step": bpu,
                         "traffic": facts["traffic"], "ncu": facts["pipes"]},
            "cpu_baseline": cpu}
    return h.finish(line)


def run_config4(h, args, cpu):
    """FP64 ramp-hold-rest protocol in fused windows with the totals of all ranks on every rank."""
    torch = h.torch
    from pymuscle_b200 import MuscleSpec, ShardedMuscleBatch
    per_rank = 32768 // h.world if args.strong else 32768
    K, W, dt = 10000, 1000, 0.02
    sb = ShardedMuscleBatch(MuscleSpec.potvin(N_UNITS), per_rank * h.world, h.dev, "f64", fresh_outputs=False)
    mode = "none"
    if h.world > 1:
        try:
            sb.enable_peer_gather(W)
            mode = "peer"
        except Exception:
            mode = "nccl"
    g = torch.Generator(device=h.dev).manual_seed(4 + h.rank)
    emax = 67.0 * (0.5 + 0.5 * torch.rand(per_rank, device=h.dev, generator=g, dtype=torch.float64))
    k = torch.arange(K, device=h.dev, dtype=torch.float64)
    shape = torch.where(k < 2000, k / 2000.0, torch.where(k < 6000, torch.ones_like(k), torch.zeros_like(k)))
    excs = [(shape[w:w + W, None] * emax[None, :]).contiguous() for w in range(0, K, W)]
    acc = torch.zeros((), dtype=torch.float64, device=h.dev)
    wi = [0]

    def window():
        exc = excs[wi[0] % len(excs)]
        if wi[0] % len(excs) == 0:
            sb.local.reset()
        wi[0] += 1
        if mode == "peer":
            acc.add_(sum(v[-1].sum() for v in sb.step_many_peer_gather(exc, dt, dense=False)))
        elif mode == "nccl":
            acc.add_(sb.step_many(exc, dt, gather=True)[-1].sum())
        else:
            acc.add_(sb.step_many(exc, dt)[-1].sum())

    steps = max(args.steps // 10 * 10, 10)                    # whole protocols
    sec, clocks = h.timed_checked(window, steps, 10, lambda: sb.local.launches)
    launches = clocks.pop("launches")
    unit_steps = per_rank * N_UNITS * W
    value = h.world * unit_steps * steps / sec
    peak = lane_op_peak(h, "f64")
    ach = unit_steps * OPS_PER_UNIT_STEP["potvin"] * steps / sec / 1e12
    facts = profile_facts("k2w_f64_config4")        # (the same kernel as config #2 in FP64 mode; captured on that shape)
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": h.world, "steps": steps, "warmup": 10,
            "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(workload_config(args), gather=mode),
            "clocks": clocks, "gpu_launches": launches, "e2e": None,
            "roofline": {"kernel": "k2w_f64", "bound": "fp64_pipe", "achieved": ach, "peak": 2 * peak / 1e12, "unit": "TFLOP/s",
                         "frac": ach / (2 * peak / 1e12), "ops_per_unit_step": OPS_PER_UNIT_STEP["potvin"],
                         "traffic": facts["traffic"], "ncu": facts["pipes"]},
            "cpu_baseline": cpu}
    return h.finish(line)


def run_config5(h, args, cpu):
    """StandardMuscle(2,000 units) x 131,072: one launch per frame; HBM-bound."""
    torch = h.torch
    from pymuscle_b200 import MuscleBatch, MuscleSpec
    M = 131072 // h.world if args.strong else 131072
    if args.precision == "f64":
        M //= 2
    b = MuscleBatch(MuscleSpec.standard(527.1), M, h.dev, args.precision, fresh_outputs=False)
    gen = torch.Generator(device=h.dev).manual_seed(2017 + h.rank)
    xs = [torch.rand(M, device=h.dev, generator=gen, dtype=b.dtype) for _ in range(4)]
    it = [0]

    def frame():
        b.step(xs[it[0] & 3], 0.02)
        it[0] += 1

    sec, clocks = h.timed_checked(frame, args.steps, args.warmup, lambda: b.launches)
    launches = clocks.pop("launches")
    value = h.world * b.n_units * args.steps / sec
    x_h = [x.cpu().pin_memory() for x in xs]
    o_h = torch.empty(M, dtype=b.dtype).pin_memory()

    def frame_host():
        o_h.copy_(b.step(x_h[it[0] & 3].to(h.dev, non_blocking=True), 0.02), non_blocking=True)
        it[0] += 1

    sec_e, _ = h.timed(frame_host, args.steps, 3)
    bpu = K1_BYTES_PER_UNIT_STEP["standard_" + args.precision]
    peak, src = measured_hbm_peak()
    ach = b.n_units * args.steps / sec * bpu / 1e9
    facts = profile_facts(f"k1_stream_{args.precision}_config5")
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": h.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": h.world * b.n_units * args.steps / sec_e, "unit": METRIC, "h2d_bytes_per_step": M * b.cpf.element_size() // 2
                    if False else M * xs[0].element_size(), "d2h_bytes_per_step": M * xs[0].element_size(),
                    "api": "MuscleBatch.step(excitations from pinned host memory) -> totals copied to pinned host memory"},
            "roofline": {"kernel": "k1_stream", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_source": src, "bytes_per_unit_step": bpu, "traffic": facts["traffic"], "ncu": facts["pipes"]},
            "cpu_baseline": cpu}
    return h.finish(line)


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()                                  # before CUDA is initialised (fork)
    h = Harness(args)
    return {2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](h, args, cpu)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config number (1-based)")
    ap.add_argument("--strong", action="store_true", help="fixed total work split over the ranks (default: weak scaling)")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--muscles", type=int, default=65536)
    ap.add_argument("--window", type=int, default=1000)
    ap.add_argument("--units-per-thread", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=250, help="sub-window length of the host-buffer pipeline")
    ap.add_argument("--totals-every", type=int, default=10,
                    help="e2e: download every t-th step's totals (5 Hz of the 50 Hz simulation by default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true")
    ap.add_argument("--no-k1", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        import socket
        import subprocess
        with socket.socket() as sock:                         # a free port: back-to-back launches must not collide
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
