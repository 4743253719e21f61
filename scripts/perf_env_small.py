"""Launch-bound environments: eager env step vs CUDA-graph replay (development aid).
Hopper-like group: 4 StandardMuscles x `agents` copies."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymuscle_b200 import MuscleGroupEnv, MuscleSpec
dev = torch.device("cuda:0")
specs = [MuscleSpec.standard(f) for f in (20.0, 32.0, 45.0, 32.0)]
for agents in (1, 256, 4096, 65536):
    env = MuscleGroupEnv(specs, agents, dev, "f32")
    a = torch.rand(agents, 4, device=dev)
    f = env.graphed_step(0.02)
    res = []
    for name, fn in (("eager", lambda: env.step(a, 0.02)), ("graph", lambda: f(a))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 2000
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        res.append(f"{name} {dt * 1e6:6.1f} us/step ({agents / dt / 1e6:8.2f} M agent-steps/s)")
    print(f"hopper group x {agents:6d} agents ({env.batch.n_units} units): " + " | ".join(res), flush=True)
