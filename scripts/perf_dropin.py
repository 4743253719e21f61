"""Latency of the drop-in scalar API (config #1 shape): StandardMuscle().step(0.6, 0.02) in a Python loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pymuscle_b200 as pymuscle
m = pymuscle.StandardMuscle()
for _ in range(200):
    m.step(0.6, 0.02)
t0 = time.perf_counter()
N = 3000
for _ in range(N):
    out = m.step(0.6, 0.02)
    f = m.current_forces
dt = time.perf_counter() - t0
print(f"StandardMuscle.step + current_forces: {dt / N * 1e6:.1f} us per step ({120 * N / dt / 1e6:.2f} M unit-steps/s); reference on one CPU core: 47 us")
p = pymuscle.PotvinFuglevandMuscle(120)
for _ in range(200):
    p.step(40.0, 0.02)
t0 = time.perf_counter()
for _ in range(N):
    out = p.step(40.0, 0.02)
dt = time.perf_counter() - t0
print(f"PotvinFuglevandMuscle.step (no forces read): {dt / N * 1e6:.1f} us per step")
