"""
History capture: decimated per-unit forces streamed to the host while the simulation keeps running.

The reference's README loop records ``muscle.current_forces`` after every step
(README.md:128-145) and ``pymuscle/vis/potvin_charts.py:9-36`` charts the resulting
``time_by_forces[steps][n]`` list.  On the GPU a per-step read of ``current_forces`` would put a
device-to-host copy and a synchronisation into every step.  ``ForceHistory`` instead lets the fused
kernel write every ``every``-th step's per-unit forces into a device buffer (``forces_every``),
and moves each window's records into a ring of pinned host buffers with an asynchronous copy on
a side stream.  Nothing in ``advance()`` waits for the device; ``collect()`` is the only
synchronisation point and returns the chart-shaped ``[records, rows, n]`` array.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


class ForceHistory:
    """Ring of pinned host buffers fed by asynchronous device-to-host copies.

    batch    a MuscleBatch
    every    keep the forces of steps every-1, 2*every-1, ...
    window   simulation steps per ``advance()`` call (a multiple of ``every``)
    slots    ring depth: windows whose copies may be in flight / unread at once
    """

    def __init__(self, batch, every: int = 1, window: int = 256, slots: int = 4):
        if every < 1 or window < every or window % every:
            raise ValueError("window must be a positive multiple of every")
        if slots < 2:
            raise ValueError("the ring needs at least two slots")
        self.batch, self.every, self.window, self.slots = batch, int(every), int(window), int(slots)
        self.records_per_window = self.window // self.every
        shape = (self.records_per_window, batch.rows, batch.L)
        self._ring = [torch.empty(shape, dtype=batch.dtype).pin_memory() for _ in range(self.slots)]
        self._done = [torch.cuda.Event() for _ in range(self.slots)]
        self._copy_stream = torch.cuda.Stream(batch.device)
        self._pending: List[tuple] = []          # (slot, records) in issue order, not yet handed to the caller
        self._drained: List[np.ndarray] = []     # records taken out of the ring because it was about to wrap
        self._head = 0
        self.steps_recorded = 0

    def advance(self, excitations, step_size: float, steps: Optional[int] = None, **kw) -> torch.Tensor:
        """One fused window (``MuscleBatch.step_many`` with ``forces_every=every``); queues the copy of
        its records into the next ring slot and returns the window's totals (device tensor).  Does not
        synchronise unless the ring is full of unread windows, in which case the oldest one -- whose
        copy finished long ago -- is moved to ordinary host memory first."""
        K = int(steps) if steps is not None else int(excitations.shape[0])
        if K > self.window or K % self.every:
            raise ValueError(f"a window holds at most {self.window} steps in multiples of {self.every}")
        if len(self._pending) == self.slots:
            self._drain_oldest()
        slot = self._head
        self._head = (self._head + 1) % self.slots
        out = self.batch.step_many(excitations, step_size, steps=steps, forces_every=self.every, **kw)
        totals, forces = out[0], out[1]
        n = forces.shape[0]
        cur = torch.cuda.current_stream(self.batch.device)
        self._copy_stream.wait_stream(cur)
        with torch.cuda.stream(self._copy_stream):
            self._ring[slot][:n].copy_(forces, non_blocking=True)
            forces.record_stream(self._copy_stream)          # the allocator must not reuse it before the copy ran
            self._done[slot].record(self._copy_stream)
        self._pending.append((slot, n))
        self.steps_recorded += K
        return totals

    def _drain_oldest(self):
        slot, n = self._pending.pop(0)
        self._done[slot].synchronize()
        self._drained.append(self._ring[slot][:n].numpy().copy())

    def collect(self) -> np.ndarray:
        """Everything recorded since the last ``collect()``, as float64 ``[records, rows, n]`` in step
        order (``time_by_forces`` of README.md:128-145 for ``rows == 1``: ``result[:, 0]``)."""
        while self._pending:
            self._drain_oldest()
        parts, self._drained = self._drained, []
        if not parts:
            return np.zeros((0, self.batch.rows, self.batch.L))
        return np.concatenate(parts, axis=0).astype(np.float64, copy=False)
