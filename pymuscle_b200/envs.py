"""
Vector-environment adapter: the muscles of many agents advanced by one launch per env step.

The reference's environments keep a Python list of muscles per agent and loop over it every
step -- ``outputs = [muscle.step(a[i], dt) for i, muscle in enumerate(self.muscles)]``
(examples/envs/muscled_hopper.py:29-36; examples/envs/pymunk_arm.py:150-179 does the same
for two muscles) -- and a vector env repeats that per agent.  ``MuscleGroupEnv`` holds the same
muscles for ``agents`` copies on the GPU: ``step(actions)`` is one streaming-kernel launch that
returns every muscle's output and, as the proprioceptive observation the library offers
(README "fatigue" signal, muscle.py:248-256), every muscle's peripheral fatigue.  With
``rest_lengths`` set, outputs are scaled by the Hill-type force-length and force-velocity
factors (pymuscle/hill_type.py:15-66) of the muscle lengths the physics reports; the curves are
evaluated inside the same launch (the kernel's per-muscle epilogue, ``mb_step_env_hill``).
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch

from .engine import MuscleBatch
from .tables import MuscleSpec


class MuscleGroupEnv:
    """``agents`` copies of a group of muscles.

    specs         the group (e.g. 4 StandardMuscles of a hopper leg, 30 of an arm)
    agents        number of environment copies
    rest_lengths  optional [S] (or [agents, S]) resting lengths: enables the Hill-type scaling
    """

    def __init__(self, specs: Union[MuscleSpec, Sequence[MuscleSpec]], agents: int, device=None,
                 precision: str = "f32", rest_lengths: Optional[torch.Tensor] = None):
        self.specs = [specs] if isinstance(specs, MuscleSpec) else list(specs)
        self.batch = MuscleBatch(self.specs if len(self.specs) > 1 else self.specs[0], agents, device,
                                 precision, fresh_outputs=False)
        self.agents, self.S = int(agents), len(self.specs)
        self.device, self.dtype = self.batch.device, self.batch.dtype
        self.rest = None
        if rest_lengths is not None:
            self.rest = torch.as_tensor(rest_lengths, dtype=self.dtype, device=self.device).expand(self.agents, self.S)
        self._prev_len = None

    def reset(self):
        self.batch.reset()
        self._prev_len = None

    def step(self, actions: torch.Tensor, step_size: float, lengths: Optional[torch.Tensor] = None):
        """actions [agents, S] in the muscles' input units (0..1 for StandardMuscle);
        lengths [agents, S] current muscle lengths (needs ``rest_lengths``).
        Returns (outputs [agents, S], fatigue [agents, S] float64)."""
        a = actions.to(device=self.device, dtype=self.dtype).reshape(self.agents, self.S)
        hill = None
        if lengths is not None:
            if self.rest is None:
                raise ValueError("lengths given but the environment has no rest_lengths")
            cur = lengths.to(device=self.device, dtype=self.dtype).reshape(self.agents, self.S)
            prev = cur if self._prev_len is None else self._prev_len      # first step: zero velocity
            hill = (self.rest, cur, prev)
            self._prev_len = cur.clone()
        out, fat = self.batch.step_env(a, step_size, want_fatigue=True, hill=hill)
        return out.view(self.agents, self.S), fat.view(self.agents, self.S)

    def graphed_step(self, step_size: float, with_gain: bool = False):
        """Capture one env step (the ``mb_step_env`` launch and, with ``with_gain``, the multiplication by a
        caller-supplied per-muscle gain) in a CUDA graph and return ``f(actions, gain=None) -> (outputs,
        fatigue)`` that replays it.  Small environments (a hopper's 4 muscles x a few hundred agents) are
        launch-bound: a replay costs one ``cudaGraphLaunch`` instead of the Python + ctypes + allocator path.
        The returned tensors are the graph's static outputs: they are overwritten by the next call."""
        dev = self.device
        a_in = torch.zeros(self.agents, self.S, dtype=self.dtype, device=dev)
        g_in = torch.ones(self.agents, self.S, dtype=self.dtype, device=dev) if with_gain else None
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        state = self.batch.state_dict()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # warm-up off the capture: function attributes, occupancy cache
            for _ in range(2):
                self.batch.step_env(a_in, step_size, gain=g_in, want_fatigue=True)
        cur.wait_stream(side)
        self.batch.load_state_dict(state)                   # the warm-up steps must not count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out, fat = self.batch.step_env(a_in, step_size, gain=g_in, want_fatigue=True)
        self.batch.load_state_dict(state)                   # (capture does not execute, but keep the contract obvious)
        out, fat = out.view(self.agents, self.S), fat.view(self.agents, self.S)

        def step(actions: torch.Tensor, gain: Optional[torch.Tensor] = None):
            a_in.copy_(actions.reshape(self.agents, self.S), non_blocking=True)
            if with_gain:
                if gain is None:
                    raise ValueError("this graph was captured with a gain input")
                g_in.copy_(gain.reshape(self.agents, self.S), non_blocking=True)
            graph.replay()
            return out, fat

        step.graph = graph                                   # keep the graph (and its memory pool) alive with the callable
        return step
