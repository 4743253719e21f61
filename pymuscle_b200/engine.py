"""
MuscleBatch: N independent muscles advanced together on one B200.

This is the batched entry point the north star adds to the reference API: the
same ``step(excitation, step_size)`` / ``current_forces`` contract as
``Muscle`` (pymuscle/muscle.py:61-92), but over a whole batch and with torch
CUDA tensors.  All arithmetic happens in libmuscle_b200.so (hand-written
sm_100a kernels behind the C ABI in include/muscle_b200.h); torch only owns
the device memory and the stream.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _native as nat
from .tables import MuscleSpec, PackedRow, pack_row, pack_rows_per_instance

_PRECISIONS = {"f32": nat.MB_F32, "fp32": nat.MB_F32, "float32": nat.MB_F32, torch.float32: nat.MB_F32,
               "f64": nat.MB_F64, "fp64": nat.MB_F64, "float64": nat.MB_F64, torch.float64: nat.MB_F64}


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class MuscleBatch:
    """``batch`` rows, each holding the muscles described by ``spec``.

    spec       one :class:`MuscleSpec` (uniform batch: ``batch`` identical muscles) or a
               sequence of them (each row is that group of muscles, e.g. the 30 muscles of an arm)
    batch      number of rows (independent copies)
    precision  "f32" (float arithmetic, float64 state; rtol 1e-4 over 10,000 steps) or
               "f64" (float64 throughout; rtol 1e-9).  Recruitment masks are exact in both.
    fresh_outputs  True (default): every step returns newly allocated tensors, like the
               reference's fresh ``current_forces`` array (fibers:258); False: reuse buffers
    row_specs  optional: one MuscleSpec PER ROW (same unit count, see
               ``tables.pack_rows_per_instance``) -- per-instance hyper-parameters for domain
               randomisation; ``spec`` is then ignored except as row 0's template
    """

    def __init__(self, spec: Union[MuscleSpec, Sequence[MuscleSpec]], batch: int,
                 device: Union[str, torch.device, None] = None, precision="f32",
                 fresh_outputs: bool = True, row_specs: Optional[Sequence[MuscleSpec]] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("pymuscle_b200 needs a CUDA device: there is no CPU fallback")
        nat.lib()
        self._ctor = dict(spec=spec, batch=batch, device=device, precision=precision, fresh_outputs=fresh_outputs,
                          row_specs=row_specs)
        self.uniform = isinstance(spec, MuscleSpec)
        specs = [spec] if self.uniform else list(spec)
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be 'f32' or 'f64', got {precision!r}")
        self.precision = _PRECISIONS[precision]
        self.dtype = torch.float32 if self.precision == nat.MB_F32 else torch.float64
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("pymuscle_b200 runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rows = int(batch)
        if self.rows <= 0:
            raise ValueError("batch must be positive")
        self.fresh_outputs = bool(fresh_outputs)

        self.per_instance = row_specs is not None
        if self.per_instance:
            row_specs = list(row_specs)
            if len(row_specs) != self.rows:
                raise ValueError(f"row_specs has {len(row_specs)} entries for a batch of {self.rows}")
            self.uniform, specs = True, [row_specs[0]]
            self.packed: PackedRow = pack_rows_per_instance(row_specs, self.precision)
        else:
            self.packed = pack_row(specs, self.precision)
        self.specs = specs
        self.cfg = self.packed.cfg
        self.S, self.L = self.packed.S, self.packed.L
        with torch.cuda.device(self.device):
            self._params = torch.from_numpy(self.packed.blob).to(self.device)
            self._seg_off = torch.from_numpy(self.packed.seg_off).to(self.device) if self.S > 1 else None
            self._seg_div = torch.from_numpy(self.packed.seg_div).to(self.device) if self.S > 1 else None
            self._ptf = torch.from_numpy(self.packed.ptf).to(self.device)
            if self.per_instance:
                self._ptf = self._ptf.reshape(self.rows, self.L)
                self.cpf = self._ptf.clone()                                  # fibers:63, each row its own peaks
            else:
                self.cpf = self._ptf.repeat(self.rows, 1).contiguous()        # fibers:63
            self.dur = torch.zeros(self.rows, self.L, dtype=torch.float64, device=self.device)  # pool:79
            self._forces = torch.zeros(self.rows, self.L, dtype=self.dtype, device=self.device)  # fibers:90
            self._totals = torch.zeros(self.rows, self.S, dtype=self.dtype, device=self.device)
        self.launches = 0          # kernels launched so far (bench.py reports it)
        # ctypes views that never change: built once (a single-step launch is a few microseconds of GPU time,
        # so per-call marshalling of 18 arguments shows)
        self._c_cfg = C.byref(self.cfg)
        self._c_params = _ptr(self._params)
        self._c_seg_off, self._c_seg_div = _ptr(self._seg_off), _ptr(self._seg_div)
        self._mb_step = nat.lib().mb_step

    # ---- shapes ---------------------------------------------------------------------
    @property
    def n_muscles(self) -> int:
        return self.rows * self.S

    @property
    def n_units(self) -> int:
        return self.rows * self.L

    @property
    def motor_unit_count(self):
        return self.specs[0].n if self.uniform else [s.n for s in self.specs]

    @property
    def max_excitation(self):
        return self.specs[0].pool.max_excitation

    def _shape_out(self, t: torch.Tensor) -> torch.Tensor:
        return t.view(self.rows) if self.uniform else t

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _as_input(self, x, per_unit_ok=True):
        """Bring an excitation argument to a contiguous device tensor of the mode's dtype.
        Returns (tensor, per_unit)."""
        if isinstance(x, torch.Tensor) and x.dtype == self.dtype and x.device == self.device and x.is_contiguous():
            n = x.numel()                                    # fast path: nothing to convert (only the pointer is used)
            if n == self.rows * self.S:
                return x, False
            if per_unit_ok and n == self.rows * self.L:
                return x, True
        if isinstance(x, (int, float)):
            t = torch.full((self.rows, self.S), float(x), dtype=self.dtype, device=self.device)
            return t, False
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"excitation must be a number, ndarray or torch tensor, not {type(x).__name__}")
        t = x.to(device=self.device, dtype=self.dtype, non_blocking=True)
        if t.numel() == self.n_muscles:
            return t.reshape(self.rows, self.S).contiguous(), False
        if per_unit_ok and t.numel() == self.n_units:
            return t.reshape(self.rows, self.L).contiguous(), True
        raise AssertionError(
            f"excitation has {t.numel()} elements; expected {self.n_muscles} (one per muscle)"
            + (f" or {self.n_units} (one per motor unit)" if per_unit_ok else ""))

    # ---- K1: one step -----------------------------------------------------------------
    def step(self, excitation, step_size: float, out: Optional[torch.Tensor] = None,
             mask_out: Optional[torch.Tensor] = None, stage: int = nat.MB_STAGE_MUSCLE,
             rates_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Advance every muscle one step (Muscle.step, muscle.py:65-92).

        excitation  number, or tensor/ndarray with one value per muscle ([rows] / [rows, S]) or
                    one per motor unit ([rows, L])
        returns     per-muscle totals, [rows] for a uniform batch, [rows, S] otherwise
        """
        x, per_unit = self._as_input(excitation)
        if stage == nat.MB_STAGE_FIBERS and not per_unit:
            raise AssertionError("firing rates must be given per motor unit")
        if out is None:
            out = torch.empty(self.rows, self.S, dtype=self.dtype, device=self.device) \
                if self.fresh_outputs else self._totals
        else:
            assert out.is_contiguous() and out.numel() == self.n_muscles and out.dtype == self.dtype
        forces = None
        if stage != nat.MB_STAGE_POOL:
            forces = torch.empty_like(self._forces) if self.fresh_outputs else self._forces
        if mask_out is not None:
            assert mask_out.dtype == torch.uint8 and mask_out.numel() == self.n_units and mask_out.is_contiguous()
        if rates_out is not None:
            assert rates_out.dtype == self.dtype and rates_out.numel() == self.n_units and rates_out.is_contiguous()
        args = (self._c_cfg, self._c_params, self.rows, self.L, self.S, self._c_seg_off, self._c_seg_div, stage,
                C.c_void_p(x.data_ptr()), int(per_unit), C.c_void_p(self.cpf.data_ptr()), C.c_void_p(self.dur.data_ptr()),
                _ptr(forces), C.c_void_p(out.data_ptr()) if stage != nat.MB_STAGE_POOL else None, _ptr(mask_out),
                _ptr(rates_out), float(step_size), self._stream())
        if torch.cuda.current_device() == self.device.index:
            rc = self._mb_step(*args)
        else:
            with torch.cuda.device(self.device):
                rc = self._mb_step(*args)
        if rc:
            nat.check(rc)
        self.launches += 1
        if forces is not None:
            self._forces = forces
        return self._shape_out(out.view(self.rows, self.S))

    # ---- K1 for the scalar drop-in API: zero-copy host I/O ---------------------------------
    def step_hostio(self, excitation, step_size: float) -> np.ndarray:
        """One step whose input, totals and per-unit forces live in pinned HOST memory that the
        kernel reads and writes directly (zero-copy over PCIe): a step is ONE kernel launch and
        ONE stream synchronisation -- no fill kernel for the scalar excitation, no device-to-host
        copies.  For the batch-of-1 drop-in classes (Muscle.step -> np.float64), where launch
        latency is everything.  Returns the totals as a host array [rows, S] (a view that the
        next call overwrites); ``host_forces()`` gives the per-unit forces of this step."""
        h = getattr(self, "_hostio", None)
        if h is None:
            npdt = np.float32 if self.dtype == torch.float32 else np.float64
            mk = lambda *shape: torch.zeros(*shape, dtype=self.dtype).pin_memory()
            h = dict(x=mk(self.rows, self.S), xu=mk(self.rows, self.L), out=mk(self.rows, self.S),
                     forces=mk(self.rows, self.L))
            h["x_np"], h["xu_np"] = h["x"].numpy(), h["xu"].numpy()
            h["out_np"], h["forces_np"] = h["out"].numpy(), h["forces"].numpy()
            h["npdt"] = npdt
            self._hostio = h
        if isinstance(excitation, (int, float)):
            h["x_np"].fill(excitation)
            src, per_unit = h["x"], 0
        else:
            a = np.asarray(excitation, dtype=h["npdt"])
            if a.size == self.n_muscles:
                h["x_np"][...] = a.reshape(self.rows, self.S)
                src, per_unit = h["x"], 0
            elif a.size == self.n_units:
                h["xu_np"][...] = a.reshape(self.rows, self.L)
                src, per_unit = h["xu"], 1
            else:
                raise AssertionError(f"excitation has {a.size} elements; expected {self.n_muscles} or {self.n_units}")
        st = torch.cuda.current_stream(self.device)
        # the argument list never changes between calls except for the input form and the step size
        fixed = h.get("args")
        key = (self.cpf.data_ptr(), self.dur.data_ptr())       # (the state tensors may have been swapped)
        if fixed is None or h.get("key") != key:
            h["key"] = key
            fixed = h["args"] = (nat.lib().mb_step, C.byref(self.cfg), _ptr(self._params), self.rows, self.L, self.S,
                                 _ptr(self._seg_off), _ptr(self._seg_div), nat.MB_STAGE_MUSCLE,
                                 (_ptr(h["x"]), _ptr(h["xu"])), _ptr(self.cpf), _ptr(self.dur),
                                 _ptr(h["forces"]), _ptr(h["out"]))
        fn, cfg, par, rows, L, S, so, sd, stage, srcs, cpf, dur, fo, out = fixed
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                rc = fn(cfg, par, rows, L, S, so, sd, stage, srcs[per_unit], per_unit, cpf, dur, fo, out, None, None,
                        float(step_size), C.c_void_p(st.cuda_stream))
        else:
            rc = fn(cfg, par, rows, L, S, so, sd, stage, srcs[per_unit], per_unit, cpf, dur, fo, out, None, None,
                    float(step_size), C.c_void_p(st.cuda_stream))
        if rc:
            nat.check(rc)
        self.launches += 1
        st.synchronize()
        return h["out_np"]

    def host_forces(self) -> np.ndarray:
        """Per-unit forces of the last ``step_hostio`` call, as a FRESH host array (fibers:258)."""
        return np.array(self._hostio["forces_np"], dtype=np.float64)

    # ---- K1 with the fused environment epilogue -------------------------------------------
    def step_env(self, excitation, step_size: float, gain: Optional[torch.Tensor] = None,
                 want_fatigue: bool = True, hill: Optional[tuple] = None):
        """One step for a vector environment: per-muscle totals (optionally multiplied by
        ``gain``, e.g. Hill-type force-length x force-velocity factors) and the peripheral
        fatigue of every muscle after the step, in ONE launch (mb_step_env).

        Equivalent to ``[m.step(a, dt) for m in muscles]`` followed by
        ``[m.get_peripheral_fatigue() for m in muscles]`` over every agent
        (examples/envs/muscled_hopper.py:33-34, muscle.py:248-256).
        ``hill=(rest, cur, prev)`` -- three [rows, S] tensors of muscle lengths -- makes the kernel's
        epilogue evaluate the Hill-type force-length and force-velocity curves (hill_type.py:15-66)
        itself and scale each muscle's total by their product; nothing else is launched.
        Returns ``(totals, fatigue)``; fatigue is float64 (None if not wanted)."""
        x, per_unit = self._as_input(excitation, per_unit_ok=False)
        out = torch.empty(self.rows, self.S, dtype=self.dtype, device=self.device)
        forces = torch.empty_like(self._forces) if self.fresh_outputs else self._forces
        fat = torch.empty(self.rows, self.S, dtype=torch.float64, device=self.device) if want_fatigue else None
        g = None
        if gain is not None:
            g = gain.to(device=self.device, dtype=self.dtype).reshape(self.rows, self.S).contiguous()
        lengths = None
        if hill is not None:
            lengths = [t.to(device=self.device, dtype=self.dtype).expand(self.rows, self.S).contiguous() for t in hill]
            if len(lengths) != 3:
                raise ValueError("hill must be (rest_lengths, current_lengths, previous_lengths)")
        try:
            with torch.cuda.device(self.device):
                if lengths is None:
                    nat.check(nat.lib().mb_step_env(
                        C.byref(self.cfg), _ptr(self._params), self.rows, self.L, self.S,
                        _ptr(self._seg_off), _ptr(self._seg_div), _ptr(x), _ptr(self.cpf), _ptr(self.dur),
                        _ptr(forces), _ptr(out), _ptr(g), _ptr(fat), float(step_size), self._stream()))
                else:
                    hp = nat.MbHill(lengths[0].data_ptr(), lengths[1].data_ptr(), lengths[2].data_ptr(), 17.33, 3.0, 1.8)
                    nat.check(nat.lib().mb_step_env_hill(
                        C.byref(self.cfg), _ptr(self._params), self.rows, self.L, self.S,
                        _ptr(self._seg_off), _ptr(self._seg_div), _ptr(x), _ptr(self.cpf), _ptr(self.dur),
                        _ptr(forces), _ptr(out), _ptr(g), _ptr(fat), C.byref(hp), float(step_size), self._stream()))
            self.launches += 1
            self._forces = forces
        except nat.UnsupportedError:                         # layouts the streaming kernel does not cover
            out = self.step(x, step_size).view(self.rows, self.S)
            if g is not None:
                out = out * g
            if lengths is not None:
                from .hill_type import (contractile_element_force_length_curve as fl,
                                        contractile_element_force_velocity_curve as fv)
                out = out * fl(lengths[0], lengths[1]) * fv(lengths[0], lengths[1], lengths[2], step_size)
            if want_fatigue:
                fat = self.get_peripheral_fatigue().view(self.rows, self.S)
        return self._shape_out(out), (self._shape_out(fat) if fat is not None else None)

    # ---- K2: K fused steps ------------------------------------------------------------
    def step_many(self, excitations, step_size: float, steps: Optional[int] = None,
                  out: Optional[torch.Tensor] = None, forces_every: int = 0,
                  return_masks: bool = False, units_per_thread: int = 0, hold: int = 1, totals_every: int = 1):
        """Advance K steps in one launch with the unit state held in registers.

        excitations  [K, M] (one row of per-muscle values per step) or [M] with ``steps=K``
                     (constant over the window, the README loop's form: README.md:122-130); with
                     ``hold=h`` [ceil(K/h), M] and ``steps=K``: row i drives steps i*h .. i*h+h-1 (a
                     controller slower than the simulation)
        totals_every  t: only the totals of steps t-1, 2t-1, ... are produced ([K // t, M])
        returns      totals [K, M]; with ``forces_every=f`` also the per-unit forces of steps
                     f-1, 2f-1, ... as [K//f, rows, L] (and their recruitment masks if asked)
        Falls back to K single-step launches (still on the GPU) for layouts the fused kernel
        does not cover (ragged rows, n > 512).
        """
        if isinstance(excitations, np.ndarray):
            excitations = torch.from_numpy(np.ascontiguousarray(excitations))
        e = excitations.to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()
        M = self.n_muscles
        hold, tev = int(hold), int(totals_every)
        if hold < 1 or tev < 1:
            raise ValueError("hold and totals_every must be >= 1")
        if e.numel() == M and steps is not None:
            K, stride_k = int(steps), 0
            e = e.reshape(M)
        elif hold > 1:
            if steps is None:
                raise ValueError("hold > 1 needs steps=K")
            K, stride_k = int(steps), M
            rows_needed = (K + hold - 1) // hold
            if e.numel() != rows_needed * M:
                raise AssertionError(f"excitations must be [{rows_needed}, {M}] for {K} steps held {hold} each (got {tuple(e.shape)})")
            e = e.reshape(rows_needed, M)
        elif e.dim() >= 2 and e.shape[0] * M == e.numel() and (steps is None or int(steps) == e.shape[0]):
            K, stride_k = int(e.shape[0]), M
            e = e.reshape(K, M)
        else:
            raise AssertionError(f"excitations must be [K, {M}] or [{M}] with steps=K (got {tuple(e.shape)})")
        if K <= 0:
            raise ValueError("steps must be positive")
        Kt = K // tev
        if out is None:
            out = torch.empty(Kt, M, dtype=self.dtype, device=self.device)
        else:
            assert out.is_contiguous() and out.dtype == self.dtype and tuple(out.shape) == (Kt, M)
        n_dec = K // forces_every if forces_every > 0 else 0
        forces_dec = masks = None
        if n_dec:
            forces_dec = torch.empty(n_dec, self.rows, self.L, dtype=self.dtype, device=self.device)
            if return_masks:
                masks = torch.empty(n_dec, self.rows, self.L, dtype=torch.uint8, device=self.device)
        forces_last = torch.empty_like(self._forces) if self.fresh_outputs else self._forces

        w = nat.MbFusedArgs()
        w.rows, w.L, w.S, w.K = self.rows, self.L, self.S, K
        w.seg_off_dev = self._seg_off.data_ptr() if self._seg_off is not None else None
        w.seg_div_dev = self._seg_div.data_ptr() if self._seg_div is not None else None
        w.exc_dev, w.exc_stride_k, w.exc_hold = e.data_ptr(), stride_k, hold
        w.totals_every = tev
        w.cpf_dev, w.dur_dev = self.cpf.data_ptr(), self.dur.data_ptr()
        w.totals_dev, w.totals_stride_k = out.data_ptr(), M
        w.forces_dec_dev = forces_dec.data_ptr() if forces_dec is not None else None
        w.mask_dec_dev = masks.data_ptr() if masks is not None else None
        w.forces_every = int(forces_every if n_dec else 0)
        w.units_per_thread = int(units_per_thread)
        w.forces_last_dev = forces_last.data_ptr()
        w.max_segment_units = max(s.n for s in self.specs)
        fused = True
        try:
            with torch.cuda.device(self.device):
                nat.check(nat.lib().mb_step_fused_ex(C.byref(self.cfg), _ptr(self._params), C.byref(w),
                                                     float(step_size), self._stream()))
            self.launches += 1
            self._forces = forces_last
        except nat.UnsupportedError:
            fused = False
        if not fused:
            # K single-step launches (still on the GPU) writing into the freshly allocated force buffer: a tensor
            # handed out earlier through current_forces is never overwritten (fibers:258)
            keep = self.fresh_outputs
            self.fresh_outputs = False
            self._forces = forces_last
            scratch = torch.empty(self.rows, self.S, dtype=self.dtype, device=self.device) if tev > 1 else None
            try:
                for k in range(K):
                    mk = masks[(k + 1) // forces_every - 1].view(-1) if (masks is not None and (k + 1) % forces_every == 0) else None
                    dst = out[k // tev].view(self.rows, self.S) if (k + 1) % tev == 0 else scratch
                    self.step(e[k // hold] if stride_k else e, step_size, out=dst, mask_out=mk)
                    if n_dec and (k + 1) % forces_every == 0:
                        forces_dec[(k + 1) // forces_every - 1].copy_(self._forces)
            finally:
                self.fresh_outputs = keep
        if n_dec:
            return (out, forces_dec, masks) if return_masks else (out, forces_dec)
        return out

    # ---- K2 with the gather of totals fused into the kernel (multi-GPU) -----------------
    def step_many_gather(self, excitations: torch.Tensor, step_size: float, out: torch.Tensor,
                         peer_ptrs: Sequence[int], peer_stride_k: int, peer_stride_pair: int,
                         peer_stride_odd: int = 1, multicast: bool = False) -> None:
        """``step_many`` whose totals go to ``out`` [K, M] (local) AND to raw device addresses in
        every peer GPU's gather buffer (``peer_ptrs``, mapped through symmetric memory and already
        offset to this rank's region; element (k, m) at (m >> 1) * peer_stride_pair +
        k * peer_stride_k + (m & 1) * peer_stride_odd) -- the kernel stores each total to all of
        them, so the all-gather rides on the kernel's own stores over NVLink.  ``multicast=True``:
        ``peer_ptrs`` is ONE NVLink multicast address and every total leaves this GPU once.  Uniform
        batches only; the caller synchronises the ranks afterwards."""
        e = excitations.to(device=self.device, dtype=self.dtype).contiguous()
        M = self.n_muscles
        assert self.S == 1 and e.dim() == 2 and e.shape[1] == M, "fused gather needs a uniform batch and exc [K, M]"
        K = int(e.shape[0])
        assert out.is_contiguous() and out.dtype == self.dtype and tuple(out.shape) == (K, M)
        forces_last = torch.empty_like(self._forces) if self.fresh_outputs else self._forces
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        w = nat.MbFusedArgs()
        w.rows, w.L, w.S, w.K = self.rows, self.L, 1, K
        w.exc_dev, w.exc_stride_k, w.exc_hold, w.totals_every = e.data_ptr(), M, 1, 1
        w.cpf_dev, w.dur_dev = self.cpf.data_ptr(), self.dur.data_ptr()
        w.totals_dev, w.totals_stride_k = out.data_ptr(), M
        w.forces_last_dev = forces_last.data_ptr()
        w.peer_totals_dev, w.n_peers = arr, len(peer_ptrs)
        w.peer_stride_k, w.peer_stride_pair, w.peer_stride_odd = int(peer_stride_k), int(peer_stride_pair), int(peer_stride_odd)
        w.peer_multicast = int(bool(multicast))
        with torch.cuda.device(self.device):
            nat.check(nat.lib().mb_step_fused_ex(C.byref(self.cfg), _ptr(self._params), C.byref(w), float(step_size),
                                                 self._stream()))
        self.launches += 1
        self._forces = forces_last

    # ---- K2 with host buffers: H2D / compute / D2H pipelined over sub-windows ---------
    def step_many_host(self, exc_host: torch.Tensor, step_size: float, out_host: torch.Tensor,
                       chunk: int = 256, overlap_calls: bool = False, steps: Optional[int] = None,
                       hold: int = 1, totals_every: int = 1) -> torch.Tensor:
        """``step_many`` for excitations and totals that live in (pinned) HOST memory.

        The window is cut into sub-windows of ``chunk`` steps; the host->device copy of
        sub-window i+1, the fused kernel of sub-window i and the device->host copy of
        sub-window i-1 run concurrently on three streams (double-buffered staging), so the
        call costs max(PCIe, kernel) rather than their sum.  Returns ``out_host``; the
        caller's current stream is made to wait for the last copy.

        exc_host   [K, M] one row per step; or [M] with ``steps=K``: constant over the window (the
                   README loop, README.md:122-130 -- 4 bytes per muscle cross PCIe instead of 4 K);
                   or [ceil(K/h), M] with ``steps=K, hold=h``: each row held for h steps
        out_host   [K // totals_every, M]: the totals of steps t-1, 2t-1, ... (t = totals_every)
        overlap_calls  False: the first upload waits for everything already queued on the current
                       stream (safe if ``exc_host`` is being produced by stream work).  True: the
                       upload only waits for the staging buffer to be free, so back-to-back calls
                       overlap this window's first upload with the previous window's last download
                       (``exc_host`` must already hold its final contents at call time, and the
                       caller joins with ``host_pipeline_join()`` before reading ``out_host``).
        """
        M = self.n_muscles
        hold, tev = int(hold), int(totals_every)
        if hold < 1 or tev < 1:
            raise ValueError("hold and totals_every must be >= 1")
        assert exc_host.device.type == "cpu" and out_host.device.type == "cpu"
        assert exc_host.dtype == self.dtype and out_host.dtype == self.dtype
        const = exc_host.numel() == M and steps is not None
        if const:
            K = int(steps)
        elif hold > 1:
            if steps is None:
                raise ValueError("hold > 1 needs steps=K")
            K = int(steps)
            assert tuple(exc_host.shape) == ((K + hold - 1) // hold, M)
        else:
            K = int(exc_host.shape[0])
            assert tuple(exc_host.shape) == (K, M) and (steps is None or int(steps) == K)
        assert tuple(out_host.shape) == (K // tev, M)
        # sub-window boundaries fall on multiples of both periods; the staging buffers are sized by the REQUESTED
        # chunk (a shorter last sub-window uses a slice), so the pipeline -- streams, events, buffers with copies
        # possibly still in flight -- is never rebuilt because a window happens to be short
        period = hold * tev // np.gcd(hold, tev)
        chunk = max(period, int(chunk) // period * period)
        st = getattr(self, "_host_pipe", None)
        if st is None or st["chunk"] != chunk or st["hold"] != hold or st["tev"] != tev:
            if st is not None:                                # retire the old pipeline only once it has drained
                cur0 = torch.cuda.current_stream(self.device)
                cur0.wait_stream(st["h2d"])
                cur0.wait_stream(st["d2h"])
                st["h2d"].wait_stream(cur0)
                st["d2h"].wait_stream(cur0)
                st["h2d"].synchronize()
                st["d2h"].synchronize()
            with torch.cuda.device(self.device):
                st = dict(chunk=chunk, hold=hold, tev=tev,
                          h2d=torch.cuda.Stream(self.device), d2h=torch.cuda.Stream(self.device),
                          exc=None,                          # staging for per-step / held excitations: allocated on first use
                          tot=[torch.empty(chunk // tev, M, dtype=self.dtype, device=self.device) for _ in range(2)],
                          cexc=[torch.empty(M, dtype=self.dtype, device=self.device) for _ in range(2)],
                          up=[torch.cuda.Event() for _ in range(2)], run=[torch.cuda.Event() for _ in range(2)],
                          down=[torch.cuda.Event() for _ in range(2)], cup=torch.cuda.Event(), crun=[None, None])
            self._host_pipe = st
        cur = torch.cuda.current_stream(self.device)
        if not overlap_calls:
            st["h2d"].wait_stream(cur)
        ce = None
        if const:                                            # one small upload per call, double-buffered across calls
            cb = st.get("cseq", 0) & 1
            st["cseq"] = st.get("cseq", 0) + 1
            if st["crun"][cb] is not None:
                st["h2d"].wait_event(st["crun"][cb])         # the last kernel that read this buffer is done
            with torch.cuda.stream(st["h2d"]):
                st["cexc"][cb].copy_(exc_host.reshape(M), non_blocking=True)
                st["cup"].record(st["h2d"])
            cur.wait_event(st["cup"])
            ce = st["cexc"][cb]
        seq = st.get("seq", 0)                               # sub-windows issued so far: the staging buffers alternate
        for k0 in range(0, K, chunk):                        # ACROSS calls too (a window may hold an odd number of them)
            i, seq = seq, seq + 1
            b, c = i & 1, min(chunk, K - k0)
            ce_rows, ct_rows = (c + hold - 1) // hold, c // tev
            if not const:
                if st["exc"] is None:
                    with torch.cuda.device(self.device):
                        st["exc"] = [torch.empty(chunk // hold, M, dtype=self.dtype, device=self.device) for _ in range(2)]
                if i >= 2:
                    st["h2d"].wait_event(st["run"][b])       # the last kernel that read exc[b] is done
                with torch.cuda.stream(st["h2d"]):
                    st["exc"][b][:ce_rows].copy_(exc_host[k0 // hold:k0 // hold + ce_rows], non_blocking=True)
                    st["up"][b].record(st["h2d"])
                cur.wait_event(st["up"][b])
            if i >= 2:
                cur.wait_event(st["down"][b])                # the last copy-out of tot[b] has drained
            if ct_rows or c:
                self.step_many(ce if const else st["exc"][b][:ce_rows], step_size, steps=(c if (const or hold > 1) else None),
                               out=st["tot"][b][:ct_rows], hold=hold, totals_every=tev)
            st["run"][b].record(cur)
            st["d2h"].wait_event(st["run"][b])
            with torch.cuda.stream(st["d2h"]):
                if ct_rows:
                    out_host[k0 // tev:k0 // tev + ct_rows].copy_(st["tot"][b][:ct_rows], non_blocking=True)
                st["down"][b].record(st["d2h"])
        if const:
            ev = torch.cuda.Event()
            ev.record(cur)
            st["crun"][cb] = ev
        st["seq"] = seq
        if not overlap_calls:
            cur.wait_stream(st["d2h"])
        return out_host

    def host_pipeline_join(self):
        """Make the current stream wait for every download queued by ``step_many_host``
        (needed after calls with ``overlap_calls=True`` before the results are consumed)."""
        st = getattr(self, "_host_pipe", None)
        if st is not None:
            torch.cuda.current_stream(self.device).wait_stream(st["d2h"])

    # ---- observers --------------------------------------------------------------------
    @property
    def current_forces(self) -> torch.Tensor:
        """Per-unit forces of the last step, [rows, L] (Muscle.current_forces, muscle.py:61-63)."""
        return self._forces

    @property
    def current_peak_forces(self) -> torch.Tensor:
        """Per-unit capacities, [rows, L] float64 (fibers:92-94)."""
        return self.cpf

    def get_peripheral_fatigue(self) -> torch.Tensor:
        """1 - sum(capacity)/max_arb_output per muscle (StandardMuscle.get_peripheral_fatigue,
        muscle.py:248-256), float64."""
        out = torch.empty(self.rows, self.S, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().mb_peripheral_fatigue(
                C.byref(self.cfg), _ptr(self._params), self.rows, self.L, self.S, _ptr(self._seg_off), _ptr(self._seg_div),
                _ptr(self.cpf), _ptr(out), self._stream()))
        self.launches += 1
        return self._shape_out(out)

    def fused_geometry(self, steps: int, units_per_thread: int = 0) -> dict:
        mpb, thr, smem, blocks = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        with torch.cuda.device(self.device):
            nat.check(nat.lib().mb_fused_geometry(C.byref(self.cfg), self.n_muscles, self.L, int(steps),
                                                  int(units_per_thread), C.byref(mpb), C.byref(thr),
                                                  C.byref(blocks), C.byref(smem)))
        return dict(muscles_per_block=mpb.value, threads=thr.value, blocks=blocks.value, smem_bytes=smem.value)

    # ---- copying / pickling (the reference's objects are plain Python and can be copied) -------
    def __getstate__(self):
        ctor = dict(self._ctor)
        ctor["device"] = str(self.device)
        if isinstance(ctor["precision"], torch.dtype):
            ctor["precision"] = "f32" if ctor["precision"] == torch.float32 else "f64"
        return {"ctor": ctor, "cpf": self.cpf.cpu(), "dur": self.dur.cpu(), "forces": self._forces.cpu(),
                "launches": self.launches}

    def __setstate__(self, state):
        self.__init__(**state["ctor"])
        self.cpf.copy_(state["cpf"])
        self.dur.copy_(state["dur"])
        self._forces = state["forces"].to(self.device)
        self.launches = state["launches"]

    # ---- state ------------------------------------------------------------------------
    def reset(self):
        """Back to the rested state (fibers:63, pool:79)."""
        self.cpf.copy_(self._ptf if self.per_instance else self._ptf.repeat(self.rows, 1))
        self.dur.zero_()
        self._forces = torch.zeros_like(self._forces)

    def state_dict(self) -> dict:
        """The whole simulator state: two float64 tensors."""
        return {"current_peak_forces": self.cpf.clone(), "recruitment_durations": self.dur.clone()}

    def load_state_dict(self, state: dict):
        self.cpf.copy_(state["current_peak_forces"].to(self.device, torch.float64).reshape(self.rows, self.L))
        self.dur.copy_(state["recruitment_durations"].to(self.device, torch.float64).reshape(self.rows, self.L))
