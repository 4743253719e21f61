"""
ctypes binding of libmuscle_b200.so (C ABI: include/muscle_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C pymuscle_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB_LIB") or os.path.join(_HERE, "csrc", "libmuscle_b200.so")   # MB_LIB: tuning builds

ABI_VERSION = 2

MB_OK, MB_EINVAL, MB_EUNSUPPORTED, MB_ECUDA = 0, -1, -2, -3
MB_F32, MB_F64 = 0, 1
MB_FIBERS_POTVIN, MB_FIBERS_PYMUSCLE = 0, 1
MB_STAGE_MUSCLE, MB_STAGE_POOL, MB_STAGE_FIBERS = 0, 1, 2
MB_PROP_ADAPT_NONNEG, MB_PROP_EXP_BOUNDED = 1, 2

#: every symbol include/muscle_b200.h declares (tests check the .so exports them all)
SYMBOLS = (
    "mb_abi_version", "mb_last_error", "mb_params_bytes", "mb_params_pack", "mb_step", "mb_step_env",
    "mb_step_fused", "mb_step_fused_gather", "mb_fused_geometry", "mb_peripheral_fatigue", "mb_peak_probe",
    "mb_step_env_hill", "mb_step_fused_ex",
)


class MbConfig(C.Structure):
    """Mirror of ``struct MbConfig`` (include/muscle_b200.h)."""
    _fields_ = [
        ("precision", C.c_int32),
        ("fibers_kind", C.c_int32),
        ("central_fatigue", C.c_int32),
        ("peripheral_fatigue", C.c_int32),
        ("max_recruitment_threshold", C.c_double),
        ("firing_gain", C.c_double),
        ("min_firing_rate", C.c_double),
        ("derecruitment_delta", C.c_double),
        ("adaptation_magnitude", C.c_double),
        ("adaptation_time_constant", C.c_double),
        ("max_duration", C.c_double),
        ("contraction_time_change_ratio", C.c_double),
        ("input_scale", C.c_double),
        ("output_divisor", C.c_double),
        ("table_props", C.c_int32),
        ("param_rows", C.c_int32),
    ]


class MbHill(C.Structure):
    """Mirror of ``struct MbHill`` (include/muscle_b200.h)."""
    _fields_ = [
        ("rest_dev", C.c_void_p),
        ("cur_dev", C.c_void_p),
        ("prev_dev", C.c_void_p),
        ("curve_width_factor", C.c_double),
        ("max_velocity", C.c_double),
        ("max_eccentric_multiple", C.c_double),
    ]


class MbFusedArgs(C.Structure):
    """Mirror of ``struct MbFusedArgs`` (include/muscle_b200.h)."""
    _fields_ = [
        ("rows", C.c_int64), ("L", C.c_int64),
        ("S", C.c_int32), ("K", C.c_int32),
        ("seg_off_dev", C.c_void_p),
        ("seg_div_dev", C.c_void_p),
        ("exc_dev", C.c_void_p),
        ("exc_stride_k", C.c_int64),
        ("exc_hold", C.c_int32),
        ("totals_every", C.c_int32),
        ("cpf_dev", C.c_void_p),
        ("dur_dev", C.c_void_p),
        ("totals_dev", C.c_void_p),
        ("totals_stride_k", C.c_int64),
        ("forces_dec_dev", C.c_void_p),
        ("mask_dec_dev", C.c_void_p),
        ("forces_every", C.c_int32),
        ("units_per_thread", C.c_int32),
        ("forces_last_dev", C.c_void_p),
        ("peer_totals_dev", C.POINTER(C.c_void_p)),
        ("n_peers", C.c_int32),
        ("max_segment_units", C.c_int32),
        ("peer_stride_k", C.c_int64), ("peer_stride_pair", C.c_int64),
        ("peer_stride_odd", C.c_int64),
        ("peer_multicast", C.c_int32), ("reserved", C.c_int32),
    ]


class NativeError(RuntimeError):
    """A libmuscle_b200 entry point returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libmuscle_b200 error {code}: {message}")
        self.code = code


class UnsupportedError(NativeError):
    """MB_EUNSUPPORTED: the requested shape is outside a kernel's range."""


_lib = None


def _declare(lib):
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    cfgp = C.POINTER(MbConfig)
    lib.mb_abi_version.restype = C.c_int
    lib.mb_abi_version.argtypes = []
    lib.mb_last_error.restype = C.c_char_p
    lib.mb_last_error.argtypes = []
    lib.mb_params_bytes.restype = C.c_size_t
    lib.mb_params_bytes.argtypes = [cfgp, i64]
    lib.mb_params_pack.restype = C.c_int
    lib.mb_params_pack.argtypes = [cfgp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i32)]
    lib.mb_step.restype = C.c_int
    lib.mb_step.argtypes = [cfgp, vp, i64, i64, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, dbl, vp]
    lib.mb_step_env.restype = C.c_int
    lib.mb_step_env.argtypes = [cfgp, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, vp]
    lib.mb_step_fused.restype = C.c_int
    lib.mb_step_fused.argtypes = [cfgp, vp, i64, i32, i32, vp, i64, vp, vp, vp, i64, vp, vp, i32, vp, dbl, i32, vp]
    lib.mb_step_fused_gather.restype = C.c_int
    lib.mb_step_fused_gather.argtypes = [cfgp, vp, i64, i32, i32, vp, i64, vp, vp, vp, i64, C.POINTER(vp), i32,
                                         i64, i64, vp, dbl, vp]
    lib.mb_fused_geometry.restype = C.c_int
    lib.mb_fused_geometry.argtypes = [cfgp, i64, i32, i32, i32, C.POINTER(i32), C.POINTER(i32),
                                      C.POINTER(i64), C.POINTER(i32)]
    lib.mb_peripheral_fatigue.restype = C.c_int
    lib.mb_peripheral_fatigue.argtypes = [cfgp, vp, i64, i64, i32, vp, vp, vp, vp, vp]
    lib.mb_step_env_hill.restype = C.c_int
    lib.mb_step_env_hill.argtypes = [cfgp, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(MbHill), dbl, vp]
    lib.mb_step_fused_ex.restype = C.c_int
    lib.mb_step_fused_ex.argtypes = [cfgp, vp, C.POINTER(MbFusedArgs), dbl, vp]
    lib.mb_peak_probe.restype = C.c_int
    lib.mb_peak_probe.argtypes = [i32, i32, i32, i32, vp, C.POINTER(i64), vp]


def lib():
    """Load (once) and return the shared library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run __graft_entry__.build() or `make -C pymuscle_b200/csrc`). "
                "pymuscle_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        _declare(handle)
        got = handle.mb_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"libmuscle_b200.so has ABI {got}, binding expects {ABI_VERSION}")
        _lib = handle
    return _lib


def check(rc: int):
    if rc == MB_OK:
        return
    msg = lib().mb_last_error().decode("utf-8", "replace")
    if rc == MB_EUNSUPPORTED:
        raise UnsupportedError(rc, msg)
    raise NativeError(rc, msg)
