"""
CPU-side checks of the C ABI: the shared library loads, exports every symbol
include/muscle_b200.h declares, and its host-side functions (parameter packing,
launch geometry, argument validation) behave.  No kernel is launched here.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle.pymuscle_oracle import PoolHyper, UnitTables, recruitment_mask
from pymuscle_b200 import _native as nat
from pymuscle_b200.tables import MuscleSpec, force_to_motor_unit_count, pack_row


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "muscle_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(nat.SYMBOLS)
    lib = nat.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mb_abi_version() == nat.ABI_VERSION


def test_config_struct_layout_matches_header():
    # 4 int32 + 10 double + 2 int32
    assert C.sizeof(nat.MbConfig) == 4 * 4 + 10 * 8 + 2 * 4


def test_tables_bit_identical_to_reference():
    g = golden("tables")
    for n in (2, 33, 120, 340, 2000):
        t = MuscleSpec(n, fibers=MuscleSpec.standard().fibers).tables()
        for key in ("thr", "peak", "ptf", "ct", "fat", "rec"):
            assert np.array_equal(t[key], g[f"{key}_{n}"]), (key, n)
    for mf, cnt in g["counts_for_force"]:
        assert force_to_motor_unit_count(float(mf), 0.0123) == int(cnt)
    assert MuscleSpec.standard().output_divisor == float(g["max_arb_output_120"])
    assert MuscleSpec.standard().input_scale == 67.0


@pytest.mark.parametrize("kind", ["potvin", "standard"])
def test_pack_f32_critical_inputs_reproduce_the_float64_predicate(kind):
    """crit[i] is the smallest float32 input the reference recruits unit i for."""
    spec = MuscleSpec.potvin(120) if kind == "potvin" else MuscleSpec.standard()
    row = pack_row([spec], nat.MB_F32)
    P = row.blob.view(np.float32).reshape(3, 120, 4)
    crit = P[0, :, 0]
    t, h = UnitTables.build(120), PoolHyper()
    scale = spec.input_scale
    for c_i, i in zip(crit, range(120)):
        below = np.nextafter(c_i, np.float32(-np.inf))
        on = recruitment_mask(np.float64(c_i) * scale, t, h)[i]
        off = recruitment_mask(np.float64(below) * scale, t, h)[i]
        assert on and not off
    # the packed float tables are the float64 tables rounded once
    tab = spec.tables()
    np.testing.assert_array_equal(P[0, :, 2], tab["peak"].astype(np.float32))
    np.testing.assert_array_equal(P[1, :, 3], tab["fat"].astype(np.float32))
    np.testing.assert_array_equal(P[2, :, 0], tab["ptf"].astype(np.float32))
    np.testing.assert_array_equal(P[2, :, 0].astype(np.float64) + P[2, :, 3], tab["ptf"].astype(np.float32).astype(np.float64)
                                  + (tab["ptf"] - tab["ptf"].astype(np.float32)).astype(np.float32))
    assert row.cfg.table_props & nat.MB_PROP_ADAPT_NONNEG


def test_pack_f64_layout():
    spec = MuscleSpec.standard()
    row = pack_row([spec], nat.MB_F64)
    P = row.blob.view(np.float64).reshape(5, 120, 2)
    tab = spec.tables()
    np.testing.assert_array_equal(P[0, :, 0], tab["thr"])
    np.testing.assert_array_equal(P[0, :, 1], tab["peak"])
    np.testing.assert_array_equal(P[3, :, 0], tab["fat"])
    np.testing.assert_array_equal(P[3, :, 1], tab["ptf"])
    np.testing.assert_array_equal(P[4, :, 0], tab["rec"])
    np.testing.assert_allclose(P[2, :, 0], tab["ct"] * 1.379 / 1000, rtol=1e-15)


def test_pack_ragged_row():
    specs = [MuscleSpec.standard(max_force=f) for f in (68.0, 90.0, 120.0)]
    row = pack_row(specs, nat.MB_F32)
    assert list(row.seg_off) == [0, specs[0].n, specs[0].n + specs[1].n, sum(s.n for s in specs)]
    assert row.L == sum(s.n for s in specs)
    np.testing.assert_array_equal(row.seg_div, [s.output_divisor for s in specs])
    with pytest.raises(ValueError):
        pack_row([MuscleSpec.standard(), MuscleSpec.potvin(120)], nat.MB_F32)


def test_argument_validation_without_a_gpu():
    lib = nat.lib()
    cfg = pack_row([MuscleSpec.potvin(120)], nat.MB_F32).cfg
    rc = lib.mb_step(C.byref(cfg), None, 1, 120, 1, None, None, 0, None, 0, None, None, None, None, None, None, 0.02, None)
    assert rc == nat.MB_EINVAL and b"NULL" in lib.mb_last_error()
    with pytest.raises(nat.NativeError):
        nat.check(rc)
    bad = nat.MbConfig()
    C.memmove(C.byref(bad), C.byref(cfg), C.sizeof(cfg))
    bad.precision = 7
    assert lib.mb_params_bytes(C.byref(bad), 120) > 0
    assert lib.mb_params_pack(C.byref(bad), 120, None, None, None, None, None, None, None, None) == nat.MB_EINVAL
    # fused kernel range
    mpb, thr, smem, blocks = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    rc = lib.mb_fused_geometry(C.byref(cfg), 65536, 2000, 1000, 0, C.byref(mpb), C.byref(thr), C.byref(blocks), C.byref(smem))
    assert rc == nat.MB_EUNSUPPORTED
    with pytest.raises(nat.UnsupportedError):
        nat.check(rc)
    # the environment step and the fused gather validate before touching the device as well
    rc = lib.mb_step_env(C.byref(cfg), None, 1, 120, 1, None, None, None, None, None, None, None, None, None, 0.02, None)
    assert rc == nat.MB_EINVAL and b"mb_step_env" in lib.mb_last_error()
    rc = lib.mb_step_fused_gather(C.byref(cfg), None, 8, 120, 10, None, 8, None, None, None, 8, None, 0, 0, 0, None, 0.02, None)
    assert rc == nat.MB_EINVAL
    # per-instance tables: param_rows must match the batch
    per = nat.MbConfig()
    C.memmove(C.byref(per), C.byref(cfg), C.sizeof(cfg))
    per.param_rows = 5
    rc = lib.mb_fused_geometry(C.byref(per), 6, 120, 10, 0, C.byref(mpb), C.byref(thr), C.byref(blocks), C.byref(smem))
    assert rc == nat.MB_EINVAL and b"param_rows" in lib.mb_last_error()
    assert lib.mb_fused_geometry(C.byref(per), 5, 120, 10, 0, C.byref(mpb), C.byref(thr), C.byref(blocks), C.byref(smem)) == 0
    assert thr.value == 256 and mpb.value == 8                                 # k2u_f32: one warp per muscle


def test_per_instance_packing_and_its_checks():
    from dataclasses import replace
    from pymuscle_b200.tables import pack_rows_per_instance
    base = MuscleSpec.potvin(24)
    rows = [replace(base, fibers=replace(base.fibers, max_fatigue_rate=0.02 + 0.001 * i)) for i in range(3)]
    pk = pack_rows_per_instance(rows, nat.MB_F32)
    one = pack_row([rows[1]], nat.MB_F32)
    assert pk.cfg.param_rows == 3 and pk.blob.size == 3 * one.blob.size
    assert np.array_equal(pk.blob[one.blob.size: 2 * one.blob.size], one.blob)    # table 1 is row 1's own table
    assert pk.ptf.shape == (3, 24)
    with pytest.raises(ValueError, match="per-launch scalar"):
        pack_rows_per_instance([rows[0], replace(rows[1], pool=replace(rows[1].pool, min_firing_rate=9))], nat.MB_F32)
    with pytest.raises(ValueError):
        pack_rows_per_instance(rows[:1], nat.MB_F32)


def test_fused_geometry_for_the_headline_config():
    lib = nat.lib()
    cfg = pack_row([MuscleSpec.potvin(120)], nat.MB_F32).cfg
    mpb, thr, smem, blocks = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    # auto: the warp kernel (one warp per pair of muscles, 8 warps per block)
    nat.check(lib.mb_fused_geometry(C.byref(cfg), 65536, 120, 1000, 0, C.byref(mpb), C.byref(thr),
                                    C.byref(blocks), C.byref(smem)))
    assert thr.value == 256 and mpb.value == 16 and blocks.value == 4096
    assert smem.value <= 110 * 1024                                           # two blocks per SM
    # explicit units_per_thread: the block kernel, 4 groups x 120 threads = 15 full warps
    nat.check(lib.mb_fused_geometry(C.byref(cfg), 65536, 120, 1000, 4, C.byref(mpb), C.byref(thr),
                                    C.byref(blocks), C.byref(smem)))
    assert thr.value == 480 and mpb.value == 16 and blocks.value == 4096
    assert smem.value <= 110 * 1024
    # FP64 mode: one warp per muscle, 8 muscles per block, two blocks per SM
    cfg64 = pack_row([MuscleSpec.potvin(120)], nat.MB_F64).cfg
    nat.check(lib.mb_fused_geometry(C.byref(cfg64), 32768, 120, 1000, 0, C.byref(mpb), C.byref(thr),
                                    C.byref(blocks), C.byref(smem)))
    assert thr.value == 256 and mpb.value == 8 and blocks.value == 4096 and smem.value <= 110 * 1024
    # ... explicit units_per_thread: the FP64 block kernel, 4 instances per thread, one block per SM
    nat.check(lib.mb_fused_geometry(C.byref(cfg64), 32768, 120, 1000, 4, C.byref(mpb), C.byref(thr),
                                    C.byref(blocks), C.byref(smem)))
    assert thr.value == 480 and mpb.value == 16 and smem.value <= 220 * 1024


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libmuscle_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        nat.lib()


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/muscle_b200.h compiles as C (gcc, no CUDA), and sizeof / offsetof of every struct that crosses the
    boundary equal the ctypes mirrors in pymuscle_b200/_native.py."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    mirrors = {"MbConfig": nat.MbConfig, "MbHill": nat.MbHill, "MbFusedArgs": nat.MbFusedArgs}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "muscle_b200.h"', "int main(void) {"]
    for name, cls in mirrors.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in mirrors.items():
        assert int(out[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, f"{name}.{field}"
