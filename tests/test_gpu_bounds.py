"""
Memory safety without compute-sanitizer (closed on the GPU pool): the library is also built with
``-DMB_BOUNDS`` (``make -C pymuscle_b200/csrc bounds``), which turns on device-side asserts on the
index of every state / output access and on the layout-dependent shared-memory slots.  This test runs
the odd-sized, ragged, partial-tile and peer-gather cases of the suite through that build in a child
process; a violated assert aborts the kernel (cudaErrorAssert) and fails the child.
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BOUNDS_LIB = os.path.join(ROOT, "pymuscle_b200", "csrc", "libmuscle_b200_bounds.so")


def test_index_asserts_hold_on_odd_and_ragged_shapes():
    if not os.path.exists(BOUNDS_LIB):
        pytest.skip("libmuscle_b200_bounds.so is not built (make -C pymuscle_b200/csrc bounds)")
    select = ("do_not_write_outside or fused_gather or ragged or wide or hold_and_totals or fused_fatigue or "
              "rows_of_small_muscles or arm_ragged or per_row_tables")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", select,
           os.path.join(ROOT, "tests", "test_gpu_round2.py"), os.path.join(ROOT, "tests", "test_gpu_env.py"),
           os.path.join(ROOT, "tests", "test_gpu_per_instance.py"), os.path.join(ROOT, "tests", "test_gpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT,
                       env={**os.environ, "MB_LIB": BOUNDS_LIB})
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
