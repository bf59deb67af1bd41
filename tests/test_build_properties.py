"""Compile-time properties the measured numbers depend on, read from the built library (no GPU needed): the stage kernels
are compiled for sm_100a and keep the register budgets their occupancy was tuned for (primary and
early-bounce trace 64 registers = 4 blocks of 256 threads per SM, late-bounce trace 80 = 3 blocks, shade 64, miss 32),
and the path has no tensor-core / TMA instructions (no stage is a contraction: BASELINE.json north_star)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raytracing_c_b200", "csrc", "libraytracer_gpu.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs the built library and cuobjdump")

BUDGET = {   # mangled-name fragment -> (max registers per thread, max bytes of stack = spills + local arrays)
    "rt_trace_kernelILb1ELi4E": (64, 64),
    "rt_trace_kernelILb0ELi4E": (64, 64),
    "rt_trace_kernelILb0ELi3E": (80, 16),
    "15rt_shade_kernel11": (64, 96),
    "14rt_miss_kernel11": (32, 32),
}


def test_stage_kernels_keep_their_register_budgets():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    found = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        for key in BUDGET:
            if key in m.group(1):
                found[key] = (int(m.group(2)), int(m.group(3)))
    assert set(found) == set(BUDGET), f"kernels missing from the library: {set(BUDGET) - set(found)}"
    for key, (regs, stack) in found.items():
        assert regs <= BUDGET[key][0], f"{key}: {regs} registers (budget {BUDGET[key][0]}) — occupancy would drop"
        assert stack <= BUDGET[key][1], f"{key}: {stack} bytes of stack (budget {BUDGET[key][1]}) — spills in the hot loops"


def test_library_is_sm_100a_without_tensor_or_tma_instructions():
    elfs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in elfs and "sm_90" not in elfs
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert len(sass) > 100000
    assert not re.search(r"UTMALDG|UTCMMA|UTCHMMA|LDTM|HMMA|WGMMA", sass)
