"""CPU checks on the built CUDA code itself (no GPU needed): the library carries sm_100a SASS only, every trace-kernel
variant fits the one-block-of-512-threads-per-SM design (<= 128 registers, static shared memory only: the 16 KB sincos table and the 24 KB exit-state staging tiles, next to no stack), and
the hot loop really runs on the FP64 pipe with the reciprocal seed (DFMA + MUFU.RCP64H, no slow-path division)."""
import os
import re
import shutil
import subprocess

import pytest

from blackhole_geodesic_calculator_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def _run(*args):
    return subprocess.run([CUOBJDUMP, *args, _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout


def test_library_holds_sm100a_code_only():
    elfs = re.findall(r"ELF file\s+\d+:\s+(\S+)", _run("-lelf"))
    assert elfs and all(e.endswith(".sm_100a.cubin") for e in elfs), elfs


def test_trace_kernel_variants_fit_the_launch_design():
    usage = _run("--dump-resource-usage")
    found = re.findall(r"Function (\S*trace_kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert len(found) >= 8, "parity / plane x SoA / AoS / f32 + disk / polyline variants expected"
    for name, reg, stack, shared, local in found:
        # 512 threads x 128 registers = the whole 64 K register file of an SM: one block per SM must fit
        assert int(reg) <= 128 and int(shared) <= 48 * 1024 and int(local) == 0 and int(stack) <= 64, (name, reg, stack)


def test_hot_kernel_uses_the_fp64_pipe_without_slow_paths():
    # the default kernel of the bench: parity mode, float64 AoS, pre-pass, local outputs
    sass = _run("-sass", "-fun", "_ZN3bhg12trace_kernelILi4ELi1ELb0ELb0ELb1ELb0EEEvNS_9TraceArgsE")
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
    count = lambda prefix: sum(o.startswith(prefix) for o in ops)
    assert count("DFMA") > 400 and count("DMUL") > 250 and count("MUFU.RCP64H") >= 8
    assert count("DFMA") + count("DMUL") + count("DADD") > 0.35 * len(ops)   # static share, service path included
    # local memory: at most the one spilled pair of the attempt loop
    assert count("STL") <= 4 and count("LDL") <= 4
