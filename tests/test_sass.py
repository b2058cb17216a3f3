"""CPU-only static checks on the SASS of the built library (cuobjdump; skipped when the CUDA toolkit is absent)."""
import functools
import re
import shutil
import subprocess

import pytest

from conformer_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@functools.lru_cache(maxsize=1)
def _kernels():
    try:
        out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    name, body, kernels = None, [], {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                kernels[name] = body
            name, body = m.group(1), []
        elif name and re.search(r"/\*[0-9a-f]{4}\*/", line):
            body.append(line)
    if name:
        kernels[name] = body
    return kernels


def test_nothing_touches_global_memory_before_the_dependency_wait():
    """Programmatic dependent launch (common.cuh): a kernel's CTAs may start while the previous kernel still runs, so no
    global load, store, atomic or TMA transfer may be issued before `griddepcontrol.wait` (SASS: ACQBULK).  Loads of
    kernel PARAMETERS are the one exception (conv_1 stages its bias first): weights are never written by a kernel of the
    same chain."""
    kernels = _kernels()
    waiting = {n: b for n, b in kernels.items() if any("ACQBULK" in ln for ln in b)}
    assert len(waiting) >= 40, "the forward and training kernels are expected to carry griddepcontrol.wait"
    touch = re.compile(r"\b(LDG|STG|ATOMG|REDG|RED\.E|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTMAPF)\b")
    for name, body in waiting.items():
        first_wait = next(i for i, ln in enumerate(body) if "ACQBULK" in ln)
        early = [ln.strip() for ln in body[:first_wait] if touch.search(ln) and "UTMAPF" not in ln]   # UTMAPF: descriptor prefetch
        if "conv1_tc_kernel" in name:
            assert len(early) <= 1, early                      # the bias (a parameter) -> shared memory
            continue
        assert not early, f"{name}: global access before griddepcontrol.wait: {early[:3]}"


def test_tensor_core_kernels_use_shared_not_generic_addressing():
    """The dynamic shared-memory base is aligned by pointer arithmetic (smem_align_1024): C++ accesses to staging buffers
    must compile to LDS / STS.  A generic LD.E / ST.E in a tcgen05 kernel means the address space got lost again
    (round 2: -4 % on the whole forward)."""
    kernels = _kernels()
    tc = {n: b for n, b in kernels.items() if any("UTCHMMA" in ln for ln in b)}
    assert len(tc) >= 20
    for name, body in tc.items():
        generic = [ln.strip() for ln in body if re.search(r"\b(LD|ST)\.E", ln)]
        assert not generic, f"{name}: generic shared-memory access: {generic[:3]}"
