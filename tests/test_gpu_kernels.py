"""Kernel-level parity on a real B200: every CUDA kernel of libdinoseg.so, called through its
C-ABI op entry point, against plain fp32 torch ops on the same tensors (tools/gpu_check.py holds
the checks; bf16 outputs: 2 % of the reference's max-abs, fp32 outputs: 0.2 %, integer/byte work
bit-exact)."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_check  # noqa: E402

pytestmark = pytest.mark.gpu

NAMES = list(gpu_check.check_names())


@pytest.mark.parametrize("name", NAMES)
def test_kernel(name):
    assert gpu_check.run_check(name)
