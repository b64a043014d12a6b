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


@pytest.mark.parametrize("unshifted", ["1", "0"])
def test_attention_shape_sweep(unshifted, monkeypatch):
    """Token counts around every tile boundary (1 .. 1153), odd numbers of (frame, head)s, 1 / 6 / 12 heads: the attention
    kernel without row maxima (default) and its classic twin alone, each against fp32 torch (tools/attn_sweep.py)."""
    import attn_sweep
    monkeypatch.setenv("DINOSEG_ATTN_UNSHIFTED", unshifted)
    assert attn_sweep.main() == 0
