#!/usr/bin/env python
"""qkv GEMM at the bench shape: single-CTA tiles vs CTA pairs (cta_group::2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dino_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    M, N, K = 64 * 3601, 1152, 384
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    p = lambda t: t.data_ptr()
    runs = {
        "single": lambda: lib.dinoseg_op_gemm(p(A), p(W), p(bias), p(out), M, N, K, N, 0, 0.125, N // 3, None, 0, 0, None),
        "pair": lambda: lib.dinoseg_op_gemm_pair(p(A), p(W), p(bias), p(out), M, N, K, N, 0, 0.125, N // 3, None),
    }
    for rep in range(2):
        for name, fn in runs.items():
            for _ in range(3):
                rc = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"qkv GEMM {name}: rc={rc} {ms:.4f} ms, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
