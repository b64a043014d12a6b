#!/usr/bin/env python
"""Where the fused head kernel's MMA issuer waits (debug build with -DDSG_HEAD_TIMING):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --shared -Xcompiler -fPIC -DDSG_HEAD_TIMING \
         -o tools/ubench/libdinoseg_htiming.so dino_b200/csrc/dinoseg_api.cu
    DINOSEG_LIB=tools/ubench/libdinoseg_htiming.so python tools/head_timing.py
Prints the cycles per row block CTA 0's issuer spent waiting on each barrier, and issuing."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dino_b200 import DINOSeg, _lib, synthetic  # noqa: E402


def main():
    cfg = synthetic.make_config("vit_small", 3, 7)
    m = DINOSeg(head="mlp", n_blocks=3, n_classes=7)
    m.load_state_dict(synthetic.init_state_dict(cfg, 0, "reference_init"), strict=True)
    m = m.to("cuda:0")
    m.set_resolution(480)
    x = synthetic.make_frames(64, 480, seed=1).cuda()
    for _ in range(3):
        m.infer(x, want_logprobs=False, want_labels=True)
    torch.cuda.synchronize()
    hb = (C.c_int * 1024)()
    assert _lib.load().dinoseg_debug_heartbeat(hb, 1024) > 0
    t = [hb[900 + 2 * i] & 0xffffffff | (hb[901 + 2 * i] << 32) for i in range(8)]
    blocks = max(1, t[6])
    names = ["wait acc1_empty", "wait full (first k-block of a row block)", "wait full (k-blocks 1..5)", "wait h1_ready",
             "wait acc2_empty", "wait w2_full", None, "issue + everything else"]
    total = sum(v for i, v in enumerate(t) if i != 6)
    print(f"CTA 0: {blocks} row blocks, {total / blocks:.0f} clk per row block")
    for n, v in zip(names, t):
        if n:
            print(f"   {n:44s} {v / blocks:9.0f} clk / row block")


if __name__ == "__main__":
    main()
