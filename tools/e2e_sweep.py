#!/usr/bin/env python
"""Sweep the host-pipeline chunk size of dinoseg_predict_host (frames/s end to end, pinned host buffers)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dino_b200 import DINOSeg, _lib, synthetic  # noqa: E402


def main():
    cfg = synthetic.make_config("vit_small", 3, 7)
    sd = synthetic.init_state_dict(cfg, 0, "reference_init")
    m = DINOSeg(head="mlp", n_blocks=3, n_classes=7)
    m.load_state_dict(sd)
    m = m.to("cuda:0")
    B = 64
    x = synthetic.make_frames(B, 480, 1).pin_memory()
    out = torch.empty((B, 480, 480), dtype=torch.int64).pin_memory()
    xd = x.cuda()
    for _ in range(3):
        m.infer(xd, want_logprobs=False, want_labels=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        m.infer(xd, want_logprobs=False, want_labels=True)
    torch.cuda.synchronize()
    print(f"device-resident: {B * 10 / (time.perf_counter() - t0):.0f} frames/s")
    lib = _lib.load()
    for chunk in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32, 64]:
        lib.dinoseg_set_host_chunk(m._handle, chunk)
        for _ in range(2):
            m.predict_batch(x, out=out)
        t0 = time.perf_counter()
        for _ in range(10):
            m.predict_batch(x, out=out)
        dt = time.perf_counter() - t0
        print(f"chunk {chunk:3d}: {B * 10 / dt:.0f} frames/s  ({dt * 100:.2f} ms/step)")


if __name__ == "__main__":
    main()
