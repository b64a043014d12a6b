#!/usr/bin/env python
"""Does a CUDA graph of the 26-kernel forward beat stream launches?  (GPU-side launch gaps vs CPU launch cost.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dino_b200 import DINOSeg, synthetic  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = synthetic.make_config("vit_small", 3, 7, head="mlp")
    sd = synthetic.init_state_dict(cfg, 0, "reference_init")
    m = DINOSeg(head="mlp", n_blocks=3, n_classes=7)
    m.load_state_dict(sd)
    m = m.to("cuda:0")
    m.set_resolution(480)
    x = synthetic.make_frames(B, 480, seed=1).cuda()
    for _ in range(3):
        m.infer(x, want_logprobs=False, want_labels=True)
    torch.cuda.synchronize()

    def timed(fn, steps=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            m.infer(x, want_logprobs=False, want_labels=True)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = m.infer(x, want_logprobs=False, want_labels=True)
    for rep in range(3):
        t_stream = timed(lambda: m.infer(x, want_logprobs=False, want_labels=True))
        t_graph = timed(g.replay)
        print(f"B={B}: stream launches {t_stream:.3f} ms/step ({B / t_stream * 1e3:.0f} frames/s), "
              f"graph replay {t_graph:.3f} ms/step ({B / t_graph * 1e3:.0f} frames/s)")


if __name__ == "__main__":
    main()
