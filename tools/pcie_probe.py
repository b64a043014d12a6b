#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, all ranks copying at once (run under torchrun, one rank per GPU).

Prints, per rank, the pinned-memory H2D, D2H and simultaneous H2D + D2H bandwidth for buffers of the size one
bench step moves (64 frames: 177 MB in, 118 MB out), once with the process left where the launcher put it and once
bound to the GPU's NUMA-local cores (dino_b200.dist.bind_to_gpu_cpus) with freshly allocated pinned buffers.  The
e2e frames/s of bench.py cannot exceed  min(H2D / 2.76 MB, D2H / 1.84 MB)  per GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dino_b200 import dist as D  # noqa: E402


def measure(tag, rank, reps=10):
    n_in, n_out = 64 * 3 * 480 * 480 * 4, 64 * 480 * 480 * 8
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_event(e0)
        s_out.wait_event(e0)
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_out)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        res[mode] = sec
    h2d, d2h = n_in / res["h2d"] / 1e9, n_out / res["d2h"] / 1e9
    both = 64 / res["both"]
    print(f"[{tag}] rank {rank}: H2D {h2d:.1f} GB/s, D2H {d2h:.1f} GB/s, both directions at once: "
          f"{n_in / res['both'] / 1e9:.1f} + {n_out / res['both'] / 1e9:.1f} GB/s -> copy ceiling {both:.0f} frames/s "
          f"(cores: {len(os.sched_getaffinity(0))})", flush=True)


def main():
    rank, local_rank, world = D.env_world()
    torch.cuda.set_device(local_rank)
    if world > 1:
        D.init("nccl")
    measure("unbound", rank)
    n = D.bind_to_gpu_cpus(local_rank)
    measure(f"bound to {n} local cores", rank)
    D.shutdown()


if __name__ == "__main__":
    main()
