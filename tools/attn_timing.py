#!/usr/bin/env python
"""Phase breakdown of the attention kernel's softmax warpgroups (debug build with -DDSG_ATTN_TIMING).
Builds a separate library under gpurun_out/, runs one attention launch at the bench shape and prints the
average cycles per key tile spent in each phase."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

so = os.environ.get("DSG_TIMING_SO", os.path.join(ROOT, "dino_b200", "lib", "libdinoseg.so"))
names = ["wait s_full", "tmem ld S", "row max", "wait pv_done (+rescale)", "ping-pong wait", "exp phase",
         "st wait + p_full", "loop/epilogue/other"]


def main():
    lib = C.CDLL(so)
    B, N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 3601, 6
    D = H * 64
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 1.0)
    qkv[..., :D] *= 0.125 * 1.4426950408889634
    qkv = qkv.to(torch.bfloat16)
    out = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    timing = torch.zeros(148 * 3 * 8, dtype=torch.int64, device="cuda")
    lib.dinoseg_debug_set_attn_timing.argtypes = [C.c_void_p]
    lib.dinoseg_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    have_timing = lib.dinoseg_debug_set_attn_timing(timing.data_ptr()) == 0
    for _ in range(3):
        lib.dinoseg_op_attention(qkv.data_ptr(), out.data_ptr(), B, N, H, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1 if have_timing else 10
    e0.record()
    for _ in range(reps):
        lib.dinoseg_op_attention(qkv.data_ptr(), out.data_ptr(), B, N, H, None)
    e1.record()
    torch.cuda.synchronize()
    if not have_timing:
        ms = e0.elapsed_time(e1) / reps
        print(f"kernel {ms:.3f} ms  ({4.0 * B * N * N * D / ms / 1e9:.0f} TFLOP/s)")
        return
    tm = timing[148 * 16:].view(148, 8).double().cpu().mean(0)
    t = timing[:148 * 16].view(148, 2, 8).double().cpu()
    q_tiles = (N + 127) // 128
    items = B * H * (q_tiles // 2) + ((B * H + 1) // 2 if q_tiles & 1 else 0)
    tiles_per_wg = items / 148.0 * ((N + 127) // 128)
    print(f"kernel {e0.elapsed_time(e1):.3f} ms; ~{tiles_per_wg:.0f} key tiles per warpgroup per CTA")
    for wg in range(2):
        tot = t[:, wg, :].mean(0)
        print(f"warpgroup {wg}: total {tot.sum().item() / tiles_per_wg:.0f} clk/tile")
        for i, n in enumerate(names):
            print(f"   {n:28s} {tot[i].item() / tiles_per_wg:8.1f} clk/tile")
    mn = ["wait kv_full", "wait s_empty0", "wait s_empty1", "wait p_full0", "wait p_full1", "-", "issue MMAs", "other"]
    print(f"MMA thread: total {tm.sum().item() / tiles_per_wg:.0f} clk/iteration")
    for i, n in enumerate(mn):
        print(f"   {n:28s} {tm[i].item() / tiles_per_wg:8.1f} clk/iteration")


if __name__ == "__main__":
    main()
