#!/usr/bin/env python
"""Cycle breakdown of the fused MLP kernel (library built with -DDSG_MLP_TIMING)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

so = os.environ.get("DSG_TIMING_SO", os.path.join(ROOT, "dino_b200", "lib", "libdinoseg.so"))
MMA = ["wait w_full", "wait acc1_empty", "wait g_full", "wait a_full", "wait acc2_empty", "-", "-", "issue + other"]
EPI = ["-", "wait acc1_full", "ld + gelu", "wait g_empty", "write G", "-", "-", "other"]


def main():
    lib = C.CDLL(so)
    lib.dinoseg_debug_set_attn_timing.argtypes = [C.c_void_p]
    lib.dinoseg_op_mlp_ex.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]
    mc = int(os.environ.get("DSG_MLP_MC", "0"))
    M = 64 * 3601
    dev = "cuda"
    x = torch.randn(M, 384, device=dev)
    g = torch.ones(384, device=dev); b = torch.zeros(384, device=dev)
    W1 = (torch.randn(1536, 384, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.zeros(1536, device=dev)
    W2 = (torch.randn(384, 1536, device=dev) * 0.03).to(torch.bfloat16); b2 = torch.zeros(384, device=dev)
    timing = torch.zeros(148 * 2 * 8, dtype=torch.int64, device=dev)
    have = lib.dinoseg_debug_set_attn_timing(timing.data_ptr()) == 0
    A = torch.randn(M, 384, device=dev).to(torch.bfloat16)
    args = (x.data_ptr(), A.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, mc, None)
    if os.environ.get("DSG_MLP_LN", "0") != "0":        # the kernel normalising x itself (LayerNorm2 fused)
        lib.dinoseg_op_mlp_ln.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p]
        args = (x.data_ptr(), 1e-6, W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), M, mc, None)
        lib.dinoseg_op_mlp_ex = lib.dinoseg_op_mlp_ln
    for _ in range(2):
        lib.dinoseg_op_mlp_ex(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1 if have else 5
    e0.record()
    for _ in range(reps):
        lib.dinoseg_op_mlp_ex(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    blocks = ((M + 127) // 128) / 148.0
    print(f"mlp_fused mc={mc} {ms:.3f} ms, {4.0 * M * 384 * 1536 / ms / 1e9:.0f} TFLOP/s, {ms * 1e-3 * 1.965e9 / blocks:.0f} clk per row block")
    if have:
        raw = timing.view(148, 2, 8).double().cpu()
        lead = raw[raw[:, 0, 6] > 0]            # CTAs that issued MMAs (every CTA, or the pair leaders)
        print(f"  SM clock inside the kernel: {(lead[:, 0, 5] / lead[:, 0, 6]).mean().item() * 1e3:.0f} MHz "
              f"(cycles / globaltimer ns of the MMA issuer's loop)")
        raw[:, 0, 5:7] = 0
        t = raw.mean(0)
        for role, names in ((0, MMA), (1, EPI)):
            print("  " + ["mma", "gelu warp 2"][role] + ": " +
                  ", ".join(f"{n}={t[role][i].item() / blocks:.0f}" for i, n in enumerate(names) if n != "-"))


if __name__ == "__main__":
    main()
