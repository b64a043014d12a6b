#!/usr/bin/env python
"""Per-role cycle breakdown of the GEMM kernel (library built with -DDSG_GEMM_TIMING)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

so = os.environ.get("DSG_TIMING_SO", os.path.join(ROOT, "tools", "ubench", "libdinoseg_gtiming.so"))
PROD = ["A refill (wait a_empty + issue)", "wait empty", "issue TMA", "-", "-", "-", "-", "loop"]
MMA = ["wait acc_empty", "wait full", "issue MMAs+commit", "-", "-", "-", "-", "loop + wait a_full"]
EPI = ["wait acc_full", "tmem ld", "bias", "leader wait_read/add", "bar1 / wait addend", "math + st.shared", "fence + bar2", "loop+store"]


def main():
    lib = C.CDLL(so)
    lib.dinoseg_debug_set_attn_timing.argtypes = [C.c_void_p]
    lib.dinoseg_op_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    M = 64 * 3601
    lib.dinoseg_op_gemm_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_float, C.c_int, C.c_void_p]
    pair = os.environ.get("GEMM_PAIR", "0") == "1"      # qkv through the CTA-pair kernel (leader CTAs carry the MMA role)
    shapes = (("qkv", 1152, 384, 0),) if pair else (("qkv", 1152, 384, 0), ("fc1", 1536, 384, 1), ("proj", 384, 384, 2), ("fc2", 384, 1536, 2))
    for name, N, K, epi in shapes:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi >= 2 else torch.bfloat16)
        timing = torch.zeros(148 * 3 * 8, dtype=torch.int64, device="cuda")
        assert lib.dinoseg_debug_set_attn_timing(timing.data_ptr()) == 0
        def run():
            if pair:
                return lib.dinoseg_op_gemm_pair(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, N, 0, 0.125,
                                                N // 3, None)
            return lib.dinoseg_op_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, N, epi, 0.125,
                                       N // 3, None, 0, 0, None)
        for _ in range(2):
            rc = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        raw = timing.view(148, 3, 8).double().cpu()
        t = raw.mean(0)
        if pair:
            t[1] = raw[0::2, 1].mean(0)                 # only even CTAs (cluster rank 0) issue MMAs
        tiles = ((M + 127) // 128) * ((N + 191) // 192) / 148.0
        print(f"== {name}: rc={rc} {ms:.3f} ms, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s, {tiles:.1f} tiles/CTA, "
              f"{ms * 1e-3 * 1.965e9 / tiles:.0f} clk/tile")
        for role, names in ((0, PROD), (1, MMA), (2, EPI)):
            print("  " + ["producer", "mma", "epilogue(leader)"][role] + ": " +
                  ", ".join(f"{n}={t[role][i].item() / tiles:.0f}" for i, n in enumerate(names) if n != "-"))


if __name__ == "__main__":
    main()
