"""Algorithmic work of the DINOSeg hot path: multiply-add = 2 flops; softmax,
GELU and LayerNorm are not counted.  Used by bench.py for the roofline figures."""
from __future__ import annotations


def flops_per_frame(cfg: dict, resolution: int) -> float:
    g = resolution // 8
    p, n = g * g, g * g + 1
    d, hid, c = cfg["embed_dim"], cfg["mlp_hidden"], cfg["n_classes"]
    f = 2.0 * p * 192 * d                                                        # patch embed
    f += cfg["n_blocks"] * (2.0 * n * d * (3 * d + d + 2 * hid) + 4.0 * n * n * d)  # qkv+proj+mlp, QK^T+PV
    f += 2.0 * p * (d * cfg["head_h1"] + cfg["head_h1"] * cfg["head_h2"] + cfg["head_h2"] * c)
    return f


def attention_flops_per_launch(batch: int, n_tokens: int, embed_dim: int) -> float:
    """QK^T + PV of one block for `batch` frames: 2 * (2 * N^2 * dh) per head, H*dh = D."""
    return 4.0 * batch * float(n_tokens) * n_tokens * embed_dim


def gemm_flops(m: int, n: int, k: int) -> float:
    return 2.0 * m * n * k
