"""Deterministic synthetic weights and frames for DINOSeg (no checkpoint or dataset is
available offline; BASELINE.json specifies random-init weights and synthetic frames).

The parameter names and shapes are exactly the reference's state_dict
(reference pl_torch_modules.py:173-183,219-222, vision_transformer.py:163-191).

Variants
  'reference_init' : the distributions the reference uses under random_init=True
                     (Linear ~ trunc_normal(0.02), bias 0, LayerNorm (1,0), pos/cls ~
                     trunc_normal(0.02): vision_transformer.py:188-200; conv and head keep
                     PyTorch's default uniform init: pl_torch_modules.py:182-183).
  'trained_like'   : larger qkv/proj/mlp weights, non-zero biases and non-trivial LayerNorm
                     affines, so that attention is peaked and every bias / affine path
                     influences the result (used by the parity tests to catch layout bugs that
                     near-uniform attention would hide).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

ARCHS = {
    # reference vision_transformer.py:300-311
    "vit_small": dict(embed_dim=384, num_heads=6, mlp_hidden=1536),
    "vit_base": dict(embed_dim=768, num_heads=12, mlp_hidden=3072),
}


def make_config(arch: str = "vit_small", n_blocks: int = 3, n_classes: int = 7, head: str = "mlp") -> dict:
    cfg = dict(ARCHS[arch])
    cfg.update(arch=arch, n_blocks=n_blocks, n_classes=n_classes, patch=8, pos_grid=28,
               head_h1=200, head_h2=100, head=head, ln_eps=1e-6)
    return cfg


def _tn(gen, shape, std):
    # trunc_normal_(std, a=-2, b=2): the +-2 bounds are absolute, i.e. >= 25 sigma for the
    # stds used here, so a plain normal clipped to [-2, 2] is the same distribution
    return (torch.randn(shape, generator=gen) * std).clamp_(-2.0, 2.0)


def _uni(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound


def init_state_dict(cfg: dict, seed: int = 0, variant: str = "reference_init") -> "OrderedDict[str, torch.Tensor]":
    """fp32 CPU state_dict with the reference's key names."""
    assert variant in ("reference_init", "trained_like")
    g = torch.Generator().manual_seed(1000 + seed)
    D, HID, C = cfg["embed_dim"], cfg["mlp_hidden"], cfg["n_classes"]
    G0, H1, H2 = cfg["pos_grid"], cfg["head_h1"], cfg["head_h2"]
    tl = variant == "trained_like"
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["dino.cls_token"] = _tn(g, (1, 1, D), 0.02 if not tl else 0.2)
    sd["dino.pos_embed"] = _tn(g, (1, G0 * G0 + 1, D), 0.02 if not tl else 0.2)
    fan = 3 * 8 * 8
    sd["dino.patch_embed.proj.weight"] = _uni(g, (D, 3, 8, 8), 1.0 / math.sqrt(fan))
    sd["dino.patch_embed.proj.bias"] = _uni(g, (D,), 1.0 / math.sqrt(fan))

    def ln(prefix):
        if tl:
            sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            sd[prefix + ".bias"] = 0.1 * torch.randn(D, generator=g)
        else:
            sd[prefix + ".weight"] = torch.ones(D)
            sd[prefix + ".bias"] = torch.zeros(D)

    def lin(prefix, out_f, in_f, std):
        sd[prefix + ".weight"] = _tn(g, (out_f, in_f), std)
        sd[prefix + ".bias"] = 0.05 * torch.randn(out_f, generator=g) if tl else torch.zeros(out_f)

    for i in range(cfg["n_blocks"]):
        p = f"dino.blocks.{i}."
        ln(p + "norm1")
        # trained_like: std chosen so that the attention logits have the same spread (std ~3) for
        # every embed dim (q.k grows with D * std^4)
        lin(p + "attn.qkv", 3 * D, D, 0.02 if not tl else 0.09 * (384.0 / D) ** 0.5)
        lin(p + "attn.proj", D, D, 0.02 if not tl else 0.04)
        ln(p + "norm2")
        lin(p + "mlp.fc1", HID, D, 0.02 if not tl else 0.05)
        lin(p + "mlp.fc2", D, HID, 0.02 if not tl else 0.03)
    ln("dino.norm")
    layers = (("clf.layer_1", (H1, D)), ("clf.layer_2", (H2, H1)), ("clf.layer_3", (C, H2))) if cfg.get("head", "mlp") == "mlp" \
        else (("clf.layer_1", (C, D)),)      # 'linear' head: pl_torch_modules.py:127-138
    for name, (o, i_) in layers:
        b = 1.0 / math.sqrt(i_)
        sd[name + ".weight"] = _uni(g, (o, i_), b)
        sd[name + ".bias"] = _uni(g, (o,), b)
    return sd


def make_frames(batch: int, resolution: int, seed: int = 1) -> torch.Tensor:
    """Synthetic, already-normalised frames: fp32 [B,3,r,r] ~ N(0,1)."""
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(batch, 3, resolution, resolution, generator=g)


def make_image_u8(height: int = 480, width: int = 640, seed: int = 3):
    """Synthetic RGB uint8 HWC image (smooth blobs + noise) for predict()."""
    import numpy as np
    g = torch.Generator().manual_seed(3000 + seed)
    low = torch.rand(1, 3, 6, 8, generator=g)
    img = torch.nn.functional.interpolate(low, size=(height, width), mode="bilinear", align_corners=False)[0]
    img = img + 0.05 * torch.randn(3, height, width, generator=g)
    return (img.clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy().astype(np.uint8)
