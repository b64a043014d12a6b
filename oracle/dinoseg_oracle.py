"""CPU ORACLE for the DINOSeg inference hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; nothing under `dino_b200/` does, and the product path raises if
its CUDA library is missing instead of falling back to anything in here.

What it is: a restatement, in plain fp32 torch ops on the CPU (plus numpy restatements of the
two pieces whose semantics are easy to get subtly wrong: the bicubic positional table and
argmax + np.kron), of the algorithm in the reference files

    dt_segmentation/src/vision_transformer.py   (PatchEmbed :143-158, interpolate_pos_encoding
                                                 :202-222, prepare_tokens :224-235, Attention
                                                 :68-107, Mlp :49-65, Block :110-140, forward :237-248)
    dt_segmentation/src/pl_torch_modules.py     (MLP head :108-124, DINOSeg.forward :239-256,
                                                 predict tail :294-298)

It works directly on a reference-named state_dict (no nn.Module), so the same tensors can be
handed to the CUDA library.

Pinning: the reference ships no tests or golden vectors, so parity is pinned
against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the unmodified
reference modules from /root/reference (with import shims for the two absent packages), loads
the same synthetic state_dict, runs `DINOSeg.forward` / `predict`, and stores the results in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this oracle against them.
Preprocessing (albumentations Resize/Normalize) is NOT part of the pinned path: albumentations
is absent here, its restatement lives in dino_b200/transforms.py and is "parity unpinned".
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# positional table
# ------------------------------------------------------------------------------------------
def interpolate_pos_encoding(pos_embed: torch.Tensor, g: int) -> torch.Tensor:
    """vision_transformer.py:202-222 for a square g x g patch grid. pos_embed: [1, G0*G0+1, D]."""
    n_src = pos_embed.shape[1] - 1
    g0 = int(math.sqrt(n_src))
    if g * g == n_src:                                   # :205-206
        return pos_embed
    dim = pos_embed.shape[-1]
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    w0 = h0 = g + 0.1                                    # :214
    patch_pos = F.interpolate(
        patch_pos.reshape(1, g0, g0, dim).permute(0, 3, 1, 2),
        scale_factor=(w0 / math.sqrt(n_src), h0 / math.sqrt(n_src)),
        mode="bicubic",
    )
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]   # :220
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def _cubic_coeffs(t: np.ndarray) -> np.ndarray:
    """ATen get_cubic_upsample_coefficients, A = -0.75, evaluated in fp32."""
    a = np.float32(-0.75)
    t = t.astype(np.float32)

    def conv1(x):
        return ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)

    def conv2(x):
        return ((a * x - np.float32(5) * a) * x + np.float32(8) * a) * x - np.float32(4) * a

    return np.stack([conv2(t + np.float32(1)), conv1(t), conv1(np.float32(1) - t), conv2(np.float32(2) - t)], axis=-1)


def bicubic_pos_table_numpy(pos_embed: np.ndarray, g: int) -> np.ndarray:
    """Independent numpy restatement of what F.interpolate(..., scale_factor=(g+0.1)/G0,
    mode='bicubic', align_corners=False) computes for the positional grid.  [G0*G0+1, D] -> [g*g+1, D].

    src = rscale*(dst+0.5)-0.5, rscale = float32(1/scale_factor); taps floor(src)-1..+2 clamped to
    [0,G0-1]; cubic convolution weights with A=-0.75; x taps summed first, then y.
    """
    pos_embed = np.asarray(pos_embed, dtype=np.float32)
    n_src = pos_embed.shape[0] - 1
    g0 = int(math.sqrt(n_src))
    if g * g == n_src:
        return pos_embed.copy()
    d = pos_embed.shape[1]
    grid = pos_embed[1:].reshape(g0, g0, d)
    rscale = np.float32(1.0 / ((g + 0.1) / float(g0)))
    dst = np.arange(g, dtype=np.float32)
    src = rscale * (dst + np.float32(0.5)) - np.float32(0.5)
    i0 = np.minimum(np.floor(src).astype(np.int64), g0 - 1)
    t = np.clip(src - i0.astype(np.float32), 0, 1).astype(np.float32)
    w = _cubic_coeffs(t)                                                   # [g, 4]
    idx = np.clip(i0[:, None] + np.arange(-1, 3)[None, :], 0, g0 - 1)      # [g, 4]
    # x direction first: tmp[y_src, ox, d]
    tmp = np.zeros((g0, g, d), dtype=np.float32)
    for b in range(4):
        tmp += w[None, :, b, None] * grid[:, idx[:, b], :]
    out = np.zeros((g, g, d), dtype=np.float32)
    for a in range(4):
        out += w[:, a, None, None] * tmp[idx[:, a], :, :]
    return np.concatenate([pos_embed[:1], out.reshape(g * g, d)], axis=0)


# ------------------------------------------------------------------------------------------
# backbone
# ------------------------------------------------------------------------------------------
def prepare_tokens(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """vision_transformer.py:224-235 (patch embed :155-158, cls concat, + positional table)."""
    b, _, w, h = x.shape
    assert w == h and w % 8 == 0
    t = F.conv2d(x, sd["dino.patch_embed.proj.weight"], sd["dino.patch_embed.proj.bias"], stride=8)
    t = t.flatten(2).transpose(1, 2)
    cls = sd["dino.cls_token"].expand(b, -1, -1)
    t = torch.cat((cls, t), dim=1)
    return t + interpolate_pos_encoding(sd["dino.pos_embed"], w // 8)


def attention(sd: dict, prefix: str, x: torch.Tensor, num_heads: int, return_attn: bool = False):
    """vision_transformer.py:80-107 with cls_mask=None."""
    b, n, c = x.shape
    hd = c // num_heads
    qkv = F.linear(x, sd[prefix + "qkv.weight"], sd[prefix + "qkv.bias"])
    qkv = qkv.reshape(b, n, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, n, c)
    y = F.linear(y, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])
    return (y, attn) if return_attn else y


def block(sd: dict, i: int, x: torch.Tensor, num_heads: int, eps: float) -> torch.Tensor:
    """vision_transformer.py:122-140 (drop_path = Identity, dropout p = 0)."""
    p = f"dino.blocks.{i}."
    d = x.shape[-1]
    y = attention(sd, p + "attn.", F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps), num_heads)
    x = x + y
    z = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    z = F.linear(z, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    z = F.gelu(z)                                        # nn.GELU() default = exact erf (:50)
    z = F.linear(z, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + z


def backbone(sd: dict, cfg: dict, x: torch.Tensor, stages: dict | None = None) -> torch.Tensor:
    """VisionTransformer.forward(all=True) over the kept blocks (vision_transformer.py:237-246,
    pl_torch_modules.py:177).  `stages`, if given, receives the residual stream after each stage."""
    t = prepare_tokens(sd, x)
    if stages is not None:
        stages["tokens"] = t
    for i in range(cfg["n_blocks"]):
        t = block(sd, i, t, cfg["num_heads"], cfg["ln_eps"])
        if stages is not None:
            stages[f"block{i}"] = t
    d = t.shape[-1]
    return F.layer_norm(t, (d,), sd["dino.norm.weight"], sd["dino.norm.bias"], cfg["ln_eps"])


@torch.no_grad()
def last_selfattention_cls(sd: dict, cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """Row 0 (CLS query) of VisionTransformer.get_last_selfattention (vision_transformer.py:273-280): [B, H, N]."""
    t = prepare_tokens(sd, x)
    nb = cfg["n_blocks"]
    for i in range(nb - 1):
        t = block(sd, i, t, cfg["num_heads"], cfg["ln_eps"])
    p = f"dino.blocks.{nb - 1}."
    d = t.shape[-1]
    y = F.layer_norm(t, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg["ln_eps"])
    _, attn = attention(sd, p + "attn.", y, cfg["num_heads"], return_attn=True)
    return attn[:, :, 0, :]


def head(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """MLP head, pl_torch_modules.py:117-124; 'linear' head (no layer_2 in the state_dict), :135-138."""
    if "clf.layer_2.weight" not in sd:
        return F.log_softmax(F.linear(x, sd["clf.layer_1.weight"], sd["clf.layer_1.bias"]), dim=1)
    x = F.relu(F.linear(x, sd["clf.layer_1.weight"], sd["clf.layer_1.bias"]))
    x = F.relu(F.linear(x, sd["clf.layer_2.weight"], sd["clf.layer_2.bias"]))
    x = F.linear(x, sd["clf.layer_3.weight"], sd["clf.layer_3.bias"])
    return F.log_softmax(x, dim=1)


@torch.no_grad()
def forward(sd: dict, cfg: dict, x: torch.Tensor, stages: dict | None = None, frame_chunk: int = 1) -> torch.Tensor:
    """DINOSeg.forward, pl_torch_modules.py:239-256: [B,3,r,r] -> log-probs [B*P, C].

    Frames are independent, so they are pushed through in chunks (the materialised N x N
    attention of one 480-px frame is already 311 MB per block)."""
    outs = []
    for s in range(0, x.shape[0], frame_chunk):
        st = {} if stages is not None else None
        t = backbone(sd, cfg, x[s:s + frame_chunk], st)[:, 1:]          # drop CLS (:243)
        t = t.reshape((-1, t.shape[-1]))                                # (:253)
        outs.append(head(sd, t))
        if stages is not None:
            for k, v in st.items():
                stages.setdefault(k, []).append(v)
    if stages is not None:
        for k in list(stages.keys()):
            stages[k] = torch.cat(stages[k], dim=0)
    return torch.cat(outs, dim=0)


def labels_from_logprobs(logprobs: torch.Tensor, batch: int, g: int):
    """predict() tail, pl_torch_modules.py:294-298, per frame: argmax -> [g,g] -> np.kron."""
    low = torch.argmax(logprobs, dim=-1).cpu().numpy().reshape((batch, g, g))
    p = 480 // g
    high = np.stack([np.kron(low[b], np.ones((p, p), dtype=int)) for b in range(batch)]) if p > 0 else \
        np.zeros((batch, 0, 0), dtype=int)
    return low, high


def half_counts(labels: np.ndarray, n_classes: int) -> np.ndarray:
    """Controller-side reduction (reference docs/index.html "Controller"): pixels of every
    class in the left (x < W/2) and right half of each label map.  labels: int [B, H, W] -> int64 [B, 2, n_classes]."""
    labels = np.asarray(labels)
    b, _, w = labels.shape
    out = np.zeros((b, 2, n_classes), dtype=np.int64)
    for i in range(b):
        out[i, 0] = np.bincount(labels[i, :, : w // 2].ravel(), minlength=n_classes)[:n_classes]
        out[i, 1] = np.bincount(labels[i, :, w // 2:].ravel(), minlength=n_classes)[:n_classes]
    return out


def argmax_replicate_numpy(logprobs: np.ndarray, batch: int, g: int):
    """Independent restatement of the tail in index form: out[b,y,x] = low[b, y//p, x//p];
    argmax = first maximum, NaN counts as maximum (torch.argmax semantics)."""
    lp = np.asarray(logprobs, dtype=np.float32)
    rows, c = lp.shape
    low = np.zeros(rows, dtype=np.int64)
    for r in range(rows):
        best, idx = lp[r, 0], 0
        for k in range(1, c):
            v = lp[r, k]
            if (v > best) or (np.isnan(v) and not np.isnan(best)):
                best, idx = v, k
        low[r] = idx
    low = low.reshape(batch, g, g)
    p = 480 // g
    yy = np.arange(g * p) // max(p, 1)
    return low, low[:, yy][:, :, yy] if p > 0 else np.zeros((batch, 0, 0), dtype=np.int64)


def flops_per_frame(cfg: dict, resolution: int) -> float:
    """Algorithmic FLOPs: multiply-add = 2; softmax/GELU/LN not counted."""
    g = resolution // 8
    p, n = g * g, g * g + 1
    d, hid, c = cfg["embed_dim"], cfg["mlp_hidden"], cfg["n_classes"]
    f = 2.0 * p * 192 * d
    f += cfg["n_blocks"] * (2.0 * n * d * (3 * d + d + 2 * hid) + 4.0 * n * n * d)
    f += 2.0 * p * (d * cfg["head_h1"] + cfg["head_h1"] * cfg["head_h2"] + cfg["head_h2"] * c)
    return f
