#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED REFERENCE (from /root/reference).

Run in the build container only (`python oracle/make_golden.py`); the fixtures are committed
because the reference cannot travel to the GPU box.  Each fixture stores the case description
(enough to regenerate the identical synthetic weights / frames with dino_b200.synthetic) and
what the reference computed for it:

    logprobs      DINOSeg.forward(x)                      (pl_torch_modules.py:239-256)
    low / high    argmax + np.kron as predict() does      (pl_torch_modules.py:294-298)
    pos_rows      rows of interpolate_pos_encoding        (vision_transformer.py:202-222)
    tok_rows, blk0_rows, norm_rows : sampled token rows of the residual stream after
                  prepare_tokens, after block 0 and after the final norm (stage-level pins)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dino_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, arch, n_blocks, res, batch, variant, seed
    ("s8_nb1_240_refinit", "vit_small", 1, 240, 1, "reference_init", 0),
    ("s8_nb3_480_trained", "vit_small", 3, 480, 1, "trained_like", 1),
    ("s8_nb3_240_b2_trained", "vit_small", 3, 240, 2, "trained_like", 2),
    ("s8_nb2_224_trained", "vit_small", 2, 224, 1, "trained_like", 3),
    ("s8_nb1_64_trained", "vit_small", 1, 64, 3, "trained_like", 4),
    ("s8_nb3_480_refinit", "vit_small", 3, 480, 1, "reference_init", 5),
    ("b8_nb4_240_trained", "vit_base", 4, 240, 1, "trained_like", 6),
    ("b8_nb4_240_refinit", "vit_base", 4, 240, 1, "reference_init", 8),
    ("s8_nb1_240_linear5_refinit", "vit_small", 1, 240, 2, "reference_init", 9),      # head='linear', 5 classes
    # BASELINE.json configs 4 and 5 at their full per-frame shape (14401 tokens / ViT-B at 3601 tokens); two frames each,
    # which are also frames 0 and 1 of the batch-16 / batch-32 runs of the GPU tests (make_frames fills frame by frame)
    ("s8_nb3_960_b2_refinit", "vit_small", 3, 960, 2, "reference_init", 12),
    ("b8_nb4_480_b2_refinit", "vit_base", 4, 480, 2, "reference_init", 13),
]


def sample_rows(n_tok: int, k: int = 24) -> np.ndarray:
    """Deterministic sample of token rows: the first rows, the last rows and a spread."""
    base = list(range(min(4, n_tok))) + list(range(max(0, n_tok - 3), n_tok))
    spread = np.linspace(0, n_tok - 1, k).round().astype(int).tolist()
    return np.array(sorted(set(base + spread)), dtype=np.int64)


def build_reference_model(plm, vt, cfg, sd):
    """The reference's own modules, loaded with our synthetic state_dict."""
    if cfg["arch"] == "vit_small":
        m = plm.DINOSeg(data_path="d", write_path="w", head=cfg["head"], n_blocks=cfg["n_blocks"],
                        n_classes=cfg["n_classes"], random_init=True)
    else:
        # the reference hard-codes ViT-S in DINOSeg; ViT-B is composed from its own parts
        # (vit_base vision_transformer.py:307-311 + MLP pl_torch_modules.py:108-124), wired as
        # DINOSeg.forward does
        m = plm.DINOSeg(data_path="d", write_path="w", head="mlp", n_blocks=1, n_classes=cfg["n_classes"],
                        random_init=True)
        dino = vt.vit_base(patch_size=8, num_classes=0)
        dino.blocks = dino.blocks[:cfg["n_blocks"]]
        m.dino = dino
        m.clf = plm.MLP(cfg["n_classes"], input_dim=768)
        m.n_blocks = cfg["n_blocks"]
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def run_case(plm, vt, name, arch, n_blocks, res, batch, variant, seed):
    linear = "linear" in name
    n_classes = 5 if linear else 7
    cfg = synthetic.make_config(arch, n_blocks, n_classes, head="linear" if linear else "mlp")
    sd = synthetic.init_state_dict(cfg, seed, variant)
    x = synthetic.make_frames(batch, res, seed)
    m = build_reference_model(plm, vt, cfg, sd)
    m.set_resolution(res)
    g = res // 8
    with torch.no_grad():
        big = res >= 480          # one frame at a time: the reference materialises [B, H, N, N] attention matrices
        lp = torch.cat([m(x[b:b + 1]) for b in range(batch)], dim=0) if big else m(x)
        tokens = m.dino.prepare_tokens(x)
        blk0 = torch.cat([m.dino.blocks[0](tokens[b:b + 1]) for b in range(batch)], dim=0) if big else m.dino.blocks[0](tokens)
        normed = torch.cat([m.dino(x[b:b + 1]) for b in range(batch)], dim=0) if big else m.dino(x)
        pos = m.dino.interpolate_pos_encoding(tokens, res, res)[0].detach()
        # CLS row of the last block's attention (vision_transformer.py:273-280), small cases only (N x N is materialised)
        cls_attn = m.dino.get_last_selfattention(x)[:, :, 0, :].numpy().astype(np.float32) if res <= 240 else np.zeros(0, np.float32)
    low = torch.argmax(lp, dim=-1).cpu().numpy().reshape(batch, g, g)
    p = 480 // g
    high = np.stack([np.kron(low[b], np.ones((p, p), dtype=int)) for b in range(batch)])
    rows = sample_rows(g * g + 1)
    meta = dict(name=name, arch=arch, n_blocks=n_blocks, res=res, batch=batch, variant=variant, seed=seed,
                n_classes=n_classes, head=cfg["head"], torch=torch.__version__)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=json.dumps(meta),
        logprobs=lp.numpy().astype(np.float32),
        low=low.astype(np.uint8),
        high_shape=np.array(high.shape, dtype=np.int64),
        high_checksum=np.array([int(high.sum()), int((high * np.arange(high.size).reshape(high.shape) % 1000003).sum())],
                               dtype=np.int64),
        rows=rows,
        pos_rows=pos[rows].numpy().astype(np.float32),
        tok_rows=tokens[:, rows].numpy().astype(np.float32),
        blk0_rows=blk0[:, rows].numpy().astype(np.float32),
        norm_rows=normed[:, rows].numpy().astype(np.float32),
        cls_attn=cls_attn,
    )
    hist = np.bincount(low.reshape(-1), minlength=n_classes).tolist()
    print(f"{name}: logprobs {tuple(lp.shape)} range [{lp.min():.3f},{lp.max():.3f}] label hist {hist}")


def survey_anchor(plm):
    """Known-answer anchor recorded in SURVEY.md §7.2-1 for the shimmed reference."""
    torch.manual_seed(0)
    m = plm.DINOSeg(data_path="d", write_path="w", head="mlp", n_blocks=3, n_classes=7, random_init=True)
    x = torch.randn(2, 3, 480, 480)
    with torch.no_grad():
        lp = torch.cat([m(x[b:b + 1]) for b in range(2)], dim=0)
    hist = np.bincount(lp.argmax(1).numpy(), minlength=7).tolist()
    print("survey anchor label histogram:", hist)
    return hist


def predict_case(plm, vt):
    """DINOSeg.predict on a synthetic uint8 image (preprocessing through the albumentations stub:
    pins everything AFTER the transforms; the transforms themselves are unpinned)."""
    from PIL import Image
    cfg = synthetic.make_config("vit_small", 1, 7)
    sd = synthetic.init_state_dict(cfg, 7, "trained_like")
    m = build_reference_model(plm, vt, cfg, sd)
    img = synthetic.make_image_u8(480, 640, 3)
    out = {}
    for res in (240, 480):
        m.set_resolution(res)
        pred = m.predict(Image.fromarray(img))
        assert pred.shape == (480, 480) and pred.dtype == np.int64
        out[f"pred_{res}"] = pred.astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "predict_s8_nb1.npz"),
                        meta=json.dumps(dict(arch="vit_small", n_blocks=1, seed=7, variant="trained_like",
                                             image_seed=3, n_classes=7)), **out)
    print("predict case: ", {k: np.bincount(v.reshape(-1), minlength=7).tolist() for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    plm, vt = ref_shims.import_reference()
    hist = survey_anchor(plm)
    with open(os.path.join(OUT, "survey_anchor.json"), "w") as f:
        json.dump({"label_histogram": hist, "expected_in_SURVEY": [80, 2, 956, 93, 344, 42, 5683]}, f)
    only = set(sys.argv[1:])          # optional: regenerate only the named cases
    for c in CASES:
        if not only or c[0] in only:
            run_case(plm, vt, *c)
    if not only or "predict" in only:
        predict_case(plm, vt)


if __name__ == "__main__":
    main()
