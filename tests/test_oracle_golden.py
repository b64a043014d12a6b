"""The oracle (oracle/dinoseg_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, GOLDEN_CASES, load_golden
from dino_b200 import synthetic
from oracle import dinoseg_oracle as O

# the oracle re-states the reference with the same ATen CPU kernels; only summation order /
# batching may differ
TOL = 2e-5


def _case(name):
    gd = load_golden(name)
    m = gd["meta"]
    cfg = synthetic.make_config(m["arch"], m["n_blocks"], m["n_classes"], head=m.get("head", "mlp"))
    sd = synthetic.init_state_dict(cfg, m["seed"], m["variant"])
    x = synthetic.make_frames(m["batch"], m["res"], m["seed"])
    return gd, m, cfg, sd, x


def test_survey_anchor_reproduced():
    """SURVEY.md §7.2-1 known answer, re-measured by make_golden.py on the shimmed reference."""
    with open(os.path.join(GOLDEN, "survey_anchor.json")) as f:
        a = json.load(f)
    assert a["label_histogram"] == a["expected_in_SURVEY"] == [80, 2, 956, 93, 344, 42, 5683]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference(name):
    gd, m, cfg, sd, x = _case(name)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    stages = {}
    lp = O.forward(sd, cfg, x, stages=stages)
    g = m["res"] // 8
    rows = gd["rows"]
    assert lp.shape == gd["logprobs"].shape
    assert np.abs(lp.numpy() - gd["logprobs"]).max() <= TOL
    pos = O.interpolate_pos_encoding(sd["dino.pos_embed"], g)[0]
    assert np.abs(pos[rows].numpy() - gd["pos_rows"]).max() <= 1e-6
    assert np.abs(stages["tokens"][:, rows].numpy() - gd["tok_rows"]).max() <= TOL
    assert np.abs(stages["block0"][:, rows].numpy() - gd["blk0_rows"]).max() <= 5 * TOL
    low, high = O.labels_from_logprobs(torch.from_numpy(gd["logprobs"]), m["batch"], g)
    assert (low == gd["low"]).all()
    assert list(high.shape) == gd["high_shape"].tolist()
    chk = [int(high.sum()), int((high * np.arange(high.size).reshape(high.shape) % 1000003).sum())]
    assert chk == gd["high_checksum"].tolist()
    # labels computed from the oracle's own log-probs: identical except where the top-2 margin is
    # below the fp32 noise between the two evaluations
    low2, _ = O.labels_from_logprobs(lp, m["batch"], g)
    assert (low2 == gd["low"]).mean() >= 0.999


@pytest.mark.parametrize("g", [8, 28, 30, 60, 120])
def test_bicubic_numpy_restatement_matches_torch(g):
    torch.manual_seed(g)
    pos = torch.randn(1, 785, 48)
    ref = O.interpolate_pos_encoding(pos, g)[0].numpy()
    got = O.bicubic_pos_table_numpy(pos[0].numpy(), g)
    assert got.shape == ref.shape == (g * g + 1, 48)
    assert np.abs(got - ref).max() <= 2e-5  # fp32 summation order differs


def test_argmax_replicate_numpy_restatement():
    rng = np.random.default_rng(0)
    lp = rng.standard_normal((2 * 30 * 30, 7)).astype(np.float32)
    lp[3] = 0.5                      # all-equal row -> index 0
    lp[5, 2] = lp[5, 6] = 7.0        # tie -> first
    lp[8, 4] = np.nan                # NaN counts as max
    lp[9, 1] = np.nan; lp[9, 3] = np.nan
    low_t, high_t = O.labels_from_logprobs(torch.from_numpy(lp), 2, 30)
    low_n, high_n = O.argmax_replicate_numpy(lp, 2, 30)
    assert (low_t == low_n).all() and (high_t == high_n).all()
    assert high_n.shape == (2, 480, 480) and high_n.dtype == np.int64
    assert low_n.reshape(-1)[3] == 0 and low_n.reshape(-1)[5] == 2 and low_n.reshape(-1)[8] == 4 and low_n.reshape(-1)[9] == 1


@pytest.mark.parametrize("res,shape", [(240, 480), (480, 480), (448, 448), (496, 434), (64, 480)])
def test_output_size_rule(res, shape):
    """README says 'always 480x480'; the code gives g * (480 // g)."""
    g = res // 8
    lp = torch.zeros(g * g, 7)
    _, high = O.labels_from_logprobs(lp, 1, g)
    assert high.shape == (1, shape, shape)


def test_flops_formula_matches_baseline_md():
    cfg = synthetic.make_config("vit_small", 3, 7)
    assert abs(O.flops_per_frame(cfg, 480) / 1e9 - 99.217) < 0.01
    cfg1 = synthetic.make_config("vit_small", 1, 7)
    assert abs(O.flops_per_frame(cfg1, 240) / 1e9 - 4.744) < 0.01
    cfgb = synthetic.make_config("vit_base", 4, 7)
    assert abs(O.flops_per_frame(cfgb, 480) / 1e9 - 365.56) < 0.05


@pytest.mark.parametrize("hw,res", [((480, 640), 480), ((480, 640), 240), ((480, 640), 960), ((600, 800), 480),
                                    ((123, 77), 64), ((480, 480), 240), ((1080, 1920), 480), ((480, 640), 496),
                                    ((480, 480), 480)])
def test_preprocessing_oracle_matches_cv2_and_transforms(hw, res):
    """The integer restatement of OpenCV's 8-bit bilinear resize is bit-exact w.r.t. cv2.resize(INTER_LINEAR)
    (what albumentations.Resize calls), and the full preprocessing equals dino_b200.transforms (the host path)."""
    import cv2
    from oracle import preproc_oracle as P
    from dino_b200.transforms import get_transforms
    rng = np.random.default_rng(hw[0] * 7 + res)
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref = cv2.resize(img, dsize=(res, res), interpolation=cv2.INTER_LINEAR)
    assert (P.resize_linear_u8(img, res) == ref).all()
    t = get_transforms(res)(image=img)["image"].numpy()
    assert np.array_equal(P.preprocess(img, res), t)


@pytest.mark.parametrize("name", ["s8_nb1_240_refinit", "s8_nb3_240_b2_trained", "b8_nb4_240_refinit"])
def test_oracle_cls_attention_matches_reference(name):
    """CLS row of VisionTransformer.get_last_selfattention (vision_transformer.py:273-280) as the reference computed it."""
    gd, m, cfg, sd, x = _case(name)
    got = O.last_selfattention_cls(sd, cfg, x).numpy()
    assert got.shape == gd["cls_attn"].shape
    assert np.abs(got - gd["cls_attn"]).max() <= 1e-6


def test_half_counts_oracle():
    """Left / right class counts: brute force over pixels on a small odd-width map."""
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 5, size=(2, 6, 9))
    got = O.half_counts(lab, 5)
    for b in range(2):
        for side in range(2):
            for c in range(5):
                n = sum(1 for y in range(6) for x in range(9) if lab[b, y, x] == c and (x >= 4) == bool(side))
                assert got[b, side, c] == n
    assert got.sum() == lab.size
