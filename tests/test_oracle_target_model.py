"""Pins oracle/target_oracle.py and oracle/model_oracle.py against the reference's outputs stored in
tests/golden/{target,model}_golden.npz (made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import mask_oracle as mo
from oracle import model_oracle as mdl
from oracle import target_oracle as tgt

STRIDE = 9973


def sample(t, n):
    f = t.detach().reshape(-1)
    idx = (torch.arange(n, dtype=torch.int64) * STRIDE) % f.numel()
    return f[idx].double().numpy()


def test_patchify_index_map():
    # probe the explicit map of oracle/target_oracle.py docstring on an arange clip
    B, C, T, H, W = 1, 3, 4, 32, 48
    x = torch.arange(B * C * T * H * W, dtype=torch.float32).reshape(B, C, T, H, W)
    p = tgt.patchify(x)
    hh, ww = H // 16, W // 16
    rng = np.random.default_rng(1)
    for _ in range(500):
        t, h, w = rng.integers(0, T // 2), rng.integers(0, hh), rng.integers(0, ww)
        p0, p1, p2, c = rng.integers(0, 2), rng.integers(0, 16), rng.integers(0, 16), rng.integers(0, 3)
        n = t * hh * ww + h * ww + w
        pix = p0 * 256 + p1 * 16 + p2
        assert p[0, n, pix, c] == x[0, c, 2 * t + p0, 16 * h + p1, 16 * w + p2]


def test_labels_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "target_golden.npz"))
    vid = tgt.synthetic_clip(2, seed=int(g["clip_seed"]))
    boxes = tgt.synthetic_boxes(2, seed=int(g["box_seed"]))
    words = mo.mt19937_words(int(g["mask_seed"]), 600)
    masks = np.stack([mo.tube_mask_bb(boxes[b], words)[0] for b in range(2)])
    labels = tgt.build_labels(vid, torch.from_numpy(masks).to(torch.bool))
    assert list(labels.shape) == g["labels_shape"].tolist()
    np.testing.assert_allclose(sample(labels, 4096), g["labels_sample"], rtol=0, atol=2e-5)
    assert abs(labels.double().abs().sum().item() - float(g["labels_abs_sum"])) < 1e-4 * float(g["labels_abs_sum"])


def test_sinusoid_closed_form():
    tab = mdl.sinusoid_table(1568, 384)[0].double()
    pos, i = 777, 101
    ang = pos / (10000 ** (2 * (i // 2) / 384))
    assert abs(tab[pos, i].item() - np.float32(np.cos(ang))) < 1e-7
    assert abs(tab[pos, 100].item() - np.float32(np.sin(pos / (10000 ** (100 / 384))))) < 1e-7


def test_schema_counts():
    n = {k: sum(int(np.prod(s)) for s in mdl.param_shapes(c).values()) for k, c in mdl.CONFIGS.items()}
    assert len(mdl.param_shapes(mdl.CONFIGS["pretrain_videomae_base_patch16_224"])) == 218
    assert round(n["pretrain_mae_small_patch16_224"] / 1e6, 2) == 24.03
    assert round(n["pretrain_videomae_base_patch16_224"] / 1e6, 2) == 94.21
    assert round(n["pretrain_videomae_large_patch16_224"] / 1e6, 2) == 317.78


@pytest.mark.parametrize("tag", ["tiny", "vit_s", "vit_b"])
def test_model_matches_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "model_golden.npz"))
    if tag == "tiny":
        cfg, B = mdl.tiny_config(img=64, frames=16), 2
    elif tag == "vit_s":
        cfg, B = mdl.CONFIGS["pretrain_mae_small_patch16_224"], 1
    else:
        cfg, B = mdl.CONFIGS["pretrain_videomae_base_patch16_224"], 1      # the headline configuration
    sd = mdl.random_state_dict(cfg, seed=42, perturb=0.05)
    vid = tgt.synthetic_clip(B, seed=100 + B, size=cfg.img)
    boxes = tgt.synthetic_boxes(B, seed=200 + B, size=cfg.img)
    masks = np.stack([mo.tube_mask_bb(boxes[b], mo.mt19937_words(10 + b, 600), cfg.grid)[0] for b in range(B)])
    loss, out, grads = mdl.pretrain_step(cfg, sd, vid, torch.from_numpy(masks).to(torch.bool))
    assert list(out.shape) == g[f"{tag}_out_shape"].tolist()
    assert abs(loss - float(g[f"{tag}_loss"])) < 2e-6 * max(1.0, abs(loss))
    np.testing.assert_allclose(sample(out, 2048), g[f"{tag}_out_sample"], rtol=0, atol=5e-5)
    names = [str(s) for s in g[f"{tag}_grad_names"]]
    assert names == list(sd.keys())
    for nme, ref in zip(names, g[f"{tag}_grad_norms"]):
        mine = float(grads[nme].double().norm())
        assert abs(mine - ref) <= 1e-4 * max(ref, 1e-6) + 1e-9, (nme, mine, ref)
    for k in g.files:
        if k.startswith(f"{tag}_gsample::"):
            nme = k.split("::", 1)[1]
            np.testing.assert_allclose(sample(grads[nme], 128), g[k], rtol=2e-3, atol=1e-7)
