#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, read-only).  It imports the
reference's ``masking_generator``, ``modeling_pretrain`` / ``modeling_finetune`` (through a
minimal in-memory ``timm`` shim: registry, ``trunc_normal_``, ``drop_path``, ``to_2tuple``,
ImageNet constants — timm 0.4.12 contributes no arithmetic on this path) and
restates ``engine_for_pretraining.py:258-304`` with einops exactly as written there.
Inputs and weights come from the seeded generators in ``oracle/`` so that the tests can
rebuild them anywhere; only the reference's OUTPUTS are stored.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.json|*.npz
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def install_timm_shim():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    registry = types.ModuleType("timm.models.registry")
    layers = types.ModuleType("timm.models.layers")
    data = types.ModuleType("timm.data")
    constants = types.ModuleType("timm.data.constants")
    _reg = {}

    def register_model(fn):
        _reg[fn.__name__] = fn
        return fn

    def create_model(name, pretrained=False, **kw):
        kw = {k: v for k, v in kw.items() if v is not None}     # timm 0.4.12 drops None kwargs
        return _reg[name](pretrained=pretrained, **kw)

    def trunc_normal_(t, mean=0., std=1., a=-2., b=2.):
        return torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def drop_path(x, drop_prob=0., training=False):
        if drop_prob == 0. or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
        return x.div(keep) * r

    registry.register_model = register_model
    models.create_model = create_model
    models.registry = registry
    models.layers = layers
    layers.trunc_normal_ = trunc_normal_
    layers.to_2tuple = to_2tuple
    layers.drop_path = drop_path
    constants.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    constants.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    data.constants = constants
    timm.models = models
    timm.data = data
    timm.create_model = create_model
    for m in (timm, models, registry, layers, data, constants):
        sys.modules[m.__name__] = m
    return create_model


def ref_labels_and_loss(videos, mask, outputs=None, patch_size=16):
    """engine_for_pretraining.py:258-304 as written (einops rearrange)."""
    from einops import rearrange
    mean = torch.as_tensor((0.485, 0.456, 0.406))[None, :, None, None, None]
    std = torch.as_tensor((0.229, 0.224, 0.225))[None, :, None, None, None]
    unnorm_videos = videos * std + mean
    videos_squeeze = rearrange(unnorm_videos, 'b c (t p0) (h p1) (w p2) -> b (t h w) (p0 p1 p2) c',
                               p0=2, p1=patch_size, p2=patch_size)
    videos_norm = (videos_squeeze - videos_squeeze.mean(dim=-2, keepdim=True)
                   ) / (videos_squeeze.var(dim=-2, unbiased=True, keepdim=True).sqrt() + 1e-6)
    videos_patch = rearrange(videos_norm, 'b n p c -> b n (p c)')
    B, _, C = videos_patch.shape
    labels = videos_patch[mask].reshape(B, -1, C)
    if outputs is None:
        return labels, None
    loss = torch.nn.MSELoss(reduction='none')(input=outputs, target=labels).mean()
    return labels, loss


MASK_CASES = [
    # (box, seed) — SURVEY §8c KATs (seed 10 = the reference's effective state, transforms.py:139) + extras
    ([60, 40, 160, 180], 10), ([0, 0, 1, 1], 10), ([0, 0, 224, 224], 10),
    ([100.5, 20.25, 130.75, 60.5], 10), ([200, 200, 224, 224], 10), ([56, 56, 168, 168], 10),
    ([60, 40, 160, 180], 0), ([17, 33, 48, 64], 1234), ([16, 16, 32, 32], 7), ([223, 0, 224, 224], 99),
    ([0, 100, 224, 101], 5), ([111.9, 111.9, 112.1, 112.1], 2024),
]

SAMPLE_STRIDE = 9973  # prime stride for sampling flat tensors


def sample(t, n=256):
    f = t.detach().reshape(-1)
    idx = (torch.arange(n, dtype=torch.int64) * SAMPLE_STRIDE) % f.numel()
    return f[idx].double().numpy()


def main():
    create_model = install_timm_shim()
    sys.path.insert(0, REF)
    import masking_generator as ref_mg
    import modeling_pretrain as ref_mp            # registers the models  # noqa: F401
    from oracle import mask_oracle, model_oracle, target_oracle

    # ---------------- mask goldens ----------------
    gen = ref_mg.TubeMaskingGenerator_BB((8, 14, 14), 0.9, 0.75)
    cases = []
    for box, seed in MASK_CASES:
        np.random.seed(seed)
        m = gen(np.tile(np.asarray([box], dtype=np.float64), (16, 1)))
        u8 = m.astype(np.uint8)
        cases.append({"box": box, "seed": seed, "sha1": hashlib.sha1(u8.tobytes()).hexdigest(),
                      "vis_slab0": np.nonzero(u8[:196] == 0)[0].tolist(), "n_masked": int(u8.sum())})
    np.random.seed(10)
    pm = ref_mg.TubeMaskingGenerator((8, 14, 14), 0.9)().astype(np.uint8)
    plain = {"seed": 10, "sha1": hashlib.sha1(pm.tobytes()).hexdigest(),
             "vis_slab0": np.nonzero(pm[:196] == 0)[0].tolist()}
    # a small non-default grid (tiny config: 64x64 clip -> 8x4x4 tokens)
    gen_t = ref_mg.TubeMaskingGenerator_BB((8, 4, 4), 0.9, 0.75)
    tiny = []
    for box, seed in [([10, 10, 40, 40], 3), ([0, 0, 1, 1], 10), ([0, 0, 64, 64], 11)]:
        np.random.seed(seed)
        m = gen_t(np.tile(np.asarray([box], dtype=np.float64), (16, 1))).astype(np.uint8)
        tiny.append({"box": box, "seed": seed, "mask": m.tolist()})
    words10 = np.random.RandomState(10)._bit_generator.random_raw(8).tolist()
    with open(os.path.join(HERE, "mask_golden.json"), "w") as f:
        json.dump({"bb_cases": cases, "plain": plain, "tiny_grid_cases": tiny,
                   "mt19937_seed10_first8": words10,
                   "generator": "TubeMaskingGenerator_BB((8,14,14),0.9,0.75) under np.random.seed(seed)"}, f, indent=1)

    # ---------------- target goldens ----------------
    vid = target_oracle.synthetic_clip(2, seed=777)
    words = mask_oracle.mt19937_words(10, 600)
    boxes = target_oracle.synthetic_boxes(2, seed=4321)
    masks = np.stack([mask_oracle.tube_mask_bb(boxes[b], words)[0] for b in range(2)])
    mask_t = torch.from_numpy(masks).to(torch.bool)
    labels, _ = ref_labels_and_loss(vid, mask_t)
    np.savez_compressed(os.path.join(HERE, "target_golden.npz"),
                        labels_sample=sample(labels, 4096), labels_shape=np.asarray(labels.shape),
                        labels_sum=np.asarray(labels.double().sum().item()),
                        labels_abs_sum=np.asarray(labels.double().abs().sum().item()),
                        clip_seed=np.asarray(777), box_seed=np.asarray(4321), mask_seed=np.asarray(10))

    # ---------------- model goldens ----------------
    out = {}
    for tag, cfg, B, ref_kwargs in [
        ("tiny", model_oracle.tiny_config(img=64, frames=16), 2,
         dict(img_size=64, patch_size=16, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2,
              encoder_num_classes=0, decoder_num_classes=1536, decoder_embed_dim=64, decoder_depth=1,
              decoder_num_heads=1, mlp_ratio=4, qkv_bias=True)),
        ("vit_s", model_oracle.CONFIGS["pretrain_mae_small_patch16_224"], 1, None),
        ("vit_b", model_oracle.CONFIGS["pretrain_videomae_base_patch16_224"], 1, None),   # the headline configuration
    ]:
        from functools import partial
        if ref_kwargs is None:
            model = create_model(cfg.name, pretrained=False, drop_path_rate=0.0, drop_block_rate=None,
                                 decoder_depth=cfg.dec_depth)
        else:
            model = ref_mp.PretrainVisionTransformer(
                norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **ref_kwargs)
        sd = model_oracle.random_state_dict(cfg, seed=42, perturb=0.05)
        assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict schema mismatch"
        model.load_state_dict(sd, strict=True)
        model.train()
        vid = target_oracle.synthetic_clip(B, seed=100 + B, size=cfg.img)
        boxes = target_oracle.synthetic_boxes(B, seed=200 + B, size=cfg.img)
        t, h, w = cfg.grid
        masks = np.stack([mask_oracle.tube_mask_bb(boxes[b], mask_oracle.mt19937_words(10 + b, 600),
                                                   (t, h, w))[0] for b in range(B)])
        mask_t = torch.from_numpy(masks).to(torch.bool)
        outputs = model(vid, mask_t)
        _, loss = ref_labels_and_loss(vid, mask_t, outputs)
        loss.backward()
        gn = {k: float(p.grad.double().norm()) for k, p in model.named_parameters()}
        out[f"{tag}_loss"] = np.asarray(float(loss))
        out[f"{tag}_out_shape"] = np.asarray(outputs.shape)
        out[f"{tag}_out_sample"] = sample(outputs, 2048)
        out[f"{tag}_grad_names"] = np.asarray(list(gn.keys()))
        out[f"{tag}_grad_norms"] = np.asarray(list(gn.values()))
        for k in ("mask_token", "encoder.patch_embed.proj.weight", "encoder.blocks.0.attn.qkv.weight",
                  "encoder.blocks.1.mlp.fc1.bias", "decoder.blocks.0.attn.q_bias", "decoder.head.weight",
                  "encoder_to_decoder.weight", "decoder.norm.weight"):
            out[f"{tag}_gsample::{k}"] = sample(dict(model.named_parameters())[k].grad, 128)
        print(tag, "loss", float(loss), "out", tuple(outputs.shape))
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
