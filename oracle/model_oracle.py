"""Oracle (test infrastructure): MOFO/VideoMAE pretraining model forward, torch fp32 on CPU
(or any torch device), written functionally over a reference-schema ``state_dict``.

Restates:
  * PatchEmbed.forward                       modeling_finetune.py:242-248
  * get_sinusoid_encoding_table              modeling_finetune.py:252-262
  * Attention.forward / Mlp.forward / Block  modeling_finetune.py:78-98, 44-51, 216-219
  * PretrainVisionTransformerEncoder         modeling_pretrain.py:83-96
  * PretrainVisionTransformerDecoder         modeling_pretrain.py:152-161
  * PretrainVisionTransformer.forward        modeling_pretrain.py:253-266
  * registry configs                         modeling_pretrain.py:268-338

Autograd gives the oracle's gradients; ``pretrain_step`` = forward + target + MSE
(engine_for_pretraining.py:258-304).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import target_oracle


@dataclass(frozen=True)
class Config:
    name: str
    enc_dim: int
    enc_depth: int
    enc_heads: int
    dec_dim: int
    dec_heads: int
    dec_depth: int = 4            # run_mae_pretraining_BB.py --decoder_depth default
    mlp_ratio: int = 4
    img: int = 224
    patch: int = 16
    tubelet: int = 2
    frames: int = 16
    num_classes: int = 1536       # 3 * tubelet * patch^2, modeling_pretrain.py:112

    @property
    def grid(self):
        return (self.frames // self.tubelet, self.img // self.patch, self.img // self.patch)

    @property
    def num_patches(self):
        t, h, w = self.grid
        return t * h * w


CONFIGS = {
    # modeling_pretrain.py:268-338
    "pretrain_mae_small_patch16_224": Config("pretrain_mae_small_patch16_224", 384, 12, 6, 192, 3),
    "pretrain_videomae_base_patch16_224": Config("pretrain_videomae_base_patch16_224", 768, 12, 12, 384, 6),
    "pretrain_videomae_large_patch16_224": Config("pretrain_videomae_large_patch16_224", 1024, 24, 16, 512, 8),
}


def tiny_config(enc_dim=128, enc_depth=2, enc_heads=2, dec_dim=64, dec_heads=1, dec_depth=1,
                img=64, frames=4) -> Config:
    """A scaled-down architecture of the same family (head_dim 64) for fast parity tests."""
    return Config("tiny", enc_dim, enc_depth, enc_heads, dec_dim, dec_heads, dec_depth, img=img, frames=frames)


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """modeling_finetune.py:252-262 — f64 numpy then FloatTensor, shape [1,n,d]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.FloatTensor(ang).unsqueeze(0)


def param_shapes(cfg: Config) -> dict:
    """The 218-tensor schema of SURVEY §8a-13 (name -> shape), in reference registration order."""
    D, Dd = cfg.enc_dim, cfg.dec_dim
    out = {"mask_token": (1, 1, Dd),
           "encoder.patch_embed.proj.weight": (D, 3, cfg.tubelet, cfg.patch, cfg.patch),
           "encoder.patch_embed.proj.bias": (D,)}

    def block(prefix, d):
        h = d * cfg.mlp_ratio
        return {f"{prefix}.norm1.weight": (d,), f"{prefix}.norm1.bias": (d,),
                f"{prefix}.attn.q_bias": (d,), f"{prefix}.attn.v_bias": (d,),
                f"{prefix}.attn.qkv.weight": (3 * d, d),
                f"{prefix}.attn.proj.weight": (d, d), f"{prefix}.attn.proj.bias": (d,),
                f"{prefix}.norm2.weight": (d,), f"{prefix}.norm2.bias": (d,),
                f"{prefix}.mlp.fc1.weight": (h, d), f"{prefix}.mlp.fc1.bias": (h,),
                f"{prefix}.mlp.fc2.weight": (d, h), f"{prefix}.mlp.fc2.bias": (d,)}

    for i in range(cfg.enc_depth):
        out.update(block(f"encoder.blocks.{i}", D))
    out["encoder.norm.weight"] = (D,)
    out["encoder.norm.bias"] = (D,)
    for i in range(cfg.dec_depth):
        out.update(block(f"decoder.blocks.{i}", Dd))
    out["decoder.norm.weight"] = (Dd,)
    out["decoder.norm.bias"] = (Dd,)
    out["decoder.head.weight"] = (cfg.num_classes, Dd)
    out["decoder.head.bias"] = (cfg.num_classes,)
    out["encoder_to_decoder.weight"] = (Dd, D)
    return out


def random_state_dict(cfg: Config, seed: int = 0, perturb: float = 0.0) -> dict:
    """Synthetic weights with the reference's init *distributions* (modeling_pretrain.py:60-67,
    129-136,234; Conv3d default init).  ``perturb`` > 0 additionally randomises biases / LN
    affine so that parity tests do not pass by symmetry (all-zero biases)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if name == "mask_token":
            t = torch.empty(shape).normal_(0, 0.02, generator=g).clamp_(-0.02, 0.02)
        elif name.endswith("proj.weight") and len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name == "encoder.patch_embed.proj.bias":
            bound = 1.0 / math.sqrt(3 * cfg.tubelet * cfg.patch * cfg.patch)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif len(shape) == 2:
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))      # xavier_uniform
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif ".norm" in name and name.endswith("weight"):
            t = torch.ones(shape) + perturb * torch.randn(shape, generator=g)
        else:
            t = perturb * torch.randn(shape, generator=g)
        sd[name] = t
    return sd


def _block(x, p, prefix, heads):
    """Block.forward with gamma None / DropPath identity (modeling_finetune.py:216-219)."""
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), p[f"{prefix}.norm1.weight"], p[f"{prefix}.norm1.bias"], eps=1e-6)
    B, N, _ = h.shape
    qb, vb = p[f"{prefix}.attn.q_bias"], p[f"{prefix}.attn.v_bias"]
    bias = torch.cat((qb, torch.zeros_like(vb), vb))                      # :82-84
    qkv = F.linear(h, p[f"{prefix}.attn.qkv.weight"], bias)
    qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)          # :85
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)                                         # :88
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)                      # :89-92
    o = (attn @ v).transpose(1, 2).reshape(B, N, -1)                      # :95
    x = x + F.linear(o, p[f"{prefix}.attn.proj.weight"], p[f"{prefix}.attn.proj.bias"])
    h = F.layer_norm(x, (d,), p[f"{prefix}.norm2.weight"], p[f"{prefix}.norm2.bias"], eps=1e-6)
    h = F.linear(h, p[f"{prefix}.mlp.fc1.weight"], p[f"{prefix}.mlp.fc1.bias"])
    h = F.gelu(h)                                                         # nn.GELU (erf)
    h = F.linear(h, p[f"{prefix}.mlp.fc2.weight"], p[f"{prefix}.mlp.fc2.bias"])
    return x + h


def forward(cfg: Config, p: dict, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """PretrainVisionTransformer.forward(x, mask) -> [B, N_mask, 1536]."""
    B = x.shape[0]
    D, Dd = cfg.enc_dim, cfg.dec_dim
    # PatchEmbed: Conv3d k=s=(2,16,16) -> flatten(2).transpose(1,2)    modeling_finetune.py:247
    t = F.conv3d(x, p["encoder.patch_embed.proj.weight"], p["encoder.patch_embed.proj.bias"],
                 stride=(cfg.tubelet, cfg.patch, cfg.patch)).flatten(2).transpose(1, 2)
    t = t + sinusoid_table(cfg.num_patches, D).to(t)                       # modeling_pretrain.py:87
    xv = t[~mask].reshape(B, -1, D)                                        # :90
    for i in range(cfg.enc_depth):
        xv = _block(xv, p, f"encoder.blocks.{i}", cfg.enc_heads)
    xv = F.layer_norm(xv, (D,), p["encoder.norm.weight"], p["encoder.norm.bias"], eps=1e-6)
    xv = F.linear(xv, p["encoder_to_decoder.weight"])                      # :256
    pos = sinusoid_table(cfg.num_patches, Dd).to(xv).expand(B, -1, -1)     # :260
    pos_vis = pos[~mask].reshape(B, -1, Dd)
    pos_msk = pos[mask].reshape(B, -1, Dd)
    xf = torch.cat([xv + pos_vis, p["mask_token"] + pos_msk], dim=1)       # :263
    for i in range(cfg.dec_depth):
        xf = _block(xf, p, f"decoder.blocks.{i}", cfg.dec_heads)
    n_mask = pos_msk.shape[1]
    xf = F.layer_norm(xf[:, -n_mask:], (Dd,), p["decoder.norm.weight"], p["decoder.norm.bias"], eps=1e-6)
    return F.linear(xf, p["decoder.head.weight"], p["decoder.head.bias"])  # :156


def pretrain_step(cfg: Config, p: dict, videos: torch.Tensor, mask: torch.Tensor,
                  normalize_target: bool = True, need_grad: bool = True):
    """One oracle step: labels (no_grad) -> forward -> MSE -> backward.
    Returns (loss float, outputs detached, {name: grad})."""
    params = {k: v.detach().clone().requires_grad_(need_grad) for k, v in p.items()}
    with torch.no_grad():
        labels = target_oracle.build_labels(videos, mask, normalize_target, cfg.patch)
    out = forward(cfg, params, videos, mask)
    loss = target_oracle.mse_loss(out, labels)
    grads = {}
    if need_grad:
        loss.backward()
        grads = {k: v.grad.detach() for k, v in params.items()}
    return float(loss.detach()), out.detach(), grads
