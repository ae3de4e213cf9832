"""Oracle (test infrastructure): regression target + loss of the MOFO pretraining step.

Restates ``/root/reference/engine_for_pretraining.py:258-304`` (``train_one_epoch_BB``)
in torch fp32 on CPU, with the patchify index map written out explicitly instead of
through einops so the map itself is what is tested:

    token   n = t*(H'*W') + h*W' + w            (t over T/2 slabs, h,w over 14x14)
    pixel   p = p0*256 + p1*16 + p2             (p0 in {0,1}, p1,p2 in 0..15)
    feature f = p*3 + c                         (channel fastest; engine...:276)
    video[b, c, 2t+p0, 16h+p1, 16w+p2]  ->  patch[b, n, p, c]

The dead ``video_masks`` work (engine...:243-249,278-279,288) has no effect
on loss or gradients and is not restated.
"""
from __future__ import annotations

import numpy as np
import torch

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)   # timm.data.constants (engine...:8)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def patchify(x: torch.Tensor, tubelet: int = 2, patch: int = 16) -> torch.Tensor:
    """'b c (t p0) (h p1) (w p2) -> b (t h w) (p0 p1 p2) c'   (engine...:268)."""
    B, C, T, H, W = x.shape
    t, h, w = T // tubelet, H // patch, W // patch
    x = x.reshape(B, C, t, tubelet, h, patch, w, patch)
    x = x.permute(0, 2, 4, 6, 3, 5, 7, 1)          # b t h w p0 p1 p2 c
    return x.reshape(B, t * h * w, tubelet * patch * patch, C)


def build_labels(videos: torch.Tensor, mask: torch.Tensor, normalize_target: bool = True,
                 patch: int = 16) -> torch.Tensor:
    """videos f32 [B,3,T,H,W] (ImageNet-normalised), mask bool [B,N] -> labels f32 [B,N_mask,1536].
    engine_for_pretraining.py:258-288."""
    mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN, dtype=videos.dtype)[None, :, None, None, None]
    std = torch.as_tensor(IMAGENET_DEFAULT_STD, dtype=videos.dtype)[None, :, None, None, None]
    unnorm = videos * std + mean                                           # :265
    sq = patchify(unnorm, 2, patch)                                        # :268
    if normalize_target:
        sq = (sq - sq.mean(dim=-2, keepdim=True)) / (
            sq.var(dim=-2, unbiased=True, keepdim=True).sqrt() + 1e-6)     # :269-270
    vp = sq.reshape(sq.shape[0], sq.shape[1], -1)                          # :276 'b n p c -> b n (p c)'
    B, _, C = vp.shape
    return vp[mask].reshape(B, -1, C)                                      # :285-286


def mse_loss(outputs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """MSELoss(reduction='none')(out, labels).mean()   (engine...:226,301-304)."""
    return ((outputs.float() - labels) ** 2).mean()


def synthetic_clip(batch: int, seed: int, frames: int = 16, size: int = 224) -> torch.Tensor:
    """SURVEY §8d synthetic input: u~U[0,1) fp32, ImageNet-normalised; CPU generator."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(batch, 3, frames, size, size, generator=g, dtype=torch.float32)
    mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN)[None, :, None, None, None]
    std = torch.as_tensor(IMAGENET_DEFAULT_STD)[None, :, None, None, None]
    return (u - mean) / std


def synthetic_boxes(batch: int, seed: int, frames: int = 16, size: int = 224) -> np.ndarray:
    """SURVEY §8d synthetic motion boxes: integer w,h~U{32..160}, one box replicated over frames."""
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, frames, 4), dtype=np.float64)
    lo, hi = (32, 161) if size >= 224 else (size // 7, size * 5 // 7 + 1)
    for b in range(batch):
        w = int(rng.integers(lo, hi)); h = int(rng.integers(lo, hi))
        x1 = int(rng.integers(0, size - w + 1)); y1 = int(rng.integers(0, size - h + 1))
        out[b, :] = (x1, y1, x1 + w, y1 + h)
    return out
