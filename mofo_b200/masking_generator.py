"""B200-native look-alike of the reference's ``masking_generator.py`` (tube masking, MOFO motion-box constraint).

``TubeMaskingGenerator_BB(input_size, mask_ratio, mask_ratio_BB).__call__(bb) -> np.float64[T*H*W]`` and
``TubeMaskingGenerator(input_size, mask_ratio).__call__()`` keep the reference contract (masking_generator.py:3-85),
including its random draw: the generator consumes numpy's *legacy global* MT19937 stream exactly as
``np.random.shuffle`` would (same words, same final RNG state), so a pipeline that seeds numpy the way the reference
does (transforms.py:139) gets bit-identical masks.  The sampling itself runs in the per-clip CUDA kernel
``mofo_tube_mask_bb`` (include/mofo_b200.h); there is no CPU implementation here.

``generate_batch`` is the batched device API the training engine / benchmark use: B boxes and B word streams in,
mask + ascending visible / masked index lists out, all resident on the GPU with no host synchronisation.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

# words one clip can consume: two Fisher-Yates passes over <= H*W elements, masked rejection accepts with p > 1/2.
# 14x14 grid consumes ~270 words on average; 4*H*W bounds it with overwhelming margin and the kernel
# reports exhaustion (words_used = -1) instead of reading past the end.
def words_per_clip(height: int, width: int) -> int:
    return max(64, 4 * height * width)


def _global_raw_words(n: int) -> np.ndarray:
    """Next ``n`` raw 32-bit outputs of numpy's global legacy RandomState WITHOUT consuming them."""
    st = np.random.get_state()
    rs = np.random.RandomState()
    rs.set_state(st)
    return rs._bit_generator.random_raw(n).astype(np.uint32)


def _advance_global(n: int) -> None:
    if n > 0:
        np.random.mtrand._rand._bit_generator.random_raw(int(n))


def mt19937_words(seed: int, n: int) -> np.ndarray:
    """The first ``n`` raw words after ``np.random.seed(seed)`` (host-side helper for building rng_words)."""
    return np.random.RandomState(seed)._bit_generator.random_raw(n).astype(np.uint32)


class TubeMaskingGenerator_BB:
    def __init__(self, input_size, mask_ratio, mask_ratio_BB, device="cuda"):
        self.frames, self.height, self.width = input_size
        self.num_patches_per_frame = self.height * self.width
        self.total_patches = self.frames * self.num_patches_per_frame
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_masks = self.frames * self.num_masks_per_frame
        self.mask_ratio = mask_ratio
        self.mask_ratio_BB = mask_ratio_BB
        self.device = device

    def __repr__(self):
        return "Maks: total patches {}, mask patches {}".format(self.total_patches, self.total_masks)

    # ---- batched device API ---------------------------------------------------------------------------
    def generate_batch(self, bb, rng_words, check=False):
        """bb: [B,16,4] / [B,4] boxes (any float/int array or tensor; only frame 0 is read, masking_generator.py:46,55);
        rng_words: uint32 [B,W] raw MT19937 words per clip (numpy array or CUDA int32/uint32 tensor).
        Returns (mask uint8 [B,N], vis_idx int32 [B,N_vis], msk_idx int32 [B,N_mask], words_used int32 [B]) on device."""
        dev = torch.device(self.device)
        if isinstance(bb, torch.Tensor) and bb.is_cuda:            # already on the device: no host round trip
            bb_t = bb.to(torch.float64)
        else:
            bb_t = torch.as_tensor(np.asarray(bb.cpu() if isinstance(bb, torch.Tensor) else bb, dtype=np.float64))
        if bb_t.dim() == 3:
            bb_t = bb_t[:, 0, :]
        bb_t = bb_t.contiguous().to(dev, non_blocking=True)
        if isinstance(rng_words, np.ndarray):
            rng_words = torch.from_numpy(np.ascontiguousarray(rng_words, dtype=np.uint32).view(np.int32)).to(dev, non_blocking=True)
        out = _lib.tube_mask_bb(bb_t, rng_words.contiguous(), (self.frames, self.height, self.width),
                                self.num_masks_per_frame, self.mask_ratio_BB)
        if check and int(out[3].min().item()) < 0:
            raise RuntimeError("tube mask: rng_words exhausted (increase W)")
        return out

    # ---- reference API --------------------------------------------------------------------------------
    def __call__(self, bb):
        W = words_per_clip(self.height, self.width)
        words = _global_raw_words(W)
        bb0 = np.asarray(bb, dtype=np.float64).reshape(-1, 4)[0:1]
        mask, _, _, used = self.generate_batch(bb0, words[None, :])
        used = int(used.item())
        if used < 0:
            raise RuntimeError("tube mask: rng_words exhausted")
        _advance_global(used)                     # leave numpy's global RNG where np.random.shuffle would have
        return mask[0].cpu().numpy().astype(np.float64)


class TubeMaskingGenerator:
    def __init__(self, input_size, mask_ratio, device="cuda"):
        self.frames, self.height, self.width = input_size
        self.num_patches_per_frame = self.height * self.width
        self.total_patches = self.frames * self.num_patches_per_frame
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_masks = self.frames * self.num_masks_per_frame
        self.device = device

    def __repr__(self):
        return "Maks: total patches {}, mask patches {}".format(self.total_patches, self.total_masks)

    def generate_batch(self, rng_words):
        dev = torch.device(self.device)
        if isinstance(rng_words, np.ndarray):
            rng_words = torch.from_numpy(np.ascontiguousarray(rng_words, dtype=np.uint32).view(np.int32)).to(dev)
        return _lib.tube_mask_bb(None, rng_words.contiguous(), (self.frames, self.height, self.width),
                                 self.num_masks_per_frame, 0.0)

    def __call__(self):
        W = words_per_clip(self.height, self.width)
        mask, _, _, used = self.generate_batch(_global_raw_words(W)[None, :])
        used = int(used.item())
        if used < 0:
            raise RuntimeError("tube mask: rng_words exhausted")
        _advance_global(used)
        return mask[0].cpu().numpy().astype(np.float64)
