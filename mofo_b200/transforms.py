"""GPU side of the reference's clip preprocessing for pretraining: ``transforms.py`` /
``datasets.py:38-58`` run, per sample and on the CPU, multi-scale crop -> bilinear resize to 224 -> stack -> /255 ->
normalise, with the motion box transformed alongside and the tube mask generated from it.  Here the host only DECIDES
(which crop) and ships raw uint8 frames; one kernel (``mofo_clip_preprocess``) produces the normalised NCTHW clip - bit
identical to cv2.resize(INTER_LINEAR) + ToTorchFormatTensor + GroupNormalize - and the transformed boxes, and the engine
generates the masks from those boxes on the GPU (``engine_for_pretraining.masks_from_bbox``).  The host moves 1 byte per
source pixel instead of 4 bytes per output sample.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib


class GroupMultiScaleCrop_BB_no_global_union:
    """Crop SELECTION of transforms.py:92-189 (same constructor, same candidate sizes / 13 fixed offsets, same use of
    ``np.random.seed(10)`` and Python's ``random.choice``); the crop itself is applied on the GPU by ``ClipPreprocessor``."""

    def __init__(self, input_size, scales=None, max_distort=1, fix_crop=True, more_fix_crop=True):
        self.scales = scales if scales is not None else [1, 875, .75, .66]
        self.max_distort = max_distort
        self.fix_crop = fix_crop
        self.more_fix_crop = more_fix_crop
        self.input_size = input_size if not isinstance(input_size, int) else [input_size, input_size]

    def _sample_crop_size(self, im_size):
        np.random.seed(10)                                    # transforms.py:139
        image_w, image_h = im_size[0], im_size[1]
        base_size = min(image_w, image_h)
        crop_sizes = [int(base_size * x) for x in self.scales]
        crop_h = [self.input_size[1] if abs(x - self.input_size[1]) < 3 else x for x in crop_sizes]
        crop_w = [self.input_size[0] if abs(x - self.input_size[0]) < 3 else x for x in crop_sizes]
        pairs = [(w, h) for i, h in enumerate(crop_h) for j, w in enumerate(crop_w) if abs(i - j) <= self.max_distort]
        crop_pair = random.choice(pairs)
        if not self.fix_crop:
            w_offset = random.randint(0, image_w - crop_pair[0])
            h_offset = random.randint(0, image_h - crop_pair[1])
        else:
            w_offset, h_offset = random.choice(self.fill_fix_offset(self.more_fix_crop, image_w, image_h, crop_pair[0], crop_pair[1]))
        return crop_pair[0], crop_pair[1], w_offset, h_offset

    @staticmethod
    def fill_fix_offset(more_fix_crop, image_w, image_h, crop_w, crop_h):
        w_step = (image_w - crop_w) // 4
        h_step = (image_h - crop_h) // 4
        ret = [(0, 0), (4 * w_step, 0), (0, 4 * h_step), (4 * w_step, 4 * h_step), (2 * w_step, 2 * h_step)]
        if more_fix_crop:
            ret += [(0, 2 * h_step), (4 * w_step, 2 * h_step), (2 * w_step, 4 * h_step), (2 * w_step, 0 * h_step),
                    (1 * w_step, 1 * h_step), (3 * w_step, 1 * h_step), (1 * w_step, 3 * h_step), (3 * w_step, 3 * h_step)]
        return ret


class ClipPreprocessor:
    """Batched device replacement of ``DataAugmentationForVideoMAE_BB.transform`` (datasets.py:38-50).

    ``pre(frames_u8, boxes, crops=None) -> (videos f32 [B,3,T,S,S], boxes f64 [B,T,4])`` with ``frames_u8`` uint8
    [B,T,H,W,3] (pinned host or CUDA), ``boxes`` [B,T,4] pascal_voc in frame pixels, ``crops`` int [B,4] =
    (x_off, y_off, crop_w, crop_h); when ``crops`` is None one crop per clip is drawn as the reference does."""

    def __init__(self, input_size=224, scales=(1, .875, .75, .66), device="cuda"):
        self.size = int(input_size)
        self.chooser = GroupMultiScaleCrop_BB_no_global_union(self.size, list(scales))
        self.device = torch.device(device)

    def sample_crops(self, B, im_w, im_h):
        out = np.empty((B, 4), dtype=np.int32)
        for b in range(B):
            cw, ch, xo, yo = self.chooser._sample_crop_size((im_w, im_h))
            out[b] = (xo, yo, cw, ch)
        return out

    def __call__(self, frames_u8, boxes, crops=None, out=None):
        B, T, H, W, _ = frames_u8.shape
        if crops is None:
            crops = self.sample_crops(B, W, H)
        crops_t = torch.as_tensor(np.asarray(crops, dtype=np.int32)).contiguous()
        c = crops_t.numpy()
        if (c[:, 0] < 0).any() or (c[:, 1] < 0).any() or (c[:, 0] + c[:, 2] > W).any() or (c[:, 1] + c[:, 3] > H).any() or (c[:, 2:] < 2).any():
            raise ValueError("crop outside the frame")
        dev = self.device
        frames_d = frames_u8.to(dev, non_blocking=True).contiguous()
        boxes_d = torch.as_tensor(np.asarray(boxes.cpu() if isinstance(boxes, torch.Tensor) else boxes, dtype=np.float64)).contiguous().to(dev, non_blocking=True)
        if out is None:
            out = torch.empty(B, 3, T, self.size, self.size, dtype=torch.float32, device=dev)
        boxes_out = torch.empty(B, T, 4, dtype=torch.float64, device=dev)
        _lib.clip_preprocess(frames_d, crops_t.to(dev, non_blocking=True), boxes_d, self.size, out, boxes_out)
        return out, boxes_out


class RawClipLoader:
    """Marks a loader of RAW batches ``(frames_u8 [B,T,H,W,3] (pinned host), boxes [B,T,4], crops int [B,4] | None)`` for
    ``train_one_epoch_BB``: the engine's prefetcher copies the uint8 frames to the GPU and runs ``mofo_clip_preprocess`` on
    its copy stream while the previous step computes, then generates the masks from the transformed boxes on the GPU
    (``mofo_gpu_masks``).  Nothing but uint8 frames, boxes and four crop integers per clip crosses PCIe."""

    mofo_gpu_masks = True
    mofo_raw_frames = True

    def __init__(self, loader, preprocessor: ClipPreprocessor, mask_ratios=(0.9, 0.75)):
        self.loader, self.preprocessor = loader, preprocessor
        self.mofo_mask_ratios = mask_ratios
        self.quiet = getattr(loader, "quiet", False)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        return iter(self.loader)
