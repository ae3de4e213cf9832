"""Pixel stages of the reference's offline motion-box pipeline on the GPU.

The reference produces the per-frame motion boxes that ``TubeMaskingGenerator_BB`` consumes in two offline CPU scripts:

* ``scripts/data/motion_map_creator.py:121-247`` (with ``scripts/motion_sts.py:5-37``): optical-flow video -> motion-map
  video.  ``motion_map`` below is the frame loop of ``make_video_flow_mag`` (:160-228), one kernel per video.
* ``scripts/data/SSV2/bounding_box_creator_SSV.py:57-475`` (and the Epic-Kitchens twin): motion-map video -> boxes.
  ``filter_motion_map`` below is the per-frame filtering of ``json_creator`` (:125-166) up to the gray image that
  ``cv2.findContours`` receives, eight launches for a whole video.

Both are bit-exact with the reference's numpy / scipy / cv2 arithmetic (tests/test_gpu_motion.py).  Video decoding / encoding,
TV-L1 optical flow and the sequential contour search, ranking and temporal smoothing (:168-475) stay host code, as in the
reference.  There is no CPU fallback: the calls fail if libmofo_sm100.so is missing.
"""
from __future__ import annotations

import math

import torch

from . import _lib


def gaussian_weights(sigma: float, truncate: float = 4.0) -> torch.Tensor:
    """float64 weights by distance 0..r, r = int(truncate*sigma + 0.5), normalised over the full 2r+1 taps in ascending-x
    order - the arithmetic of scipy.ndimage's ``_gaussian_kernel1d(sigma, 0, r)`` (numpy does the same exp / sum)."""
    import numpy as np
    sd = float(sigma)
    r = int(truncate * sd + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return torch.from_numpy(phi[r:].copy())


def motion_map(flows: torch.Tensor, ws: int = 8, border: int = 8, channels: int = 3) -> torch.Tensor:
    """``flows`` uint8 [T,H,W,C] (CUDA) -> uint8 [T,H,W,channels]: the frames the reference writes to the motion-map video
    (motion_map_creator.py:160-228; ws = 8, or 4 for kinetics :141-142)."""
    out = torch.empty(flows.shape[:3] + (channels,), dtype=torch.uint8, device=flows.device)
    return _lib.motion_map(flows.contiguous(), int(ws), int(border), out)


class MotionMapFilter:
    """``filter(frames)``: uint8 [T,H,W,3] motion-map frames (CUDA) -> (filtered uint8 [T,H,W,3], gray uint8 [T,H,W]) as
    bounding_box_creator_SSV.py:125-166 computes them frame by frame (before_sigma = 1, remove_thrd = 0.4, 1.5 * (std + 1e-5),
    after_sigma = 30).  The weight tables and the workspace are cached per device / shape."""

    def __init__(self, before_sigma: float = 1, remove_thrd: float = 0.4, std_k: float = 1.5, std_eps: float = 1e-5,
                 after_sigma: float = 30):
        self.before_sigma, self.after_sigma = before_sigma, after_sigma
        self.remove_thrd, self.std_k, self.std_eps = float(remove_thrd), float(std_k), float(std_eps)
        self._w = {}
        self._work = None

    def filter(self, frames: torch.Tensor):
        dev = frames.device
        if dev not in self._w:
            self._w[dev] = (gaussian_weights(self.before_sigma).to(dev), gaussian_weights(self.after_sigma).to(dev))
        wb, wa = self._w[dev]
        frames = frames.contiguous()
        T = frames.shape[0]
        need = 2 * frames.numel()
        if self._work is None or self._work[0].numel() < need or self._work[0].device != dev or self._work[1].numel() < 4 * T:
            self._work = (torch.empty(need, dtype=torch.uint8, device=dev), torch.empty(4 * T, dtype=torch.int64, device=dev))
        filtered = torch.empty_like(frames)
        gray = torch.empty(frames.shape[:3], dtype=torch.uint8, device=dev)
        _lib.motion_box_filter(frames, wb, wa, self.remove_thrd, self.std_k, self.std_eps, self._work[0], self._work[1], filtered, gray)
        return filtered, gray


def filter_motion_map(frames: torch.Tensor, **kw):
    return MotionMapFilter(**kw).filter(frames)
