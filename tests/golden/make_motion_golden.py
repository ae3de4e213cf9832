#!/usr/bin/env python
"""Generates tests/golden/motion_golden.npz by running the REFERENCE's own motion-box code in this container:
scripts/motion_sts.py's compute_motion_boudary / zero_boundary are imported from /root/reference and driven exactly as
scripts/data/motion_map_creator.py:160-228 drives them; the per-frame filtering is the statement sequence of
scripts/data/SSV2/bounding_box_creator_SSV.py:125-166 with the libraries it calls (scipy.ndimage.gaussian_filter, numpy, cv2).
The nested functions that hold those lines in the reference cannot be imported (they open videos with decord), so the lines
are replayed here on synthetic arrays.   usage: python tests/golden/make_motion_golden.py
"""
import os
import sys
import warnings

import cv2
import numpy as np
from scipy.ndimage import gaussian_filter

REF = os.environ.get("MOFO_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "scripts"))
import motion_sts as ms  # noqa: E402  (the reference's file)

warnings.simplefilter("ignore")       # the reference's astype(uint8) of out-of-range floats warns on newer numpy


def reference_motion_map(flows, ws):
    """motion_map_creator.py:160-228 (the frame loop of make_video_flow_mag) -> uint8 [T,H,W,3]."""
    duration = len(flows)
    frame_mags = []
    for idx in range(1, duration + 1):
        if ws == 1:
            flow_clip = [flows[idx - 1]]
        else:
            if idx - ws // 2 >= 0 and idx + ws // 2 <= duration:
                flow_clip = flows[idx - ws // 2: idx + ws // 2]
            elif idx - ws // 2 >= 0 and idx + ws // 2 > duration:
                flow_clip = flows[-ws:]
            elif idx + ws // 2 <= duration and idx - ws // 2 < 0:
                flow_clip = flows[:ws]
            else:
                flow_clip = flows[:]
        flows_u = list([flow[:, :, 0].astype(np.float32) for flow in flow_clip])
        flows_v = list([flow[:, :, 1].astype(np.float32) for flow in flow_clip])
        _, _, mb_x_u, mb_y_u = ms.compute_motion_boudary(flows_u)
        _, _, mb_x_v, mb_y_v = ms.compute_motion_boudary(flows_v)
        frame_mag_u, _ = cv2.cartToPolar(mb_x_u, mb_y_u, angleInDegrees=True)
        frame_mag_v, _ = cv2.cartToPolar(mb_x_v, mb_y_v, angleInDegrees=True)
        frame_mag = (frame_mag_u + frame_mag_v) / 2
        frame_mag = ms.zero_boundary(frame_mag)
        frame_mag = np.repeat(frame_mag[:, :, np.newaxis], 3, axis=2)
        frame_mags.append(frame_mag)
    return np.stack([frame.astype(np.uint8) for frame in frame_mags])


def reference_filter_frame(frame):
    """bounding_box_creator_SSV.py:125-166 for one frame -> (filtered, gray)."""
    frame = gaussian_filter(frame, sigma=1)
    max_pixel_after_gaussian = np.max(frame)
    frame[frame < 0.4 * max_pixel_after_gaussian] = 0
    sigma = np.std(frame) + 1e-5
    frame[frame < 1.5 * sigma] = 0
    frame = gaussian_filter(frame, sigma=30)
    gray = cv2.cvtColor(frame.astype(np.uint8), cv2.COLOR_BGR2GRAY)
    return frame, gray


def synthetic_flow_video(seed, T, H, W):
    """A flow-like video: mid-grey background with noise, one blob moving with a different flow value."""
    rng = np.random.default_rng(seed)
    v = np.clip(128 + 3 * rng.standard_normal((T, H, W, 3)), 0, 255)
    yy, xx = np.mgrid[:H, :W]
    for t in range(T):
        cy, cx = H * 0.45 + 0.6 * t, W * 0.3 + 1.1 * t
        blob = ((yy - cy) ** 2 / (H * 0.14) ** 2 + (xx - cx) ** 2 / (W * 0.12) ** 2) < 1
        v[t, blob, 0] += 60; v[t, blob, 1] -= 45
    return np.clip(v, 0, 255).astype(np.uint8)


def main():
    out = {}
    cases = [("a", 1, 12, 40, 56, 8), ("b", 2, 5, 33, 47, 8), ("c", 3, 9, 40, 40, 4), ("d", 4, 3, 30, 34, 1)]
    for name, seed, T, H, W, ws in cases:
        rng = np.random.default_rng(seed)
        flows = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8) if name != "a" else synthetic_flow_video(seed, T, H, W)
        mm = reference_motion_map(flows, ws)
        assert np.array_equal(mm[..., 0], mm[..., 1]) and np.array_equal(mm[..., 0], mm[..., 2])
        out[f"flows_{name}"] = flows; out[f"ws_{name}"] = np.int64(ws); out[f"map_{name}"] = mm[..., 0].copy()
    # stage B: frames of the synthetic video's map, plus noisy 3-channel frames (what an mp4 decode hands back is not exactly grey)
    mm = reference_motion_map(synthetic_flow_video(7, 6, 72, 96), 8)
    rng = np.random.default_rng(11)
    noisy = np.clip(mm[2:4].astype(np.int64) + rng.integers(-3, 4, mm[2:4].shape), 0, 255).astype(np.uint8)
    frames = np.concatenate([mm[:2], noisy, np.zeros_like(mm[:1]), np.full_like(mm[:1], 7)])
    filt, gray = zip(*[reference_filter_frame(f.copy()) for f in frames])
    out["box_frames"] = frames; out["box_filtered"] = np.stack(filt); out["box_gray"] = np.stack(gray)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "motion_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})
    print("gray max per frame:", [int(g.max()) for g in gray])


if __name__ == "__main__":
    main()
