#!/usr/bin/env python
"""Generates tests/golden/motion_golden.npz by running the REFERENCE's own motion-box code in this container:
scripts/motion_sts.py's compute_motion_boudary / zero_boundary are imported from /root/reference and driven exactly as
scripts/data/motion_map_creator.py:160-228 drives them; the per-frame filtering is the statement sequence of
scripts/data/SSV2/bounding_box_creator_SSV.py:125-166 with the libraries it calls (scipy.ndimage.gaussian_filter, numpy, cv2)
- both replays live in baseline/motion_ref.py (the reference's own nested functions open videos with decord and cannot be
imported).   usage: python tests/golden/make_motion_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baseline.motion_ref import reference_filter_frame, reference_motion_map, synthetic_flow_video  # noqa: E402


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
