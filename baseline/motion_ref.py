"""Drives the reference's own motion-box code on the CPU: bench.py's cpu_baseline / --impl reference leg for
``--workload motion`` and the golden-file generator (tests/golden/make_motion_golden.py).  Nothing here is on the product path.

``scripts/motion_sts.py`` is staged UNMODIFIED as baseline/_ref/motion_sts.py (baseline/setup_ref.py) and imported from there
(or straight from $MOFO_REFERENCE/scripts when that tree is present).  The frame loops that call it live inside nested
functions of the reference that open videos with decord (scripts/data/motion_map_creator.py:121-247,
scripts/data/SSV2/bounding_box_creator_SSV.py:57-475) and cannot be imported, so their statement sequences are replayed here
on in-memory arrays with the libraries they call (scipy.ndimage, numpy, cv2).
"""
from __future__ import annotations

import importlib.util
import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(os.environ.get("MOFO_REFERENCE", "/root/reference"), "scripts", "motion_sts.py"),
               os.path.join(HERE, "_ref", "motion_sts.py"))
_ms = None


def available() -> bool:
    try:
        import cv2  # noqa: F401
        import scipy.ndimage  # noqa: F401
    except Exception:
        return False
    return any(os.path.exists(c) for c in _CANDIDATES)


def load():
    """The reference's motion_sts module."""
    global _ms
    if _ms is None:
        path = next(c for c in _CANDIDATES if os.path.exists(c))
        spec = importlib.util.spec_from_file_location("mofo_reference_motion_sts", path)
        _ms = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ms)
    return _ms


def reference_motion_map(flows, ws=8):
    """motion_map_creator.py:160-228 (the frame loop of make_video_flow_mag): uint8 [T,H,W,3] flow frames -> uint8 [T,H,W,3]."""
    import cv2
    ms = load()
    duration = len(flows)
    frame_mags = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")       # the reference's astype(uint8) of out-of-range floats warns on newer numpy
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
    """bounding_box_creator_SSV.py:125-166 for one uint8 [H,W,3] motion-map frame -> (filtered, gray).  ``frame`` is modified
    in place where the reference modifies it."""
    import cv2
    from scipy.ndimage import gaussian_filter
    frame = gaussian_filter(frame, sigma=1)
    max_pixel_after_gaussian = np.max(frame)
    frame[frame < 0.4 * max_pixel_after_gaussian] = 0
    sigma = np.std(frame) + 1e-5
    frame[frame < 1.5 * sigma] = 0
    frame = gaussian_filter(frame, sigma=30)
    gray = cv2.cvtColor(frame.astype(np.uint8), cv2.COLOR_BGR2GRAY)
    return frame, gray


def synthetic_flow_video(seed, T, H, W):
    """A flow-like video: mid-grey background with noise, one blob that moves with a different flow value."""
    rng = np.random.default_rng(seed)
    v = np.clip(128 + 3 * rng.standard_normal((T, H, W, 3)), 0, 255)
    yy, xx = np.mgrid[:H, :W]
    for t in range(T):
        cy, cx = H * 0.45 + 0.6 * t * H / 40, W * 0.3 + 1.1 * t * W / 56
        blob = ((yy - cy) ** 2 / (H * 0.14) ** 2 + (xx - cx) ** 2 / (W * 0.12) ** 2) < 1
        v[t, blob, 0] += 60; v[t, blob, 1] -= 45
    return np.clip(v, 0, 255).astype(np.uint8)
