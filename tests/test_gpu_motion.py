"""GPU parity of the motion-box pixel stages through the C ABI: bit-exact against the golden file produced by
the reference's own code and against the CPU oracle on seeded inputs, including shapes that are not multiples of the tile,
frames smaller than the filter radius, ws = 1 / odd / larger than the video, and real SSv2-sized frames."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "motion_golden.npz"))


def _diff(got, want):
    bad = np.argwhere(got != want)
    return f"{len(bad)} of {want.size} differ; first at {bad[:4].tolist()} got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}" if len(bad) else ""


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_motion_map_golden(name):
    from mofo_b200 import motion_boxes as mb
    flows = torch.from_numpy(GOLD[f"flows_{name}"]).cuda()
    got = mb.motion_map(flows, ws=int(GOLD[f"ws_{name}"])).cpu().numpy()
    want = GOLD[f"map_{name}"]
    for c in range(3):
        assert np.array_equal(got[..., c], want), _diff(got[..., c], want)


@pytest.mark.parametrize("T,H,W,C,ws,border,oc", [(17, 240, 320, 3, 8, 8, 3), (6, 61, 95, 2, 5, 8, 1), (4, 20, 18, 3, 16, 3, 1),
                                                   (40, 100, 176, 4, 4, 0, 2), (2, 9, 7, 3, 8, 8, 1)])
def test_motion_map_vs_oracle(T, H, W, C, ws, border, oc):
    from mofo_b200 import motion_boxes as mb
    from oracle import motion_oracle as mo
    rng = np.random.default_rng(T * 1000 + H)
    flows = rng.integers(0, 256, (T, H, W, C), dtype=np.uint8)
    got = mb.motion_map(torch.from_numpy(flows).cuda(), ws=ws, border=border, channels=oc).cpu().numpy()
    want = mo.motion_map(flows, ws=ws, border=border) if border else None
    if border == 0:                                        # 0 = no border here (numpy's [-0:] would blank the frame)
        big = np.pad(flows, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="edge")
        want = mo.motion_map(big, ws=ws, border=1)[:, 1:-1, 1:-1]
    assert got.shape == (T, H, W, oc)
    for c in range(oc):
        assert np.array_equal(got[..., c], want), _diff(got[..., c], want)


def test_box_filter_golden():
    from mofo_b200 import motion_boxes as mb
    frames = torch.from_numpy(GOLD["box_frames"]).cuda()
    keep = frames.clone()
    filt, gray = mb.filter_motion_map(frames)
    assert torch.equal(frames, keep)
    assert np.array_equal(filt.cpu().numpy(), GOLD["box_filtered"]), _diff(filt.cpu().numpy(), GOLD["box_filtered"])
    assert np.array_equal(gray.cpu().numpy(), GOLD["box_gray"]), _diff(gray.cpu().numpy(), GOLD["box_gray"])


@pytest.mark.parametrize("T,H,W,generic", [(3, 240, 320, False), (2, 37, 53, False), (1, 130, 70, False), (1, 700, 24, False),
                                             (2, 61, 95, True)])
def test_box_filter_vs_oracle(T, H, W, generic, monkeypatch):
    """generic=True forces the table-driven fallback kernels (what frames too large for the staged kernels' shared-memory
    buffers get; H = 700 takes the fallback for the H pass on its own)."""
    from mofo_b200 import motion_boxes as mb
    from oracle import motion_oracle as mo
    if generic:
        monkeypatch.setenv("MOFO_MOTION_GENERIC", "1")
    else:
        monkeypatch.delenv("MOFO_MOTION_GENERIC", raising=False)
    rng = np.random.default_rng(H)
    yy, xx = np.mgrid[:H, :W]
    frames = np.zeros((T, H, W, 3), np.uint8)
    for t in range(T):
        blob = ((yy - H * (0.3 + 0.1 * t)) ** 2 / (H * 0.12) ** 2 + (xx - W * 0.55) ** 2 / (W * 0.1) ** 2) < 1
        f = np.clip(rng.integers(0, 12, (H, W)) + 180 * blob + rng.integers(-20, 20, (H, W)) * blob, 0, 255)
        frames[t] = np.clip(f[:, :, None] + rng.integers(-2, 3, (H, W, 3)), 0, 255)
    flt = mb.MotionMapFilter()
    filt, gray = flt.filter(torch.from_numpy(frames).cuda())
    filt2, gray2 = flt.filter(torch.from_numpy(frames).cuda())          # cached workspace, statistics cleared by the call
    assert torch.equal(filt, filt2) and torch.equal(gray, gray2)
    for t in range(T):
        wf, wg = mo.filter_frame(frames[t])
        assert np.array_equal(filt[t].cpu().numpy(), wf), (t, _diff(filt[t].cpu().numpy(), wf))
        assert np.array_equal(gray[t].cpu().numpy(), wg), (t, _diff(gray[t].cpu().numpy(), wg))
        assert wg.max() > 0


def test_motion_pipeline_properties_full_size():
    """Size-independent properties at a full SSv2 video (48 frames of 240 x 320): a static flow field has no motion boundaries
    away from the blob edge; shifting every flow byte by a constant changes nothing (the stencil sums to zero)."""
    from mofo_b200 import motion_boxes as mb
    rng = np.random.default_rng(5)
    flows = rng.integers(0, 200, (48, 240, 320, 3), dtype=np.uint8)
    a = mb.motion_map(torch.from_numpy(flows).cuda())
    b = mb.motion_map(torch.from_numpy(flows + 55).cuda())
    assert torch.equal(a, b)
    const = torch.full((48, 240, 320, 3), 93, dtype=torch.uint8, device="cuda")
    assert int(mb.motion_map(const).max()) == 0
    assert int(a[:, :8].max()) == 0 and int(a[:, :, -8:].max()) == 0 and int(a[:, 8:-8, 8:-8].max()) > 0


def test_other_sigmas_and_thresholds_vs_oracle():
    """Nothing in the kernels is tied to the reference's constants (sigma 1 / 30, 0.4, 1.5)."""
    from mofo_b200 import motion_boxes as mb
    from oracle import motion_oracle as mo
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (2, 45, 77, 3), dtype=np.uint8)
    frames[:, 10:30, 20:50] = np.clip(frames[:, 10:30, 20:50].astype(np.int64) + 90, 0, 255)
    kw = dict(before_sigma=2.5, remove_thrd=0.3, std_k=0.7, after_sigma=6)
    filt, gray = mb.filter_motion_map(torch.from_numpy(frames).cuda(), **kw)
    for t in range(2):
        wf, wg = mo.filter_frame(frames[t], **kw)
        assert np.array_equal(filt[t].cpu().numpy(), wf), _diff(filt[t].cpu().numpy(), wf)
        assert np.array_equal(gray[t].cpu().numpy(), wg)


def test_motion_calls_reject_bad_arguments():
    from mofo_b200 import _lib
    flows = torch.zeros(4, 16, 16, 3, dtype=torch.uint8, device="cuda")
    out = torch.empty(4, 16, 16, 1, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.MofoError):
        _lib.motion_map(flows, 0, 8, out)                                   # ws < 1
    with pytest.raises(_lib.MofoError):
        _lib.motion_map(flows[..., :1].contiguous(), 8, 8, out)             # needs the u and v channels
    with pytest.raises(AssertionError):
        _lib.motion_map(flows.cpu(), 8, 8, out)                             # host tensors never reach the library
    _lib.motion_map(flows, 8, 8, out)                                       # and the library is still usable afterwards
    assert int(out.max()) == 0


@pytest.mark.parametrize("chunk", ["1", "3", "1000"])
def test_motion_map_frame_chunking_is_invisible(chunk, monkeypatch):
    """The kernel splits a video into chunks of frames (each rebuilds its first window): any chunk length gives the same map."""
    from mofo_b200 import motion_boxes as mb
    from oracle import motion_oracle as mo
    rng = np.random.default_rng(9)
    flows = rng.integers(0, 256, (19, 40, 36, 3), dtype=np.uint8)
    monkeypatch.setenv("MOFO_MOTION_CHUNK", chunk)
    got = mb.motion_map(torch.from_numpy(flows).cuda(), ws=8, channels=1).cpu().numpy()[..., 0]
    want = mo.motion_map(flows, ws=8)
    assert np.array_equal(got, want), _diff(got, want)
