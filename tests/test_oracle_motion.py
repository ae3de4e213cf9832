"""The motion-box oracle (oracle/motion_oracle.py) pinned on the CPU: against the golden file the REFERENCE's own
code produced (tests/golden/make_motion_golden.py), against scipy.ndimage (float64 data exposes the summation order) and cv2."""
import os

import numpy as np
import pytest

from oracle import motion_oracle as mo

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "motion_golden.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_motion_map_matches_reference_golden(name):
    got = mo.motion_map(GOLD[f"flows_{name}"], int(GOLD[f"ws_{name}"]))
    assert np.array_equal(got, GOLD[f"map_{name}"])


def test_filter_matches_reference_golden():
    for t, frame in enumerate(GOLD["box_frames"]):
        filt, gray = mo.filter_frame(frame)
        assert np.array_equal(filt, GOLD["box_filtered"][t]), t
        assert np.array_equal(gray, GOLD["box_gray"][t]), t
    assert GOLD["box_gray"][5].max() == 6          # a flat frame of 7s: sum(w) * 7 < 7 in float64, the uint8 store truncates


def test_window_schedule():
    assert [mo.flow_window(i, 12, 8) for i in (1, 3, 4, 5, 8, 9, 12)] == [(0, 8), (0, 8), (0, 8), (1, 9), (4, 12), (4, 12), (4, 12)]
    assert [mo.flow_window(i, 5, 8) for i in (1, 4, 5)] == [(0, 5), (0, 5), (0, 5)]
    assert [mo.flow_window(i, 3, 1) for i in (1, 2, 3)] == [(0, 1), (1, 2), (2, 3)]
    los, his = zip(*[mo.flow_window(i, 40, 8) for i in range(1, 41)])
    assert list(los) == sorted(los) and list(his) == sorted(his)            # the window never moves backwards


def test_correlate_order_and_cast_vs_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    a = rng.random((50, 40, 3))
    for sigma in (1, 30):
        w, r = mo.gaussian_weights(sigma)
        assert r == int(4 * sigma + 0.5)
        for axis in range(3):
            assert np.array_equal(ndi.correlate1d(a, w[::-1], axis, mode="reflect"), mo.correlate1d_symmetric(a, w, axis))
    img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    flat = np.full((60, 80, 3), 7, np.uint8); flat[20:30, 30:50] = 201
    for sigma in (1, 30):
        assert np.array_equal(ndi.gaussian_filter(img, sigma), mo.gaussian_filter_u8(img, sigma))
        assert np.array_equal(ndi.gaussian_filter(flat, sigma), mo.gaussian_filter_u8(flat, sigma))


def test_magnitude_gray_and_cast_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    x = rng.integers(-12000, 12000, (64, 64)).astype(np.float32); y = rng.integers(-12000, 12000, (64, 64)).astype(np.float32)
    mag, _ = cv2.cartToPolar(x, y, angleInDegrees=True)
    assert np.array_equal(mag, mo.magnitude_f32(x, y))
    img = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), mo.bgr2gray_u8(img))
    big = (rng.random(4096) * 9000).astype(np.float32)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(big.astype(np.uint8), mo.wrap_u8(big))


def test_product_weight_table_matches_scipy_formula():
    """Host logic of mofo_b200/motion_boxes.py (no GPU): the weight table handed to the kernels is scipy's kernel, by distance."""
    from mofo_b200 import motion_boxes as mb
    for sigma in (1, 30, 2.5):
        w, r = mo.gaussian_weights(sigma)
        got = mb.gaussian_weights(sigma).numpy()
        assert got.dtype == np.float64 and got.shape == (r + 1,)
        assert np.array_equal(got, w[r:]) and np.array_equal(got, w[:r + 1][::-1])
