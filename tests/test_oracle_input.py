"""The input-pipeline oracle (oracle/input_oracle.py) pinned against the libraries the reference calls: cv2.resize
(INTER_LINEAR, uint8) bit for bit, and torch's ToTensor / normalise arithmetic.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as io

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("h,w", [(240, 320), (256, 340), (168, 224), (224, 224), (180, 180), (158, 211), (320, 240), (210, 240), (159, 158)])
def test_resize_bit_exact_vs_cv2(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = cv2.resize(img, (224, 224), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(io.resize_linear_u8(img, 224, 224), want)


def test_normalize_matches_torch_ops():
    rng = np.random.default_rng(1)
    fr = rng.integers(0, 256, (4, 32, 32, 3), dtype=np.uint8)
    t = torch.from_numpy(fr).permute(0, 3, 1, 2).float().div(255.0)                 # ToTorchFormatTensor(div=True)
    mean = torch.tensor(io.MEAN)[None, :, None, None]; std = torch.tensor(io.STD)[None, :, None, None]
    want = ((t - mean) / std).permute(1, 0, 2, 3).contiguous()                      # GroupNormalize, then C,T,H,W
    assert np.array_equal(io.to_tensor_normalize(fr), want.numpy())


def test_crop_candidates_and_box_transform():
    pairs = io.crop_pairs(320, 240)
    assert (240, 240) in pairs and (210, 240) in pairs and (158, 158) in pairs and (240, 180) not in pairs
    offs = io.fix_offsets(320, 240, 210, 180)
    assert len(offs) == 13 and offs[0] == (0, 0) and offs[3] == (4 * 27, 4 * 15)
    b = io.transform_box([60, 40, 160, 180], 320, 240, (20, 10, 210, 180))
    assert np.allclose(b, [(60 - 20) / 210 * 224, (40 - 10) / 180 * 224, (160 - 20) / 210 * 224, (180 - 10) / 180 * 224])
    assert np.array_equal(io.transform_box([0, 0, 10, 10], 320, 240, (100, 100, 158, 158)), [0, 0, 1, 1])   # dropped -> fallback
    clipped = io.transform_box([0, 0, 150, 150], 320, 240, (100, 100, 100, 100))
    assert np.allclose(clipped, [0, 0, 112, 112])
