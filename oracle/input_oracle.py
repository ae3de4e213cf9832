"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's clip preprocessing on the pretraining path:
``GroupMultiScaleCrop_BB_no_global_union`` (transforms.py:92-189: albumentations Crop + Resize(224, interpolation=1) with the
pascal_voc box riding along) -> ``Stack`` -> ``ToTorchFormatTensor(div=True)`` -> ``GroupNormalize`` (datasets.py:44-50).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the product path never does.

Pixels.  albumentations' Resize calls ``cv2.resize(img, (224, 224), interpolation=cv2.INTER_LINEAR)`` on the uint8 HWC
frame.  OpenCV (imgproc/resize.cpp, 8-bit linear path) is restated here and PINNED against cv2 itself, which is present
in this image (tests/test_oracle_input.py): per output column dx, fx = (float)((dx + 0.5) * scale_x - 0.5),
sx = floor(fx), fx -= sx, taps outside the row are reset (sx < 0 -> sx = 0, fx = 0; sx >= W - 1 -> sx = W - 1, fx = 0);
coefficients are 11-bit fixed point, a1 = round_half_even(fx * 2048), a0 = round_half_even((1 - fx) * 2048); rows use the
same taps but are CLIPPED instead of reset (fy is kept); horizontal pass S = p[sx] * a0 + p[sx + 1] * a1 (int32), vertical
pass out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.

Boxes.  albumentations is NOT in this image and the reference does not pin its version (README.md:99 ``pip install -U``), so
the box transform is restated from the library's published algorithm for ``A.Compose([A.Crop, A.Resize],
bbox_params=A.BboxParams(format='pascal_voc'))`` and is UNPINNED: normalise by the frame size, shift / rescale into the crop,
drop a box that falls completely outside it (the reference then substitutes [0, 0, 1, 1], transforms.py:120-123), clip to
the crop, scale to the output size.
"""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def linear_taps(dn, sn, is_y):
    """(index, a0, a1) of OpenCV's 8-bit INTER_LINEAR for a dimension resized sn -> dn."""
    scale = 1.0 / (dn / sn)                      # scale_x = 1. / inv_scale_x, inv_scale_x = (double)dsize / ssize
    idx = np.empty(dn, np.int32); a0 = np.empty(dn, np.int32); a1 = np.empty(dn, np.int32)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if not is_y:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= sn - 1:
                s, f = sn - 1, np.float32(0)
        idx[d] = s
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
        a0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
    return idx, a0, a1


def resize_linear_u8(src, dh, dw):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit for bit."""
    sh, sw = src.shape[:2]
    xi, xa0, xa1 = linear_taps(dw, sw, False)
    yi, ya0, ya1 = linear_taps(dh, sh, True)
    s = src.astype(np.int64)
    xi1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi] * xa0[None, :, None] + s[:, xi1] * xa1[None, :, None]
    y0 = np.clip(yi, 0, sh - 1); y1 = np.clip(yi + 1, 0, sh - 1)
    out = (((ya0[:, None, None].astype(np.int64) * (rows[y0] >> 4)) >> 16) + ((ya1[:, None, None].astype(np.int64) * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def to_tensor_normalize(frames_u8):
    """[T,H,W,3] uint8 -> f32 [3,T,H,W]: x / 255 then (x - mean) / std with fp32 roundings as torch does
    (ToTorchFormatTensor(div=True) + GroupNormalize, transforms.py:346-382; the dataset then reshapes to C,T,H,W)."""
    x = frames_u8.astype(np.float32) / np.float32(255.0)
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(3, 0, 1, 2))


def transform_box(box, im_w, im_h, crop, out_size=224):
    """One pascal_voc box (x1, y1, x2, y2) of an im_w x im_h frame through Crop(crop) + Resize(out_size); returns the
    transformed box or the reference's fallback [0, 0, 1, 1] when albumentations would have dropped it."""
    x_off, y_off, cw, ch = crop
    x1, y1, x2, y2 = [np.float64(v) for v in box]
    nx1, ny1, nx2, ny2 = x1 / im_w, y1 / im_h, x2 / im_w, y2 / im_h            # normalize_bbox
    cx1 = (nx1 * im_w - x_off) / cw; cy1 = (ny1 * im_h - y_off) / ch            # bbox_crop (denormalise, shift, re-normalise)
    cx2 = (nx2 * im_w - x_off) / cw; cy2 = (ny2 * im_h - y_off) / ch
    kx1, ky1, kx2, ky2 = (min(max(v, 0.0), 1.0) for v in (cx1, cy1, cx2, cy2))  # filter_bboxes: clip to the crop ...
    if (kx2 - kx1) * (ky2 - ky1) <= 0.0:                                          # ... and drop what has no area left
        return np.array([0.0, 0.0, 1.0, 1.0])
    return np.array([kx1 * out_size, ky1 * out_size, kx2 * out_size, ky2 * out_size])   # Resize keeps normalised boxes; denormalise


def preprocess_clip(frames_u8, boxes, crop, out_size=224):
    """frames_u8 [T,H,W,3], boxes [T,4], crop (x_off, y_off, crop_w, crop_h) -> (clip f32 [3,T,S,S], boxes f64 [T,4])."""
    x_off, y_off, cw, ch = crop
    T, H, W, _ = frames_u8.shape
    res = np.stack([resize_linear_u8(f[y_off:y_off + ch, x_off:x_off + cw], out_size, out_size) for f in frames_u8])
    out_boxes = np.stack([transform_box(b, W, H, crop, out_size) for b in boxes])
    return to_tensor_normalize(res), out_boxes


def fix_offsets(image_w, image_h, crop_w, crop_h, more_fix_crop=True):
    """GroupMultiScaleCrop.fill_fix_offset (transforms.py:161-185): the 13 candidate crop offsets."""
    w_step = (image_w - crop_w) // 4
    h_step = (image_h - crop_h) // 4
    ret = [(0, 0), (4 * w_step, 0), (0, 4 * h_step), (4 * w_step, 4 * h_step), (2 * w_step, 2 * h_step)]
    if more_fix_crop:
        ret += [(0, 2 * h_step), (4 * w_step, 2 * h_step), (2 * w_step, 4 * h_step), (2 * w_step, 0),
                (1 * w_step, 1 * h_step), (3 * w_step, 1 * h_step), (1 * w_step, 3 * h_step), (3 * w_step, 3 * h_step)]
    return ret


def crop_pairs(image_w, image_h, input_size=224, scales=(1, .875, .75, .66), max_distort=1):
    """transforms.py:139-155: candidate (crop_w, crop_h) pairs."""
    base = min(image_w, image_h)
    sizes = [int(base * x) for x in scales]
    ch = [input_size if abs(x - input_size) < 3 else x for x in sizes]
    cw = [input_size if abs(x - input_size) < 3 else x for x in sizes]
    return [(w, h) for i, h in enumerate(ch) for j, w in enumerate(cw) if abs(i - j) <= max_distort]
