"""GPU clip preprocessing (mofo_clip_preprocess) against the oracle (oracle/input_oracle.py, pinned to
cv2.resize in tests/test_oracle_input.py) and, where cv2 is importable, against cv2 + torch directly: pixels bit for bit,
boxes bit for bit against the oracle's restatement; then the whole chain raw frames -> clip -> GPU masks -> fused step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import input_oracle as io


@pytest.mark.parametrize("H,W,T,B", [(240, 320, 16, 3), (256, 340, 4, 2), (224, 224, 2, 1), (180, 200, 2, 2)])
def test_clip_preprocess_bit_exact(H, W, T, B):
    from mofo_b200 import transforms as tr
    rng = np.random.default_rng(H + W + T)
    frames = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    pairs = io.crop_pairs(W, H)
    crops, boxes = [], []
    for b in range(B):
        cw, ch = pairs[rng.integers(len(pairs))]
        xo, yo = io.fix_offsets(W, H, cw, ch)[rng.integers(13)]
        crops.append((xo, yo, cw, ch))
        x1 = rng.random(T) * (W - 40); y1 = rng.random(T) * (H - 40)
        bb = np.stack([x1, y1, x1 + 8 + rng.random(T) * 100, y1 + 8 + rng.random(T) * 100], 1)
        if b == 0:
            bb[0] = [0, 0, 1.5, 1.5]                       # (usually) cropped away -> fallback box
        boxes.append(bb)
    crops = np.array(crops, dtype=np.int32); boxes = np.stack(boxes)
    pre = tr.ClipPreprocessor(224)
    vid, bout = pre(torch.from_numpy(frames), boxes, crops)
    torch.cuda.synchronize()
    assert tuple(vid.shape) == (B, 3, T, 224, 224) and vid.dtype == torch.float32
    for b in range(B):
        want, wbox = io.preprocess_clip(frames[b], boxes[b], tuple(int(v) for v in crops[b]))
        assert np.array_equal(vid[b].cpu().numpy(), want), f"clip {b} differs from the oracle"
        assert np.array_equal(bout[b].cpu().numpy(), wbox), (bout[b].cpu().numpy(), wbox)
    try:
        import cv2
    except ImportError:
        return
    b = B - 1
    xo, yo, cw, ch = (int(v) for v in crops[b])
    res = np.stack([cv2.resize(f[yo:yo + ch, xo:xo + cw], (224, 224), interpolation=cv2.INTER_LINEAR) for f in frames[b]])
    t = torch.from_numpy(res).permute(0, 3, 1, 2).float().div(255.0)
    mean = torch.tensor(io.MEAN)[None, :, None, None]; std = torch.tensor(io.STD)[None, :, None, None]
    want = ((t - mean) / std).permute(1, 0, 2, 3).contiguous()
    assert torch.equal(vid[b].cpu(), want)


def test_raw_frames_to_fused_step():
    """uint8 frames + raw boxes -> GPU crop/resize/normalise + box transform -> GPU masks from the transformed boxes ->
    fused step; equals the step fed the oracle-preprocessed clip and the oracle's mask for the same transformed box."""
    from mofo_b200 import engine_for_pretraining as eng
    from mofo_b200 import modeling_pretrain as mp
    from mofo_b200 import transforms as tr
    from oracle import mask_oracle as mo
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    B, T, H, W = 2, 16, 240, 320
    frames = rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)
    boxes = np.repeat(np.array([[60., 40., 160., 180.], [100.5, 20.25, 230.75, 160.5]])[:, None, :], T, 1)
    crops = np.array([[27, 15, 210, 180], [0, 0, 240, 240]], dtype=np.int32)
    vid, bout = tr.ClipPreprocessor(224)(torch.from_numpy(frames), boxes, crops)
    mask, vis, msk = eng.masks_from_bbox(bout, dev)
    torch.manual_seed(0)
    model = mp.create_model("pretrain_mae_small_patch16_224", pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4).to(dev)
    model.use_cuda_graph = False
    loss = model.pretrain_step(vid, vis_idx=vis, msk_idx=msk).item()
    want_vid = np.stack([io.preprocess_clip(frames[b], boxes[b], tuple(int(v) for v in crops[b]))[0] for b in range(B)])
    want_box = np.stack([io.preprocess_clip(frames[b], boxes[b], tuple(int(v) for v in crops[b]))[1] for b in range(B)])
    want_mask = np.stack([mo.tube_mask_bb(want_box[b][0], mo.mt19937_words(10, 800), (8, 14, 14))[0] for b in range(B)])
    assert np.array_equal(mask.cpu().numpy(), want_mask.astype(np.uint8))
    loss2 = model.pretrain_step(torch.from_numpy(want_vid).to(dev), torch.from_numpy(want_mask).bool().to(dev)).item()
    assert abs(loss - loss2) <= 1e-6 * abs(loss2)


def test_raw_clip_loader_drives_the_engine():
    """RawClipLoader: raw uint8 frames + boxes -> (GPU preprocessing on a side stream) -> train_one_epoch_BB with GPU masks;
    same losses as the engine fed the oracle-preprocessed clips and the oracle's masks through a plain loader."""
    from mofo_b200 import engine_for_pretraining as eng
    from mofo_b200 import modeling_pretrain as mp
    from mofo_b200 import transforms as tr
    from mofo_b200 import utils as U
    from mofo_b200.optim_factory import FusedAdamW
    from oracle import mask_oracle as mo
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(8)
    B, T, H, W = 2, 16, 240, 320
    raw = []
    for _ in range(3):
        frames = torch.from_numpy(rng.integers(0, 256, (B, T, H, W, 3), dtype=np.uint8)).pin_memory()
        x1 = rng.integers(0, 150, B).astype(np.float64); y1 = rng.integers(0, 100, B).astype(np.float64)
        boxes = np.repeat(np.stack([x1, y1, x1 + 90, y1 + 110], 1)[:, None, :], T, 1)
        crops = np.array([[27, 15, 210, 180], [40, 0, 240, 240]], dtype=np.int32)
        raw.append((frames, boxes, crops))

    def run(loader):
        torch.manual_seed(0)
        model = mp.create_model("pretrain_mae_small_patch16_224", pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4).to(dev)
        opt = FusedAdamW([{"params": list(model.parameters()), "weight_decay": 0.05, "lr_scale": 1.0}], lr=1e-4, betas=(0.9, 0.95))
        log = []

        class Log:
            def update(self, head="x", **kw):
                if "loss" in kw:
                    log.append(kw["loss"])

            def set_step(self):
                pass
        eng.train_one_epoch_BB(model, loader, opt, dev, 0, U.NativeScalerWithGradNormCount(), max_norm=None, patch_size=16,
                               normlize_target=True, log_writer=Log(), start_steps=0)
        return log

    class Quiet(list):
        quiet = True
    got = run(tr.RawClipLoader(Quiet(raw), tr.ClipPreprocessor(224)))
    plain = Quiet()
    for frames, boxes, crops in raw:
        vids, masks = [], []
        for b in range(B):
            v, bo = io.preprocess_clip(frames[b].numpy(), boxes[b], tuple(int(c) for c in crops[b]))
            vids.append(v); masks.append(mo.tube_mask_bb(bo[0], mo.mt19937_words(10, 800), (8, 14, 14))[0])
        plain.append((torch.from_numpy(np.stack(vids)), torch.zeros(B, T, 4, dtype=torch.long), torch.from_numpy(np.stack(masks)).double()))
    want = run(plain)
    assert len(got) == len(want) == 3
    assert all(abs(a - b) <= 2e-5 * abs(b) for a, b in zip(got, want)), (got, want)
