"""GPU parity of the HBM-bound kernels (mask, gather, target+MSE, LayerNorm, assemble, helpers) through the C ABI,
against the CPU oracle (oracle/) and plain torch fp32 math on the same inputs."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mask_oracle as mo
from oracle import target_oracle as tgt


@pytest.fixture(scope="module")
def lib():
    from mofo_b200 import _lib
    _lib.load()
    return _lib


def dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def words_tensor(seeds, W):
    w = np.stack([mo.mt19937_words(int(s), W) for s in seeds]).astype(np.uint32)
    return torch.from_numpy(w.view(np.int32)).cuda()


def test_mask_golden_and_random(lib, golden_dir):
    with open(os.path.join(golden_dir, "mask_golden.json")) as f:
        gold = json.load(f)
    cases = gold["bb_cases"]
    bb = dev(np.asarray([c["box"] for c in cases], dtype=np.float64))
    words = words_tensor([c["seed"] for c in cases], 640)
    mask, vis, msk, used = lib.tube_mask_bb(bb, words, (8, 14, 14), 176, 0.75)
    torch.cuda.synchronize()
    mask = mask.cpu().numpy()
    for i, c in enumerate(cases):
        assert hashlib.sha1(mask[i].tobytes()).hexdigest() == c["sha1"], c
        v_ref, m_ref = mo.index_lists(mask[i])
        assert np.array_equal(vis[i].cpu().numpy(), v_ref) and np.array_equal(msk[i].cpu().numpy(), m_ref)
    # random float boxes, different seeds per clip, bit-exact vs oracle incl. words consumed
    rng = np.random.default_rng(3)
    B = 257
    boxes = np.zeros((B, 4))
    for b in range(B):
        x1, y1 = rng.uniform(0, 200, 2); w, h = rng.uniform(0.5, 224, 2)
        boxes[b] = (x1, y1, min(224, x1 + w), min(224, y1 + h))
    boxes[0] = (0, 0, 1, 1); boxes[1] = (0, 0, 224, 224); boxes[2] = (16, 16, 16, 16); boxes[3] = (300, 300, 400, 400)
    seeds = np.arange(B) + 1000
    mask, vis, msk, used = lib.tube_mask_bb(dev(boxes), words_tensor(seeds, 700), (8, 14, 14), 176, 0.75)
    torch.cuda.synchronize()
    for b in range(B):
        ref, n_used = mo.tube_mask_bb(np.tile(boxes[b], (16, 1)), mo.mt19937_words(int(seeds[b]), 700))
        assert np.array_equal(mask[b].cpu().numpy(), ref.astype(np.uint8)), b
        assert int(used[b]) == n_used
    assert int(mask.sum()) == B * 1408


def test_mask_plain_and_small_grid_and_exhaustion(lib, golden_dir):
    with open(os.path.join(golden_dir, "mask_golden.json")) as f:
        gold = json.load(f)
    mask, vis, msk, used = lib.tube_mask_bb(None, words_tensor([gold["plain"]["seed"]], 700), (8, 14, 14), 176, 0.0)
    assert hashlib.sha1(mask[0].cpu().numpy().tobytes()).hexdigest() == gold["plain"]["sha1"]
    for c in gold["tiny_grid_cases"]:
        mask, vis, msk, used = lib.tube_mask_bb(dev(np.asarray([c["box"]], dtype=np.float64)),
                                                words_tensor([c["seed"]], 200), (8, 4, 4), 14, 0.75)
        assert mask[0].cpu().numpy().tolist() == c["mask"]
    # too few rng words -> words_used == -1 (host raises on it), never a hang or OOB read
    mask, vis, msk, used = lib.tube_mask_bb(dev(np.asarray([[60., 40, 160, 180]])), words_tensor([10], 16), (8, 14, 14), 176, 0.75)
    assert int(used[0]) == -1


def test_gather_matches_conv_patch_order(lib):
    torch.manual_seed(0)
    B, size, frames = 2, 64, 8
    video = torch.randint(0, 200, (B, 3, frames, size, size)).float().cuda()
    hw = size // 16
    ntok = (frames // 2) * hw * hw
    idx = torch.stack([torch.randperm(ntok)[:10].sort().values for _ in range(B)]).int().cuda()
    A = lib.gather_tubes(video, idx)
    # reference: unfold exactly as Conv3d flattens its weight (c, p0, p1, p2)
    v = video.reshape(B, 3, frames // 2, 2, hw, 16, hw, 16).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, ntok, 1536)
    ref = torch.stack([v[b][idx[b].long()] for b in range(B)]).reshape(-1, 1536)
    assert torch.equal(A.float(), ref)            # small integers are exact in bf16
    # and A @ W^T equals the Conv3d patch embedding on those tokens (modeling_finetune.py:247)
    W = torch.randint(-2, 3, (32, 3, 2, 16, 16)).float().cuda()
    conv = torch.nn.functional.conv3d(video, W, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    ref2 = torch.stack([conv[b][idx[b].long()] for b in range(B)]).reshape(-1, 32)
    assert torch.equal(A.float() @ W.reshape(32, -1).t(), ref2)


@pytest.mark.parametrize("normalize", [True, False])
def test_target_mse(lib, normalize):
    B = 3
    vid = tgt.synthetic_clip(B, seed=5)
    boxes = tgt.synthetic_boxes(B, seed=6)
    masks = np.stack([mo.tube_mask_bb(boxes[b], mo.mt19937_words(20 + b, 600))[0] for b in range(B)])
    mask_t = torch.from_numpy(masks).bool()
    labels = tgt.build_labels(vid, mask_t, normalize)
    msk_idx = torch.from_numpy(np.stack([mo.index_lists(masks[b])[1] for b in range(B)])).cuda()
    g = torch.Generator().manual_seed(1)
    pred = (labels + 0.3 * torch.randn(labels.shape, generator=g)).bfloat16()
    n = B * 1408
    lp = torch.zeros(n, device="cuda"); loss = torch.zeros(1, device="cuda")
    dpred = torch.empty(n, 1536, dtype=torch.bfloat16, device="cuda")
    lab_out = torch.empty(n, 1536, device="cuda")
    lib.target_mse(vid.cuda(), msk_idx, pred.cuda().reshape(n, 1536), lp, loss, dpred, normalize, 1.0, lab_out)
    torch.cuda.synchronize()
    lab_out = lab_out.cpu().reshape(labels.shape)
    if normalize:
        assert (lab_out - labels).abs().max().item() < 3e-5
    else:
        assert torch.equal(lab_out, labels)        # pure indexing + (x*std+mean): bit-exact
    ref_loss = tgt.mse_loss(pred.float(), lab_out).item()
    assert abs(loss.item() - ref_loss) < 1e-5 * ref_loss
    ref_d = (2.0 / pred.numel()) * (pred.float() - lab_out)
    err = (dpred.cpu().float().reshape(labels.shape) - ref_d).norm() / ref_d.norm()
    assert err < 5e-3, err


@pytest.mark.parametrize("M,D", [(64, 384), (1000, 768), (77, 64), (300, 1024), (129, 192)])
def test_layernorm_fwd_bwd(lib, M, D):
    torch.manual_seed(M + D)
    x = (torch.randn(M, D) * 2 + 0.5).cuda()
    gamma = (1 + 0.1 * torch.randn(D)).cuda(); beta = (0.1 * torch.randn(D)).cuda()
    y = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    lib.layernorm_fwd(x, gamma, beta, y, mean, rstd, M, D)
    xr = x.clone().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps=1e-6)
    assert (y.float() - yr).abs().max().item() < 0.03
    assert ((y.float() - yr).norm() / yr.norm()).item() < 4e-3
    dy = torch.randn(M, D, device="cuda").bfloat16()
    dres = torch.randn(M, D, device="cuda")
    dx = torch.empty(M, D, device="cuda"); dxb = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
    lib.layernorm_bwd(dy, x, gamma, mean, rstd, dres, M, D, dx, dxb, dg, db)
    yr.backward(dy.float())
    ref_dx = xr.grad + dres
    assert ((dx - ref_dx).norm() / ref_dx.norm()).item() < 1e-5
    assert ((dxb.float() - ref_dx).norm() / ref_dx.norm()).item() < 5e-3
    assert ((dg - gr.grad).norm() / gr.grad.norm()).item() < 1e-5
    assert ((db - br.grad).norm() / br.grad.norm()).item() < 1e-5


def test_layernorm_row_mapping(lib):
    B, N, nm, D = 3, 20, 14, 128
    torch.manual_seed(1)
    x = torch.randn(B * N, D).cuda(); gamma = torch.ones(D).cuda() * 1.5; beta = torch.zeros(D).cuda() + 0.25
    M = B * nm
    y = torch.empty(M, D, dtype=torch.bfloat16, device="cuda"); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    lib.layernorm_fwd(x, gamma, beta, y, mean, rstd, M, D, 1e-6, nm, N, N - nm)
    ref = torch.nn.functional.layer_norm(x.reshape(B, N, D)[:, -nm:], (D,), gamma, beta, eps=1e-6).reshape(M, D)
    assert ((y.float() - ref).norm() / ref.norm()).item() < 4e-3
    dy = torch.randn(M, D, device="cuda").bfloat16()
    dx = torch.zeros(B * N, D, device="cuda")
    dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
    lib.layernorm_bwd(dy, x, gamma, mean, rstd, None, M, D, dx, None, dg, db, nm, N, N - nm)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr.reshape(B, N, D)[:, -nm:], (D,), gamma, beta, eps=1e-6).reshape(M, D).backward(dy.float())
    assert ((dx - xr.grad).norm() / xr.grad.norm()).item() < 1e-5
    assert dx.reshape(B, N, D)[:, : N - nm].abs().max().item() == 0.0


def test_assemble_fwd_bwd(lib):
    B, nv, nm, Dd = 3, 6, 26, 192
    torch.manual_seed(2)
    N = nv + nm
    pos = torch.randn(N, Dd).cuda(); mt = torch.randn(Dd).cuda()
    msk_idx = torch.stack([torch.randperm(N)[:nm].sort().values for _ in range(B)]).int().cuda()
    xf = torch.zeros(B * N, Dd, device="cuda")
    lib.decoder_assemble_fwd(mt, pos, msk_idx, B, nv, nm, Dd, xf)
    xf = xf.reshape(B, N, Dd)
    for b in range(B):
        assert torch.equal(xf[b, nv:], mt[None] + pos[msk_idx[b].long()])
        assert xf[b, :nv].abs().max().item() == 0
    dxf = torch.randn(B * N, Dd, device="cuda")
    dmt = torch.zeros(Dd, device="cuda"); dvis = torch.empty(B * nv, Dd, dtype=torch.bfloat16, device="cuda")
    lib.decoder_assemble_bwd(dxf, B, nv, nm, Dd, dmt, dvis)
    r = dxf.reshape(B, N, Dd)
    assert ((dmt - r[:, nv:].sum((0, 1))).abs().max() / r[:, nv:].sum((0, 1)).abs().max()).item() < 1e-5
    assert torch.equal(dvis.reshape(B, nv, Dd), r[:, :nv].bfloat16())


def test_helpers(lib):
    torch.manual_seed(3)
    W = torch.randn(96, 3, 2, 16, 16).cuda()
    Wb = torch.empty(96, 1536, dtype=torch.bfloat16, device="cuda"); Wt = torch.empty(1536, 96, dtype=torch.bfloat16, device="cuda")
    lib.cast_weight(W, Wb, Wt)
    assert torch.equal(Wb, W.reshape(96, -1).bfloat16()) and torch.equal(Wt, W.reshape(96, -1).t().bfloat16())
    W2 = torch.randn(70, 50).cuda(); Wt2 = torch.empty(50, 70, dtype=torch.bfloat16, device="cuda")
    lib.cast_weight(W2, None, Wt2)
    assert torch.equal(Wt2, W2.t().bfloat16())
    qb, vb = torch.randn(192).cuda(), torch.randn(192).cuda()
    out = torch.empty(576, device="cuda")
    lib.pack_qkv_bias(qb, vb, out)
    assert torch.equal(out, torch.cat([qb, torch.zeros_like(qb), vb]))
    X = torch.randn(3000, 1152, device="cuda").bfloat16()
    cs = torch.zeros(384, device="cuda")
    lib.colsum_bf16(X[:, 768:], 3000, 384, cs)
    ref = X[:, 768:].float().sum(0)
    assert ((cs - ref).norm() / ref.norm()).item() < 1e-5
    x = torch.randn(1_000_003, device="cuda")
    o = torch.zeros(1, device="cuda")
    lib.sq_norm_f32(x, o)
    assert abs(o.item() - (x.double() ** 2).sum().item()) < 1e-4 * o.item()


def test_normalize_u8_bit_exact_and_step_equivalence(lib):
    """uint8 input path: bit-identical to ToTorchFormatTensor(div=True) + GroupNormalize (datasets.py:44-50) on CPU."""
    g = torch.Generator().manual_seed(4)
    u8 = torch.randint(0, 256, (2, 3, 16, 64, 64), dtype=torch.uint8, generator=g)
    out = torch.empty(u8.shape, dtype=torch.float32, device="cuda")
    lib.normalize_u8(u8.cuda(), out)
    mean = torch.tensor((0.485, 0.456, 0.406))[None, :, None, None, None]
    std = torch.tensor((0.229, 0.224, 0.225))[None, :, None, None, None]
    ref = u8.float().div(255).sub_(mean).div_(std)
    assert torch.equal(out.cpu(), ref)
    # the fused step gives the same loss for the raw uint8 clip and for its normalised fp32 version
    from functools import partial
    from mofo_b200 import modeling_pretrain as mp
    from oracle import model_oracle as mdl
    cfg = mdl.tiny_config(img=64, frames=16)
    m = mp.PretrainVisionTransformer(img_size=64, patch_size=16, encoder_embed_dim=cfg.enc_dim, encoder_depth=cfg.enc_depth,
                                     encoder_num_heads=cfg.enc_heads, decoder_embed_dim=cfg.dec_dim, decoder_depth=cfg.dec_depth,
                                     decoder_num_heads=cfg.dec_heads, mlp_ratio=4, qkv_bias=True,
                                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    m.load_state_dict(mdl.random_state_dict(cfg, seed=2, perturb=0.05)); m.cuda()
    masks = np.stack([mo.tube_mask_bb([[8, 8, 40, 40]] * 16, mo.mt19937_words(b, 300), cfg.grid)[0] for b in range(2)])
    mask = torch.from_numpy(masks).bool().cuda()
    l_u8 = m.pretrain_step(u8.cuda(), mask).item()
    l_f32 = m.pretrain_step(ref.cuda(), mask).item()
    assert l_u8 == l_f32


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 3, 1568, 1568), (3, 1, 40, 2048), (1, 2, 7, 12)])
def test_masked_softmax_fwd_bwd(B, H, Nq, Nk):
    """Key-masked softmax pair of the box-focused classifier's cross attention vs torch fp32 (modeling_finetune.py:150-154)."""
    from mofo_b200 import _lib
    torch.manual_seed(B * 100 + Nk)
    dev = "cuda"
    S = torch.randn(B, H, Nq, Nk, device=dev) * 6
    allowed = (torch.rand(B, Nk, device=dev) < 0.6)
    allowed[:, 0] = True
    allowed[0] = True
    scale = 0.37
    P = torch.empty(B, H, Nq, Nk, dtype=torch.bfloat16, device=dev)
    _lib.masked_softmax_fwd(S, allowed.to(torch.uint8), scale, P)
    want = (S * scale).masked_fill(~allowed[:, None, None, :], float("-inf")).softmax(-1)
    assert torch.all(P.float()[~allowed[:, None, None, :].expand_as(P)] == 0)
    assert (P.float() - want).abs().max().item() <= 2 ** -8 * want.max().item() + 1e-6      # bf16 rounding of values <= 1
    assert (P.float().sum(-1) - 1).abs().max().item() < 5e-3
    dP = torch.randn(B, H, Nq, Nk, device=dev)
    dS = torch.empty_like(P)
    _lib.masked_softmax_bwd(P, dP, scale, dS)
    Pf = P.float()
    want_dS = scale * Pf * (dP - (Pf * dP).sum(-1, keepdim=True))
    err = (dS.float() - want_dS).abs().max().item()
    assert err <= 2 ** -8 * want_dS.abs().max().item() + 1e-6, err


def test_cast_f32_bf16_strided():
    from mofo_b200 import _lib
    torch.manual_seed(5)
    src = torch.randn(777, 1536, device="cuda")
    dst = torch.zeros(777, 2304, dtype=torch.bfloat16, device="cuda")
    _lib.cast_f32_bf16(src[:, 256:1280], dst[:, 768:1792])
    assert torch.equal(dst[:, 768:1792], src[:, 256:1280].bfloat16())
    assert dst[:, :768].abs().max().item() == 0 and dst[:, 1792:].abs().max().item() == 0
