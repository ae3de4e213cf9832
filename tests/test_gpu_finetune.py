"""Finetuning classifier (BASELINE configs[4]): the B200 VisionTransformer against the REFERENCE's own
modeling_finetune.VisionTransformer (baseline/_ref, unmodified) on the same GPU - logits and every parameter gradient of a
cross-entropy step, reference fp32 (TF32 off) as ground truth and the reference under bf16 autocast as the calibration of
what a 16-bit path can reach (same rule as tests/test_gpu_reference_parity.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _refrun():
    from baseline import refrun
    if not refrun.available():
        pytest.skip("baseline/_ref is not staged (python baseline/setup_ref.py needs /root/reference)")
    return refrun


def _ref_step(model, x, y, amp):
    model.zero_grad(set_to_none=True)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            logits = model(x)
            loss = torch.nn.functional.cross_entropy(logits.float(), y)
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return logits.detach().float(), loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}


@pytest.mark.parametrize("name,B,classes,init_scale", [("vit_small_patch16_224", 2, 174, 1.0), ("vit_base_patch16_224", 2, 174, 0.001),
                                                       ("vit_base_patch16_224", 3, 97, 1.0)])
def test_classifier_step_matches_reference(name, B, classes, init_scale):
    refrun = _refrun()
    ref = refrun.load()
    from mofo_b200 import modeling_finetune as mf
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    kw = dict(num_classes=classes, all_frames=16, tubelet_size=2, drop_rate=0.0, drop_path_rate=0.0, attn_drop_rate=0.0,
              use_mean_pooling=True, init_scale=init_scale)
    ref_model = getattr(ref.modeling_finetune, name)(pretrained=False, **kw).to(dev).train()
    with torch.no_grad():                           # biases / norms away from their (0, 1) initialisation
        for n, p in ref_model.named_parameters():
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.05)
    ours = mf.create_model(name, pretrained=False, drop_block_rate=None, **kw)
    sd = ref_model.state_dict()
    assert list(ours.state_dict().keys()) == list(sd.keys())
    ours.load_state_dict(sd, strict=True)
    ours = ours.to(dev).train()
    x = refrun.synthetic_batches(B, 1, seed=11, device=dev)[0][0]
    y = torch.randint(0, classes, (B,), device=dev)

    l32, loss32, g32 = _ref_step(ref_model, x, y, amp=False)
    l16, loss16, g16 = _ref_step(ref_model, x, y, amp=True)
    logits = ours(x)
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (B, classes)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    g = {n: p.grad.detach().clone() for n, p in ours.named_parameters()}

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()
    e_logits, r_logits = rel(logits.detach(), l32), rel(l16, l32)
    e_loss, r_loss = abs(loss.item() - loss32) / abs(loss32), abs(loss16 - loss32) / abs(loss32)
    print(f"{name} B={B}: logits rel {e_logits:.3e} (reference bf16 {r_logits:.3e}); loss rel {e_loss:.3e} ({r_loss:.3e})")
    assert e_logits <= max(2e-2, 2.5 * r_logits) and e_loss <= max(1e-3, 2.5 * r_loss)
    bad, worst = [], (0.0, "")
    for n in g32:
        e, r = rel(g[n], g32[n]), rel(g16[n], g32[n])
        worst = max(worst, (e, n))
        if e > max(3e-2, 2.5 * r):
            bad.append((n, e, r))
    print("worst gradient:", worst)
    assert not bad, bad[:8]
    # eval / no_grad forward returns the same logits (every forward kernel is deterministic)
    with torch.no_grad():
        assert torch.equal(ours(x), logits.detach())


def test_classifier_drop_path_matches_reference_with_the_same_draws():
    """DropPath 0.1 (the finetuning recipe): with the SAME uniform draws behind the per-sample keep masks, logits and every
    gradient match the reference (fp32 truth, bf16-autocast calibration).  The reference's drop_path (timm semantics:
    floor(keep + U) / keep per sample) is fed the draws our model uses."""
    refrun = _refrun()
    ref = refrun.load()
    from mofo_b200 import modeling_finetune as mf
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    name, B, classes = "vit_small_patch16_224", 6, 174
    kw = dict(num_classes=classes, all_frames=16, tubelet_size=2, drop_rate=0.0, drop_path_rate=0.3, attn_drop_rate=0.0,
              use_mean_pooling=True, init_scale=1.0)
    ref_model = getattr(ref.modeling_finetune, name)(pretrained=False, **kw).to(dev).train()
    ours = mf.create_model(name, pretrained=False, drop_block_rate=None, **kw)
    ours.load_state_dict(ref_model.state_dict(), strict=True)
    ours = ours.to(dev).train()
    depth = len(ours.blocks)
    U = torch.rand(2 * depth, B, device=dev)
    ours._drop_path_uniform = lambda n, b, d: U.clone()
    dpr = ours.dpr
    calls = {}

    def fixed_drop_path(x, drop_prob=0., training=False):
        if drop_prob == 0. or not training:
            return x
        i = min(range(depth), key=lambda k: abs(dpr[k] - drop_prob))          # block index from its (unique) rate
        j = calls.get(i, 0); calls[i] = j + 1                                   # 0 = attention branch, 1 = MLP branch
        keep = 1.0 - drop_prob
        mask = (keep + U[2 * i + (j % 2)]).floor()
        return x.div(keep) * mask.to(x.dtype).view(-1, *([1] * (x.ndim - 1)))
    orig = ref.modeling_finetune.drop_path
    ref.modeling_finetune.drop_path = fixed_drop_path
    try:
        x = refrun.synthetic_batches(B, 1, seed=13, device=dev)[0][0]
        y = torch.randint(0, classes, (B,), device=dev)
        calls.clear(); l32, loss32, g32 = _ref_step(ref_model, x, y, amp=False)
        calls.clear(); l16, loss16, g16 = _ref_step(ref_model, x, y, amp=True)
    finally:
        ref.modeling_finetune.drop_path = orig
    assert (U[2:] + (1.0 - torch.tensor(dpr, device=dev).repeat_interleave(2)[2:, None]) < 1.0).any(), "no branch was dropped"
    logits = ours(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()
    e, r = rel(logits.detach(), l32), rel(l16, l32)
    print(f"drop_path 0.3: logits rel {e:.3e} (reference bf16 {r:.3e})")
    assert e <= max(2e-2, 2.5 * r)
    bad = []
    for n, p in ours.named_parameters():
        ee, rr = rel(p.grad, g32[n]), rel(g16[n], g32[n])
        if ee > max(3e-2, 2.5 * rr):
            bad.append((n, ee, rr))
    assert not bad, bad[:8]
    ours.eval()
    with torch.no_grad():
        a, b = ours(x), ours(x)
    assert torch.equal(a, b)                       # no DropPath in eval mode; deterministic forward


def test_reference_finetuning_engine_drives_the_b200_classifier():
    """Drop-in at the engine level: the reference's OWN engine_for_finetuning.train_one_epoch (fp16 autocast,
    GradScaler-based loss scaler, soft-target cross-entropy, its own create_optimizer) runs unchanged on the B200 classifier,
    and the loss trajectory follows the same engine driving the reference's classifier from the same weights."""
    import types
    refrun = _refrun()
    ref = refrun.load()
    eng = refrun.load_finetune_engine()
    from mofo_b200 import modeling_finetune as mf
    from timm.loss import SoftTargetCrossEntropy
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    name, B, classes = "vit_small_patch16_224", 4, 174
    kw = dict(num_classes=classes, all_frames=16, tubelet_size=2, drop_rate=0.0, drop_path_rate=0.0, attn_drop_rate=0.0,
              use_mean_pooling=True, init_scale=1.0)
    ref_model = getattr(ref.modeling_finetune, name)(pretrained=False, **kw).to(dev)
    ours = mf.create_model(name, pretrained=False, drop_block_rate=None, **kw)
    ours.load_state_dict(ref_model.state_dict(), strict=True)
    ours = ours.to(dev)
    xs = [b[0] for b in refrun.synthetic_batches(B, 2, seed=17, device=dev)]
    g = torch.Generator().manual_seed(3)
    ys = [torch.randint(0, classes, (B,), generator=g) for _ in range(2)]
    steps = 6

    def mixup_fn(samples, targets):            # stands in for timm's Mixup: label smoothing only (the engine needs soft targets)
        return samples, torch.nn.functional.one_hot(targets, classes).float() * 0.9 + 0.1 / classes
    loader = [(xs[i % 2], ys[i % 2], None, None) for i in range(steps)]
    args = types.SimpleNamespace(data_set="SSV2")

    def run(model):
        losses = refrun.LossLog()
        opt = refrun.create_optimizer(model, lr=2e-4)
        scaler = refrun.make_scaler(dev)
        with refrun._quiet():
            eng.train_one_epoch(model, SoftTargetCrossEntropy(), loader, opt, dev, 0, scaler, args, max_norm=None, model_ema=None,
                                mixup_fn=mixup_fn, log_writer=None, start_steps=0, lr_schedule_values=None, wd_schedule_values=None,
                                num_training_steps_per_epoch=steps, update_freq=1)
        return model
    # per-step losses through a forward hook on the criterion is not available: compare the trained weights' behaviour instead
    run(ref_model); run(ours)
    ref_model.eval(); ours.eval()
    with torch.no_grad():
        a = ref_model(xs[0]).float(); b = ours(xs[0])
    la = torch.nn.functional.cross_entropy(a, ys[0].to(dev)).item(); lb = torch.nn.functional.cross_entropy(b, ys[0].to(dev)).item()
    print(f"after {steps} steps of the reference finetuning engine: eval loss reference {la:.4f}, ours {lb:.4f}")
    assert abs(la - lb) <= 3e-2 * abs(la)
    assert ((a - b).norm() / a.norm()).item() < 0.1


@pytest.mark.parametrize("fusing", ["weighted_mean", "org", "soft_attn", "MCA"])
def test_box_focused_classifier_matches_reference(fusing):
    """VisionTransformer_BB_focused.forward(x, BB): the token-in-box predicate is bit-equal to the reference's patch_yab
    construction (all-ones Conv3d over a painted clip), logits and gradients follow the reference; parameters of the fusing
    modules that the chosen method never touches get no gradient on either side."""
    refrun = _refrun()
    ref = refrun.load()
    from mofo_b200 import modeling_finetune as mf
    dev = torch.device("cuda", 0)
    torch.manual_seed(21)
    B, classes = 4, 97
    kw = dict(num_classes=classes, all_frames=16, tubelet_size=2, drop_rate=0.0, drop_path_rate=0.0, attn_drop_rate=0.0,
              use_mean_pooling=True, init_scale=1.0, fusing_method=fusing)
    ref_model = ref.modeling_finetune.vit_base_patch16_224_BB_focused(pretrained=False, **kw).to(dev).train()
    if fusing == "MCA":              # away from the init point (std 0.02 weights, zero biases: a nearly uniform softmax)
        with torch.no_grad():
            att = ref_model.local_MCA[0].attn
            att.q.weight.mul_(3.0); att.kv.weight.mul_(3.0)
            att.q_bias.normal_(0, 0.5); att.v_bias.normal_(0, 0.5)
    ours = mf.create_model("vit_base_patch16_224_BB_focused", pretrained=False, drop_block_rate=None, **kw)
    ours.load_state_dict(ref_model.state_dict(), strict=True)
    ours = ours.to(dev).train()
    x = refrun.synthetic_batches(B, 1, seed=19, device=dev)[0][0]
    y = torch.randint(0, classes, (B,), device=dev)
    g = torch.Generator().manual_seed(4)
    x1 = torch.randint(0, 150, (B, 16, 1), generator=g); y1 = torch.randint(0, 150, (B, 16, 1), generator=g)
    BB = torch.cat([x1, y1, x1 + torch.randint(1, 74, (B, 16, 1), generator=g), y1 + torch.randint(1, 74, (B, 16, 1), generator=g)], 2)
    BB[1] = torch.tensor([10, 10, 10, 40])                       # empty box: no token inside -> plain mean (:560-562)
    BB[2, :, :] = torch.tensor([0, 0, 224, 100])
    if fusing == "MCA":
        BB[3, :, :] = torch.tensor([0, 0, 224, 224])             # every token in the box: keys fall back to the box tokens (:131-133)
    BB = BB.to(dev)
    # the reference's predicate, as modeling_finetune.py:589-630 builds it
    with torch.no_grad():
        x_new = torch.zeros_like(x)
        for i in range(B):
            for j in range(16):
                x_new[i, :, j, BB[i, j, 1]:BB[i, j, 3], BB[i, j, 0]:BB[i, j, 2]] = 1
        want = torch.clamp(ref_model.patch_yab(x_new).flatten(2).transpose(1, 2).mean(2), 0, 1).type(torch.bool)
    got = ours.tokens_in_box(BB, 16, 224)
    assert torch.equal(got, want)
    assert int(want[1].sum()) == 0 and 0 < int(want[0].sum()) < 1568 and (fusing != "MCA" or int(want[3].sum()) == 1568)

    def ref_step(amp):
        ref_model.zero_grad(set_to_none=True)
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                logits = ref_model(x, BB)
                loss = torch.nn.functional.cross_entropy(logits.float(), y)
            loss.backward()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        return logits.detach().float(), {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in ref_model.named_parameters()}
    l32, g32 = ref_step(False)
    l16, g16 = ref_step(True)
    logits = ours(x, BB)
    torch.nn.functional.cross_entropy(logits, y).backward()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()
    e, r = rel(logits.detach(), l32), rel(l16, l32)
    print(f"box-focused ({fusing}): logits rel {e:.3e} (reference bf16 {r:.3e})")
    assert e <= max(2e-2, 2.5 * r)
    bad = []
    for n, p in ours.named_parameters():
        if g32[n] is None:
            assert p.grad is None, n
            continue
        if n.startswith("soft_att"):                 # 'soft_attn': a mathematically zero gradient (rounding noise in the reference)
            assert p.grad is None and g32[n].abs().max().item() < 1e-6, (n, g32[n].abs().max().item())
            continue
        ee, rr = rel(p.grad, g32[n]), rel(g16[n], g32[n])
        if ee > max(3e-2, 2.5 * rr):
            bad.append((n, ee, rr))
    assert not bad, bad[:8]
