"""Fused AdamW (mofo_adamw_step) vs torch.optim.AdamW on identical gradients, incl. weight-decay groups, per-step lr
changes, gradient clipping coefficient, the non-finite-loss guard and the bf16 operand copies it maintains."""
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model_oracle as mdl


def build():
    from mofo_b200 import modeling_pretrain as mp
    cfg = mdl.tiny_config(img=64, frames=16)
    m = mp.PretrainVisionTransformer(img_size=64, patch_size=16, encoder_embed_dim=cfg.enc_dim, encoder_depth=cfg.enc_depth,
                                     encoder_num_heads=cfg.enc_heads, decoder_embed_dim=cfg.dec_dim, decoder_depth=cfg.dec_depth,
                                     decoder_num_heads=cfg.dec_heads, mlp_ratio=4, qkv_bias=True,
                                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    m.load_state_dict(mdl.random_state_dict(cfg, seed=5, perturb=0.05))
    return m.cuda()


def groups(model):
    from mofo_b200.optim_factory import get_parameter_groups
    return get_parameter_groups(model, 0.05, model.no_weight_decay())


def test_fused_adamw_matches_torch_adamw():
    from mofo_b200.optim_factory import FusedAdamW
    torch.manual_seed(0)
    ours, ref = build(), build()
    opt = FusedAdamW(groups(ours), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05).attach(ours)
    opt_ref = torch.optim.AdamW(groups(ref), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    assert [g["weight_decay"] for g in opt.param_groups] == [g["weight_decay"] for g in opt_ref.param_groups]
    names = [n for n, _ in ours.named_parameters()]
    arena = ours._runner.grad_arena()
    gen = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        lr = 1e-3 * (1 + step)                       # the engine rewrites lr / wd every step (engine...:230-236)
        for g, gr in zip(opt.param_groups, opt_ref.param_groups):
            g["lr"] = gr["lr"] = lr * g.get("lr_scale", 1.0)
        coef = torch.tensor([0.5 if step == 3 else 1.0], device="cuda")
        arena.zero_()
        for (n, p), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
            gval = torch.randn(p.shape, device="cuda", generator=gen) * 0.1
            p.grad.copy_(gval)
            pr.grad = gval * coef
        opt.step(clip_coef=coef)
        opt_ref.step()
    worst = 0.0
    for (n, p), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        worst = max(worst, ((p - pr).abs().max() / pr.abs().max().clamp_min(1e-12)).item())
        m_ref = opt_ref.state[pr]["exp_avg"]
        assert torch.allclose(opt.state[p]["exp_avg"], m_ref, rtol=1e-5, atol=1e-8), n
    assert worst < 2e-6, worst
    # bf16 operand copies follow the parameters
    r = ours._runner
    blk = ours.encoder.blocks[1]
    wb, wt = r.wcache["enc1.fc1"]
    assert torch.equal(wb, blk.mlp.fc1.weight.detach().bfloat16()) and torch.equal(wt, blk.mlp.fc1.weight.detach().t().bfloat16())
    wb, wt = r.wcache["pe"]
    assert wt is None and torch.equal(wb, ours.encoder.patch_embed.proj.weight.detach().reshape(wb.shape).bfloat16())
    qb = r.buf("dec0.qkvbias", (3 * 64,), torch.float32)
    a = ours.decoder.blocks[0].attn
    assert torch.equal(qb, torch.cat([a.q_bias.detach(), torch.zeros_like(a.q_bias), a.v_bias.detach()]))
    # non-finite loss guard: update skipped entirely
    before = [p.detach().clone() for p in ours.parameters()]
    opt.step(loss_guard=torch.tensor([float("nan")], device="cuda"))
    assert all(torch.equal(a_, b_) for a_, b_ in zip(before, ours.parameters()))
    # state_dict has torch AdamW's layout and round-trips
    sd = opt.state_dict()
    assert set(sd["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"} and len(sd["state"]) == len(names)
    opt.load_state_dict(sd)
    assert opt._step == 7


def test_engine_with_fused_optimizer_tracks_torch_optimizer():
    """The same 8 steps through train_one_epoch_BB with FusedAdamW and with torch.optim.AdamW give the same losses."""
    import numpy as np
    from mofo_b200 import engine_for_pretraining as eng, utils as U
    from mofo_b200.optim_factory import FusedAdamW
    from oracle import mask_oracle as mo, target_oracle as tgt
    cfg = mdl.tiny_config(img=64, frames=16)
    vid = tgt.synthetic_clip(4, seed=9, size=64)
    boxes = tgt.synthetic_boxes(4, seed=10, size=64)
    masks = np.stack([mo.tube_mask_bb(boxes[b], mo.mt19937_words(b, 300), cfg.grid)[0] for b in range(4)])

    class Loader(list):
        quiet = True
    batch = (vid, torch.zeros(4, 16, 4, dtype=torch.long), torch.from_numpy(masks))
    losses = []
    for fused in (True, False):
        model = build()
        if fused:
            opt = FusedAdamW(groups(model), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
        else:
            opt = torch.optim.AdamW(groups(model), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
        per_step = []
        for _ in range(8):
            st = eng.train_one_epoch_BB(model, Loader([batch]), opt, torch.device("cuda"), 0, U.NativeScalerWithGradNormCount(),
                                        max_norm=1.0, start_steps=0)
            per_step.append(st["loss"])
        losses.append(per_step)
    a, b = np.asarray(losses[0]), np.asarray(losses[1])
    assert a[-1] < a[0]
    assert np.abs(a - b).max() < 2e-3 * np.abs(b).max(), (a, b)


def test_engine_pipelined_host_reports_the_same_meters_as_strict_sync(monkeypatch, capsys):
    """One step of host/device pipelining (loss read a step late) must not change what the engine reports: every
    step's loss / grad norm reaches the meters, and a non-finite loss still exits with the reference's message."""
    import numpy as np
    from mofo_b200 import engine_for_pretraining as eng, utils as U
    from mofo_b200.optim_factory import FusedAdamW
    from oracle import mask_oracle as mo, target_oracle as tgt
    cfg = mdl.tiny_config(img=64, frames=16)
    boxes = tgt.synthetic_boxes(4, seed=10, size=64)
    masks = torch.from_numpy(np.stack([mo.tube_mask_bb(boxes[b], mo.mt19937_words(b, 300), cfg.grid)[0] for b in range(4)]))

    class Loader(list):
        quiet = True
    batches = Loader([(tgt.synthetic_clip(4, seed=20 + i, size=64).pin_memory(), torch.zeros(4, 16, 4, dtype=torch.long), masks)
                      for i in range(7)])
    stats = []
    for strict in ("1", "0"):
        monkeypatch.setenv("MOFO_SYNC_EVERY_STEP", strict)
        model = build()
        opt = FusedAdamW(groups(model), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
        stats.append(eng.train_one_epoch_BB(model, batches, opt, torch.device("cuda"), 0, U.NativeScalerWithGradNormCount(),
                                            max_norm=1.0, start_steps=0))
    for k in ("loss", "grad_norm", "lr", "weight_decay"):
        # two runs differ by the summation order of the fp32 atomics (wgrad / LayerNorm reductions): ~1e-6 relative
        assert abs(stats[0][k] - stats[1][k]) <= 1e-4 * abs(stats[0][k]) + 1e-12, (k, stats)
    # non-finite loss: same message and exit status as engine_for_pretraining.py:418-420, one step late
    monkeypatch.setenv("MOFO_SYNC_EVERY_STEP", "0")
    bad = Loader(list(batches[:3]))
    v = bad[1][0].clone(); v[0, 0, 0, 0, 0] = float("nan")
    bad[1] = (v.pin_memory(), bad[1][1], bad[1][2])
    model = build()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt = FusedAdamW(groups(model), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
    with pytest.raises(SystemExit) as ex:
        eng.train_one_epoch_BB(model, bad, opt, torch.device("cuda"), 0, U.NativeScalerWithGradNormCount(), max_norm=1.0)
    assert ex.value.code == 1
    assert "stopping training" in capsys.readouterr().out
    assert all(torch.isfinite(p).all() for p in model.parameters())      # the device-side guard skipped the bad update
    assert any(not torch.equal(before[n], p) for n, p in model.named_parameters())   # step 0 was applied


def test_example_script_trains_and_checkpoint_round_trips(tmp_path, monkeypatch):
    """examples/pretrain_synthetic.py = the reference's main() flow on synthetic data; the checkpoint it writes has the
    reference's state_dict schema and loads back into a fresh model + optimizer."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("pretrain_synthetic", os.path.join(root, "examples", "pretrain_synthetic.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["x", "--model", "pretrain_mae_small_patch16_224", "--batch_size", "2", "--epochs", "2",
                                      "--steps_per_epoch", "3", "--output_dir", str(tmp_path)])
    stats = mod.main()
    assert np.isfinite(stats["loss"]) and stats["lr"] > 0
    ck = torch.load(os.path.join(tmp_path, "checkpoint-1.pth"), map_location="cpu", weights_only=False)
    assert list(ck["model"].keys()) == list(mdl.param_shapes(mdl.CONFIGS["pretrain_mae_small_patch16_224"]).keys())
    from mofo_b200 import modeling_pretrain as mp
    from mofo_b200.optim_factory import FusedAdamW, get_parameter_groups
    m2 = mp.create_model("pretrain_mae_small_patch16_224", decoder_depth=4).cuda()
    m2.load_state_dict(ck["model"])
    o2 = FusedAdamW(get_parameter_groups(m2, 0.05, m2.no_weight_decay()), lr=1e-4, betas=(0.9, 0.95)).attach(m2)
    o2.load_state_dict(ck["optimizer"])
    assert o2._step == 6


import numpy as np  # noqa: E402
