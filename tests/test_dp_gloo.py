"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes drive mofo_b200.dp.GradSync over an
arena laid out exactly as the model lays it out (backward-completion order, stage boundaries), and the result must be
the rank-sum of every slice with the 1/world factor folded in.  No GPU and no kernel is involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mofo_b200.dp import GradSync
        from mofo_b200.modeling_pretrain import create_model
        torch.manual_seed(0)
        model = create_model("pretrain_mae_small_patch16_224", decoder_depth=4)
        runner = model._runner
        runner.device = torch.device("cpu")
        arena, views = runner._make_arena()
        stage_end = runner.stage_end
        order, stage_of, n_stages = runner.backward_order()
        assert stage_end == sorted(stage_end) and stage_end[-1] == arena.numel()
        # every parameter lives wholly inside its stage's slice
        lo = [0] + stage_end[:-1]
        for n in order:
            v = views[n]
            off = (v.data_ptr() - arena.data_ptr()) // 4
            k = stage_of[n]
            assert lo[k] <= off and off + v.numel() <= stage_end[k], n
        sync = GradSync()
        assert sync.world == world and abs(sync.grad_scale - 1.0 / world) < 1e-12
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(arena.numel(), generator=g)
        arena.copy_(local * sync.grad_scale)            # the loss gradient carries the 1/world factor
        sync.begin(arena, stage_end)
        calls = []
        for k in range(n_stages):                        # the order the backward pass reports stages
            sync.stage_done(k)
            calls.append(sync.prev_end)
        sync.finish()
        assert calls == stage_end and sync.launched == n_stages
        expect = sum(torch.randn(arena.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        err = (arena - expect).abs().max().item()
        # a stage that never reports is still flushed by finish()
        arena.copy_(local * sync.grad_scale)
        sync.begin(arena, stage_end)
        sync.stage_done(0)
        sync.finish()
        err2 = (arena - expect).abs().max().item()
        # on_stage hook: called once per stage, in order, with that stage's slice bounds, AFTER the slice was reduced
        arena.copy_(local * sync.grad_scale)
        seen = []
        sync.begin(arena, stage_end, on_stage=lambda k, lo, hi: seen.append((k, lo, hi, (arena[lo:hi] - expect[lo:hi]).abs().max().item())))
        for k in range(n_stages):
            sync.stage_done(k)
        sync.finish()
        assert [(k, lo, hi) for k, lo, hi, _ in seen] == [(k, ([0] + stage_end)[k], stage_end[k]) for k in range(n_stages)]
        assert all(e < 1e-6 for *_, e in seen)
        # parameters: replicas initialised from different seeds adopt rank 0's values (DDP constructor semantics)
        torch.manual_seed(1000 + rank)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        sync.sync_parameters(model)
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered) and getattr(model, "_mofo_params_synced", False)
        q.put((rank, err, err2, None))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_gradsync_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, err2, tb in res:
        assert tb is None, tb
        assert err < 1e-6 and err2 < 1e-6, (rank, err, err2)


def test_single_process_sync_is_a_noop():
    from mofo_b200.dp import GradSync
    s = GradSync()
    a = torch.arange(10, dtype=torch.float32)
    s.begin(a, [4, 10])
    s.stage_done(0); s.stage_done(1); s.finish()
    assert s.world == 1 and s.grad_scale == 1.0 and torch.equal(a, torch.arange(10, dtype=torch.float32))


def test_state_dict_schema_matches_reference_layout():
    """state_dict keys / shapes / order equal the reference's (oracle.param_shapes is pinned to it by the golden test)."""
    from mofo_b200.modeling_pretrain import create_model
    from oracle import model_oracle as mdl
    for name, cfg in mdl.CONFIGS.items():
        if "large" in name:
            continue
        m = create_model(name, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4)
        sd = m.state_dict()
        want = mdl.param_shapes(cfg)
        assert list(sd.keys()) == list(want.keys())
        assert all(tuple(sd[k].shape) == tuple(want[k]) for k in want)
        assert m.encoder.patch_embed.patch_size == (16, 16)
        assert m.no_weight_decay() == {'pos_embed', 'cls_token', 'mask_token'}
        assert not any(k.endswith("pos_embed") for k in sd)          # tables are not buffers (modeling_pretrain.py:42)


def test_model_refuses_cpu_inputs():
    from mofo_b200.modeling_pretrain import create_model
    m = create_model("pretrain_mae_small_patch16_224", decoder_depth=4)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 16, 224, 224), torch.zeros(1, 1568, dtype=torch.bool))


def test_backward_order_stages_cover_every_parameter_once(monkeypatch):
    """Gradient-sync stages (mofo_b200/modeling_pretrain.py::_Runner.backward_order): decoder first, encoder blocks from
    the last to the first in groups, patch embedding with the final group; every parameter in exactly one stage, stage
    indices non-decreasing along the order (so every stage is a contiguous arena slice), for the default grouping and
    for explicit group sizes."""
    from mofo_b200.modeling_pretrain import create_model
    m = create_model("pretrain_videomae_base_patch16_224", decoder_depth=4)
    names = [n for n, _ in m.named_parameters()]
    for spec, want_sizes in ((None, [4, 4, 4]), ("5,4,3", [5, 4, 3]), ("4,4,2,2", [4, 4, 2, 2]), ("7", [7, 5])):
        if spec is None:
            monkeypatch.delenv("MOFO_ENC_STAGES", raising=False)
        else:
            monkeypatch.setenv("MOFO_ENC_STAGES", spec)
        order, stage_of, n_stages = m._runner.backward_order()
        assert sorted(order) == sorted(names) and len(set(order)) == len(order)
        stages = [stage_of[n] for n in order]
        assert stages == sorted(stages) and stages[0] == 0 and stages[-1] == n_stages - 1
        assert all(stage_of[n] == 0 for n in names if n.startswith(("decoder.", "mask_token", "encoder_to_decoder", "encoder.norm")))
        table = m._runner.enc_stage_table()
        sizes = [sum(1 for i in table if table[i] == s) for s in range(1, n_stages)]
        assert sizes == want_sizes and table[11] == 1
        assert stage_of["encoder.patch_embed.proj.weight"] == table[0] == n_stages - 1
