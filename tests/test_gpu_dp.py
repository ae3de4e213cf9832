"""Data-parallel correctness on real GPUs (needs >= 2 CUDA devices; skipped otherwise): after the overlapped NCCL
all-reduce of the gradient arena, every rank holds the gradients of the CONCATENATED batch (mean over ranks), i.e. what a
single GPU computes on all clips — SURVEY.md §8e correctness check."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _inputs(cfg, B, seed):
    from oracle import mask_oracle as mo, target_oracle as tgt
    vid = tgt.synthetic_clip(B, seed=seed, size=cfg.img)
    boxes = tgt.synthetic_boxes(B, seed=seed + 1, size=cfg.img)
    masks = np.stack([mo.tube_mask_bb(boxes[b], mo.mt19937_words(50 + b, 400), cfg.grid)[0] for b in range(B)])
    return vid, torch.from_numpy(masks).bool()


def _build(cfg):
    from functools import partial
    from mofo_b200 import modeling_pretrain as mp_
    from oracle import model_oracle as mdl
    m = mp_.PretrainVisionTransformer(img_size=cfg.img, patch_size=16, encoder_embed_dim=cfg.enc_dim, encoder_depth=cfg.enc_depth,
                                      encoder_num_heads=cfg.enc_heads, decoder_embed_dim=cfg.dec_dim, decoder_depth=cfg.dec_depth,
                                      decoder_num_heads=cfg.dec_heads, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    m.load_state_dict(mdl.random_state_dict(cfg, seed=21, perturb=0.05))
    return m


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mofo_b200.dp import GradSync
        from oracle import model_oracle as mdl
        cfg = mdl.tiny_config(img=64, frames=16)
        vid, mask = _inputs(cfg, 2 * world, seed=400)
        model = _build(cfg).cuda()
        sync = GradSync()
        r = model._runner
        r._ensure_device(torch.device("cuda", rank))
        arena = r.grad_arena()
        sync.begin(arena, r.stage_end)
        sl = slice(2 * rank, 2 * rank + 2)
        loss = model.pretrain_step(vid[sl].cuda(), mask[sl].cuda(), grad_scale=sync.grad_scale, stage_done=sync.stage_done)
        sync.finish()
        torch.cuda.synchronize()
        q.put((rank, loss.item(), {n: p.grad.detach().cpu() for n, p in model.named_parameters()}, None))
    except Exception:
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_equal_single_gpu_on_concatenated_batch():
    from oracle import model_oracle as mdl
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for rank, loss, grads, tb in res:
        assert tb is None, tb
    cfg = mdl.tiny_config(img=64, frames=16)
    vid, mask = _inputs(cfg, 2 * world, seed=400)
    single = _build(cfg).cuda()
    single.use_cuda_graph = False
    l1 = single.pretrain_step(vid.cuda(), mask.cuda()).item()
    assert abs(0.5 * (res[0][1] + res[1][1]) - l1) < 1e-5 * abs(l1)
    for n, p in single.named_parameters():
        g1 = p.grad.detach().cpu()
        for rank in range(world):
            d = ((res[rank][2][n] - g1).norm() / g1.norm().clamp_min(1e-30)).item()
            # fp32 arena reductions; the 1/world factor folded into the bf16 dpred is a power of two (exact), so the only
            # difference is the fp32 summation order of the weight-gradient reductions (1e-5)
            assert d < 1e-5, (n, rank, d)
        assert torch.equal(res[0][2][n], res[1][2][n])   # both ranks hold identical reduced gradients
