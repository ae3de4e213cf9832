"""B200-native look-alike of the reference's ``optim_factory.py`` for the pretraining path.

``create_optimizer(args, model, ...)`` and ``get_parameter_groups`` keep the reference's signatures and grouping
rules (optim_factory.py:49-127: no weight decay for 1-D parameters, ``.bias`` and the model's ``no_weight_decay()``
set; ``lr_scale`` per group).  For ``--opt adamw`` (the pretraining recipe, PRETRAIN.md) it returns ``FusedAdamW``:
a ``torch.optim.Optimizer`` whose ``step()`` is ONE kernel (``mofo_adamw_step``) over flat fp32 arenas that also
emits the bf16 operand copies (W, W^T) the next forward/backward reads — replacing torch's multi-tensor AdamW, the
per-step weight casts and the qkv-bias packing.  ``param_groups`` (lr / weight_decay edited per step by the engine,
engine_for_pretraining.py:230-236) and ``state_dict()`` (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``) keep
torch.optim.AdamW's layout, so reference checkpoints of the optimizer round-trip.
"""
from __future__ import annotations

import json
import math

import numpy as np
import torch

from . import _lib


def get_parameter_groups(model, weight_decay=1e-5, skip_list=(), get_num_layer=None, get_layer_scale=None, verbose=False):
    """optim_factory.py:49-88"""
    names, groups = {}, {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or name in skip_list:
            group_name, this_wd = "no_decay", 0.
        else:
            group_name, this_wd = "decay", weight_decay
        layer_id = None
        if get_num_layer is not None:
            layer_id = get_num_layer(name)
            group_name = "layer_%d_%s" % (layer_id, group_name)
        if group_name not in names:
            scale = get_layer_scale(layer_id) if get_layer_scale is not None else 1.
            names[group_name] = {"weight_decay": this_wd, "params": [], "lr_scale": scale}
            groups[group_name] = {"weight_decay": this_wd, "params": [], "lr_scale": scale}
        groups[group_name]["params"].append(param)
        names[group_name]["params"].append(name)
    if verbose:
        print("Param groups = %s" % json.dumps(names, indent=2))
    return list(groups.values())


class FusedAdamW(torch.optim.Optimizer):
    """AdamW with torch.optim.AdamW semantics, executed by one CUDA kernel over flat arenas.

    ``attach(model)`` (called by ``create_optimizer`` / lazily by the engine) re-points every parameter's storage into a
    flat fp32 arena laid out exactly like the model's gradient arena, allocates the moment arenas and the bf16 operand
    arena, and hands the bf16 views to the model's kernel runner.  Values are preserved."""

    fused_mofo = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._attached = None
        self._step = 0

    # ---- wiring -------------------------------------------------------------------------------------------
    def attach(self, model):
        core = model.module if hasattr(model, "module") else model
        if self._attached is core:
            return self
        runner = core._runner
        dev = next(core.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA only: move the model to the GPU before creating / attaching the optimizer")
        runner._ensure_device(dev)
        named = dict(core.named_parameters())
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                group_of[id(p)] = gi
        missing = [n for n, p in named.items() if id(p) not in group_of]
        if missing:
            raise RuntimeError(f"FusedAdamW must own every model parameter; missing e.g. {missing[:3]}")
        if len(self.param_groups) > 28:
            raise RuntimeError("FusedAdamW supports at most 28 parameter groups")
        garena = runner.grad_arena()
        gviews = runner.arena_views
        n = garena.numel()
        self.p_arena = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m_arena = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v_arena = torch.zeros(n, dtype=torch.float32, device=dev)
        two_d = {}   # name -> key in runner.wcache, needs transposed copy
        two_d["encoder.patch_embed.proj.weight"] = ("pe", False)
        for tag, blocks in (("enc", core.encoder.blocks), ("dec", core.decoder.blocks)):
            for i, blk in enumerate(blocks):
                for sub, key in (("attn.qkv", "qkv"), ("attn.proj", "proj"), ("mlp.fc1", "fc1"), ("mlp.fc2", "fc2")):
                    two_d[f"{blk._mofo_name}.{sub}.weight"] = (f"{tag}{i}.{key}", True)
        two_d["encoder_to_decoder.weight"] = ("e2d", True)
        two_d["decoder.head.weight"] = ("head", True)
        segs, tiles, w16_elems = [], [], 0
        order, _, _ = runner.backward_order()
        w16_plan = {}
        for name in order:
            p = named[name]
            gv = gviews[name]
            off = (gv.data_ptr() - garena.data_ptr()) // 4
            if name in two_d:
                rows = p.shape[0]; cols = p.numel() // rows
                key, need_t = two_d[name]
                w_off = w16_elems; w16_elems += rows * cols
                wt_off = -1
                if need_t:
                    wt_off = w16_elems; w16_elems += rows * cols
                w16_plan[key] = (w_off, wt_off, rows, cols)
            else:
                rows, cols, w_off, wt_off = 1, p.numel(), -1, -1
            seg = len(segs)
            segs.append([off, rows, cols, group_of[id(p)], w_off, wt_off])
            tr, tc = _lib.ADAMW_TILE
            ntile = ((rows + tr - 1) // tr) * ((cols + tc - 1) // tc) if wt_off >= 0 else (rows * cols + _lib.ADAMW_RUN - 1) // _lib.ADAMW_RUN
            tiles.extend([seg, t] for t in range(ntile))
        self.w16 = torch.zeros(max(w16_elems, 8), dtype=torch.bfloat16, device=dev)
        self.segs = torch.tensor(segs, dtype=torch.int64, device=dev)
        self.tiles = torch.tensor(tiles, dtype=torch.int32, device=dev)
        # tiles follow the arena (= backward completion) order, so the tiles of gradient-sync stage k are the contiguous
        # range [stage_tile_end[k-1], stage_tile_end[k]): the staged step updates a stage as soon as its slice is reduced
        self.stage_tile_end = []
        for end in runner.stage_end:
            self.stage_tile_end.append(sum(1 for sg, _ in tiles if segs[sg][0] < end))
        # move parameter storage into the arena (values preserved); moments become views too (state_dict layout of torch AdamW)
        with torch.no_grad():
            for name, p in named.items():
                gv = gviews[name]
                off = (gv.data_ptr() - garena.data_ptr()) // 4
                view = self.p_arena[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                st = self.state[p]
                old_m, old_v = st.get("exp_avg"), st.get("exp_avg_sq")
                st["exp_avg"] = self.m_arena[off:off + p.numel()].view(p.shape)
                st["exp_avg_sq"] = self.v_arena[off:off + p.numel()].view(p.shape)
                if old_m is not None:
                    st["exp_avg"].copy_(old_m); st["exp_avg_sq"].copy_(old_v)
                st.setdefault("step", torch.tensor(float(self._step)))
        # hand the bf16 operand views (and the packed qkv-bias windows of the parameter arena) to the runner
        wc = {}
        for key, (w_off, wt_off, rows, cols) in w16_plan.items():
            wb = self.w16[w_off:w_off + rows * cols].view(rows, cols)
            wt = self.w16[wt_off:wt_off + rows * cols].view(cols, rows) if wt_off >= 0 else None
            wc[key] = (wb, wt)
        qkvbias = {}
        for tag, blocks in (("enc", core.encoder.blocks), ("dec", core.decoder.blocks)):
            for i, blk in enumerate(blocks):
                qb = blk.attn.q_bias
                D = qb.numel()
                off = (qb.data_ptr() - self.p_arena.data_ptr()) // 4
                assert blk.attn.v_bias.data_ptr() == qb.data_ptr() + 8 * D, "arena must hold [q_bias | gap | v_bias]"
                qkvbias[f"{tag}{i}.qkvbias"] = self.p_arena[off:off + 3 * D]
        runner.adopt_external_weights(wc, qkvbias)
        # ring of pinned staging buffers: the async H2D copy of step i reads host memory when it EXECUTES, which may be
        # after the host has started preparing step i+1 (the engine pipelines the host one step ahead of the device)
        self.hyper_hosts = [torch.zeros(8 + 2 * len(self.param_groups), dtype=torch.float32).pin_memory() for _ in range(4)]
        self.hyper_host = self.hyper_hosts[0]
        self.hyper = torch.zeros_like(self.hyper_host, device=dev)
        self._attached = core
        self._runner = runner
        return self

    # ---- staged step: one launch per gradient-sync stage, driven by mofo_b200.dp.GradSync ------------------------
    @torch.no_grad()
    def begin_staged(self, loss_guard=None):
        """Call BEFORE the step's backward is enqueued: advances the step count, uploads this step's hyper-parameters and
        returns ``(on_stage, sq_norm_acc)``.  ``on_stage(k, lo, hi)`` (run by GradSync on its stream once stage k's slice
        is final) adds that slice's squared norm to ``sq_norm_acc`` and updates its parameters / bf16 operand copies;
        stage k's weights are not read again in this step once its gradients exist.  No clipping on this path (the
        clip coefficient needs the norm of ALL slices first): the scaler falls back to ``step()`` when clipping is on."""
        if self._attached is None:
            raise RuntimeError("FusedAdamW.begin_staged() before attach(model)")
        garena = self._runner.grad_arena()
        self._upload_hyper()
        acc = torch.zeros(1, dtype=torch.float32, device=garena.device)
        ends = self.stage_tile_end

        def on_stage(k, lo, hi):
            t0 = ends[k - 1] if k > 0 else 0
            if ends[k] > t0:
                _lib.adamw_step(self.p_arena, garena, self.m_arena, self.v_arena, self.w16, self.segs, self.tiles[t0:ends[k]],
                                self.hyper, None, loss_guard, sq_norm_out=acc)
        return on_stage, acc

    def _upload_hyper(self):
        self._step += 1
        beta1, beta2 = self.param_groups[0]["betas"]
        h = self.hyper_host = self.hyper_hosts[self._step % len(self.hyper_hosts)]
        h[0], h[1], h[2] = beta1, beta2, self.param_groups[0]["eps"]
        h[3] = 1.0 - beta1 ** self._step
        h[4] = math.sqrt(1.0 - beta2 ** self._step)
        for gi, g in enumerate(self.param_groups):
            h[8 + 2 * gi] = g["lr"]
            h[9 + 2 * gi] = g["weight_decay"]
        self.hyper.copy_(h, non_blocking=True)

    # ---- torch.optim.Optimizer API ------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, clip_coef=None, loss_guard=None, sq_norm_out=None):
        if self._attached is None:
            raise RuntimeError("FusedAdamW.step() before attach(model)")
        loss = closure() if closure is not None else None
        runner = self._runner
        garena = runner.grad_arena()
        for name, p in self._attached.named_parameters():      # gradients produced outside the arena (autograd path)
            gv = runner.arena_views[name]
            if p.grad is not None and p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
                p.grad = gv
        self._upload_hyper()
        _lib.adamw_step(self.p_arena, garena, self.m_arena, self.v_arena, self.w16, self.segs, self.tiles, self.hyper,
                        clip_coef, loss_guard, sq_norm_out=sq_norm_out)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        if self._attached is not None:
            self._runner.grad_arena().zero_()
        else:
            super().zero_grad(set_to_none)

    def state_dict(self):
        for st in self.state.values():
            if "step" in st:
                st["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        views = {id(p): (st.get("exp_avg"), st.get("exp_avg_sq")) for p, st in self.state.items()}
        super().load_state_dict(state_dict)
        steps = []
        with torch.no_grad():
            for p, st in self.state.items():
                m, v = views.get(id(p), (None, None))
                if m is not None and "exp_avg" in st and st["exp_avg"].data_ptr() != m.data_ptr():
                    m.copy_(st["exp_avg"]); v.copy_(st["exp_avg_sq"])
                    st["exp_avg"], st["exp_avg_sq"] = m, v
                if "step" in st:
                    steps.append(int(float(st["step"])))
        if steps:
            self._step = max(steps)


def create_optimizer(args, model, get_num_layer=None, get_layer_scale=None, filter_bias_and_bn=True, skip_list=None):
    """optim_factory.py:91-175 for the optimizer the pretraining recipe uses (``--opt adamw``)."""
    opt_lower = args.opt.lower()
    weight_decay = args.weight_decay
    if weight_decay and filter_bias_and_bn:
        skip = {}
        if skip_list is not None:
            skip = skip_list
        elif hasattr(model, 'no_weight_decay'):
            skip = model.no_weight_decay()
        parameters = get_parameter_groups(model, weight_decay, skip, get_num_layer, get_layer_scale)
        weight_decay = 0.
    else:
        parameters = model.parameters()
    opt_args = dict(lr=args.lr, weight_decay=weight_decay)
    if hasattr(args, 'opt_eps') and args.opt_eps is not None:
        opt_args['eps'] = args.opt_eps
    if hasattr(args, 'opt_betas') and args.opt_betas is not None:
        opt_args['betas'] = tuple(args.opt_betas)
    if opt_lower.split('_')[-1] != 'adamw':
        raise NotImplementedError(f"mofo_b200.optim_factory implements the pretraining recipe's optimizer (adamw); got {args.opt!r}")
    opt = FusedAdamW(parameters, **opt_args)
    if next(model.parameters()).is_cuda:
        opt.attach(model)
    return opt
