"""B200-native look-alike of the reference's ``modeling_pretrain.py``.

Same public surface as /root/reference/modeling_pretrain.py (:163-338):
  * ``PretrainVisionTransformer(...)`` with the same constructor keywords, ``forward(x, mask)``,
    ``no_weight_decay()``, ``encoder.patch_embed.patch_size`` and — key for checkpoints — the same 218
    ``state_dict`` names and shapes, so reference checkpoints load and finetuning's
    ``encoder.``-prefix stripping still works;
  * registry entry points ``pretrain_mae_small_patch16_224``, ``pretrain_videomae_base_patch16_224``,
    ``pretrain_videomae_large_patch16_224`` (registered with timm when timm is importable).

The nn.Module tree only HOLDS parameters (fp32 masters).  All arithmetic — forward and backward — runs in
hand-written sm_100a kernels behind the C ABI of libmofo_sm100.so (include/mofo_b200.h), orchestrated by
``_Runner`` below.  There is no PyTorch fallback: on a machine without CUDA ``forward`` raises.

Two entry points share the runner:
  * ``forward(x, mask)``: drop-in; returns predictions [B, N_mask, 1536] (bf16) wired into autograd through a
    single ``autograd.Function`` whose backward is the manual kernel backward.
  * ``pretrain_step(videos, mask | index lists)``: the fused step the engine uses — forward, target + MSE and
    backward in one pass, gradients accumulated straight into the flat gradient arena (``.grad`` views).

Numerics (vs. the reference under autocast): GEMM inputs bf16, fp32 accumulation; the residual
stream, LayerNorm statistics, softmax and the loss are fp32; parameters, gradients and the arena are fp32.
"""
from __future__ import annotations

import math
import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["PretrainVisionTransformer", "pretrain_mae_small_patch16_224", "pretrain_videomae_base_patch16_224",
           "pretrain_videomae_large_patch16_224", "get_sinusoid_encoding_table", "create_model"]


def get_sinusoid_encoding_table(n_position, d_hid):
    """Sinusoid table, f64 numpy then FloatTensor [1,n,d] (same arithmetic as modeling_finetune.py:252-262,
    vectorised)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.FloatTensor(tab).unsqueeze(0)


def _trunc_normal_(tensor, mean=0., std=1.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=-std, b=std)   # modeling_pretrain.py:13-14


# ----------------------------------------------------------------------------------------------------------
# parameter holders (names / shapes / init identical to the reference modules; their forward is never used)
# ----------------------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            raise NotImplementedError("mofo_b200 implements the registry models, which all use qkv_bias=True")
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, num_frames=16, tubelet_size=2):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.tubelet_size = int(tubelet_size)
        self.num_patches = (img_size // patch_size) ** 2 * (num_frames // self.tubelet_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(self.tubelet_size, patch_size, patch_size),
                              stride=(self.tubelet_size, patch_size, patch_size))


def _init_weights(m):
    if isinstance(m, nn.Linear):                      # modeling_pretrain.py:60-67
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class _Encoder(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, qkv_bias, norm_layer,
                 tubelet_size):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim, tubelet_size=tubelet_size)
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.apply(_init_weights)

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}


class _Decoder(nn.Module):
    def __init__(self, patch_size, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias, norm_layer,
                 tubelet_size):
        super().__init__()
        assert num_classes == 3 * tubelet_size * patch_size ** 2          # modeling_pretrain.py:112
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.apply(_init_weights)

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}


# ----------------------------------------------------------------------------------------------------------
# kernel orchestration
# ----------------------------------------------------------------------------------------------------------
class _Runner:
    """Owns the device-side state of one model: bf16 weight copies, position tables, activation workspace and the
    flat gradient arena, and sequences the C-ABI calls of one forward / backward."""

    def __init__(self, model: "PretrainVisionTransformer"):
        self.m = model
        self.bufs = {}
        self.wcache = {}
        self.wversion = None
        self.device = None
        self.arena = None
        self.arena_views = None
        self.scratch_arena = None
        self.enc_group = int(os.environ.get("MOFO_ENC_GROUP", "4"))   # encoder blocks per gradient-sync stage
        self.stage_end = None
        self.graphs = {}            # (B, Nv, Nm, normalize, grad_scale) -> [CUDAGraph per sync stage] | None (capture failed)
        self.static_in = {}         # (B, Nv, Nm, T, size) -> (videos, vis_idx, msk_idx) device buffers the graphs read
        self.graph_pool = None
        self.force_cast = False
        self.external_weights = False   # bf16 operand copies maintained by mofo_b200.optim_factory.FusedAdamW
        # weight-gradient GEMMs are off the backward critical path (nothing reads dW before the stage's all-reduce), so
        # they CAN run on a side stream and fill SMs that the dgrad / attention / LayerNorm chain leaves idle at kernel
        # tails.  Measured on B200 (B=32 ViT-B): 13.31 -> 13.27 ms/step, i.e. within run-to-run noise - a wgrad CTA
        # needs ~190 KB of shared memory and rarely finds an SM to co-reside on - so it is opt-in: MOFO_SIDE_WGRAD=1.
        self.side_wgrad = os.environ.get("MOFO_SIDE_WGRAD", "0") == "1"
        # the four weight gradients of a block (fc2, fc1, proj, qkv) in ONE grouped launch at the end of the block's backward
        # (their inputs are all still alive there); MOFO_GROUPED_WGRAD=0 launches them one by one where they arise
        self.grouped_wgrad = os.environ.get("MOFO_GROUPED_WGRAD", "1") == "1" and not self.side_wgrad
        self._side = None
        self._side_dirty = False
        self._readers = {}          # data_ptr of a scratch tensor -> event after the side-stream wgrad that reads it

    # ---- memory -------------------------------------------------------------------------------------------
    def buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(*shape, dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t

    # ---- side stream for the weight-gradient GEMMs --------------------------------------------------------
    def _wgrad(self, reads, *args, **kw):
        """``_lib.gemm_wgrad(*args, **kw)`` ordered after everything enqueued so far, on the side stream.  ``reads``:
        the scratch tensors it reads that the main stream overwrites later (guarded by ``_before_write``)."""
        if not self.side_wgrad:
            _lib.gemm_wgrad(*args, **kw)
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            _lib.gemm_wgrad(*args, **kw)
            done = torch.cuda.Event()
            done.record(self._side)
        for t in reads:
            self._readers[t.data_ptr()] = done
        self._side_dirty = True

    def _before_write(self, *tensors):
        """The main stream is about to overwrite these scratch tensors: wait for their side-stream readers."""
        if not self._readers:
            return
        main = torch.cuda.current_stream(self.device)
        for t in tensors:
            ev = self._readers.pop(t.data_ptr(), None)
            if ev is not None:
                main.wait_event(ev)

    def _join_side(self):
        """Main stream waits for every weight gradient enqueued so far (before a stage's all-reduce / the optimizer)."""
        if self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._readers.clear()
            self._side_dirty = False

    def _ensure_device(self, device):
        if self.device != device:
            self.device = device
            self.bufs.clear(); self.wcache.clear(); self.wversion = None
            self.arena = None; self.arena_views = None; self.scratch_arena = None
            self.graphs.clear(); self.static_in.clear(); self.graph_pool = None
            m = self.m
            self.pos_enc = m.encoder_pos_embed[0].to(device).contiguous()
            self.pos_dec = m.pos_embed[0].to(device).contiguous()

    def enc_stage_table(self):
        """block index -> gradient-sync stage (1-based; stage 0 is the decoder).  Blocks are grouped from the last
        (first to finish in backward) to the first.  Default: groups of ``enc_group``; ``MOFO_ENC_STAGES="4,4,2,2"``
        gives explicit group sizes (a smaller final group shortens the all-reduce that nothing overlaps)."""
        ne = len(self.m.encoder.blocks)
        spec = os.environ.get("MOFO_ENC_STAGES")
        sizes = [max(1, int(v)) for v in spec.split(",")] if spec else [self.enc_group] * ((ne + self.enc_group - 1) // self.enc_group)
        table, i, stage = {}, ne - 1, 1
        while i >= 0:
            sz = sizes[min(stage - 1, len(sizes) - 1)]
            for _ in range(sz):
                if i >= 0:
                    table[i] = stage
                    i -= 1
            stage += 1
        return table

    def backward_order(self):
        """Parameter names in the order the backward pass finishes their gradients, with the stage index at which
        each becomes final: stage 0 = decoder (+ head, mask_token, encoder_to_decoder), then encoder blocks in
        groups of ``enc_group`` from the last to the first, the patch embedding riding with the final group."""
        m = self.m
        names = dict(m.named_parameters())
        order, stage_of = [], {}

        def add(prefix, stage):
            for n in names:
                if (n == prefix or n.startswith(prefix + ".")) and n not in stage_of:
                    order.append(n); stage_of[n] = stage

        add("decoder.head", 0); add("decoder.norm", 0)
        for i in range(len(m.decoder.blocks) - 1, -1, -1):
            add(f"decoder.blocks.{i}", 0)
        add("mask_token", 0); add("encoder_to_decoder", 0); add("encoder.norm", 0)
        ne = len(m.encoder.blocks)
        table = self.enc_stage_table()
        for i in range(ne - 1, -1, -1):
            add(f"encoder.blocks.{i}", table[i])
        last = table[0]
        add("encoder.patch_embed", last)
        assert len(order) == len(names)
        return order, stage_of, last + 1

    def _named(self):
        """(name, parameter) pairs the kernels produce gradients for (subclasses exclude parameters their path never touches)."""
        return list(self.m.named_parameters())

    def _make_arena(self):
        params = dict(self._named())
        order, stage_of, n_stages = self.backward_order()
        total = sum((params[n].numel() + 3) // 4 * 4 * (2 if n.endswith(".attn.q_bias") else 1) for n in order)  # 16-B aligned
        arena = torch.zeros(total, dtype=torch.float32, device=self.device)
        views, off = {}, 0
        stage_end = [0] * n_stages
        for n in order:
            p = params[n]
            views[n] = arena[off:off + p.numel()].view(p.shape)
            off += (p.numel() + 3) // 4 * 4
            if n.endswith(".attn.q_bias"):
                off += (p.numel() + 3) // 4 * 4       # always-zero gap: [q_bias | gap | v_bias] is one [3D] window
            stage_end[stage_of[n]] = off
        self.stage_end = stage_end
        return arena, views

    def grad_arena(self):
        """Flat fp32 gradient arena; every parameter's ``.grad`` is a view into it (reverse-order slices make it
        directly usable for bucketed NCCL all-reduce)."""
        if self.arena is None:
            self.arena, self.arena_views = self._make_arena()
        for n, p in self._named():
            v = self.arena_views[n]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
        return self.arena

    # ---- weights ------------------------------------------------------------------------------------------
    def adopt_external_weights(self, wcache, qkvbias):
        """Called by FusedAdamW.attach: from now on the optimizer kernel rewrites the bf16 W / W^T copies in place after
        every update (and the packed qkv bias is a window of the parameter arena), so ``prepare_weights`` only has to
        re-cast when something else modified the fp32 parameters (e.g. ``load_state_dict``)."""
        self.wcache = dict(wcache)
        for k, v in qkvbias.items():
            self.bufs[(k, tuple(v.shape), torch.float32)] = v
        self.external_weights = True
        self.external_qkvbias = set(qkvbias)
        self.wversion = None
        self.graphs.clear()

    def prepare_weights(self):
        m = self.m
        version = tuple(p._version for p in m.parameters())
        if version == self.wversion and not (self.force_cast and not self.external_weights):
            return
        self.wversion = version
        wc = self.wcache

        def cast(name, W, need_t=True):
            R = W.shape[0]
            C = W.numel() // R
            if name not in wc:
                wc[name] = (torch.empty(R, C, dtype=torch.bfloat16, device=self.device),
                            torch.empty(C, R, dtype=torch.bfloat16, device=self.device) if need_t else None)
            wb, wt = wc[name]
            _lib.cast_weight(W.detach(), wb, wt)

        cast("pe", m.encoder.patch_embed.proj.weight, need_t=False)
        for tag, blocks in (("enc", m.encoder.blocks), ("dec", m.decoder.blocks)):
            for i, blk in enumerate(blocks):
                pre = f"{tag}{i}"
                cast(pre + ".qkv", blk.attn.qkv.weight)
                cast(pre + ".proj", blk.attn.proj.weight)
                cast(pre + ".fc1", blk.mlp.fc1.weight)
                cast(pre + ".fc2", blk.mlp.fc2.weight)
                if not (self.external_weights and (pre + ".qkvbias") in self.external_qkvbias):
                    qb = self.buf(pre + ".qkvbias", (3 * blk.attn.q_bias.numel(),), torch.float32)
                    _lib.pack_qkv_bias(blk.attn.q_bias.detach(), blk.attn.v_bias.detach(), qb)
        cast("e2d", m.encoder_to_decoder.weight)
        cast("head", m.decoder.head.weight)

    # ---- forward ------------------------------------------------------------------------------------------
    def _block_fwd(self, pre, blk, x, M, S, B, D, dp=None):
        """``dp``: None, or (s1, s2) f32 [B] DropPath scales (mask_b / keep_prob) of the attention / MLP branches."""
        H = blk.attn.num_heads
        bf, f32 = torch.bfloat16, torch.float32
        wc = self.wcache
        h1 = self.buf(pre + ".h1", (M, D), bf); mean1 = self.buf(pre + ".mean1", (M,), f32); rstd1 = self.buf(pre + ".rstd1", (M,), f32)
        _lib.layernorm_fwd(x, blk.norm1.weight, blk.norm1.bias, h1, mean1, rstd1, M, D, blk.norm1.eps)
        qkv = self.buf(pre + ".qkv", (M, 3 * D), bf)
        _lib.gemm_tn(h1, wc[pre + ".qkv"][0], _lib.EPI_BIAS_BF16, qkv, bias=self.buf(pre + ".qkvbias", (3 * D,), f32))
        o = self.buf(pre + ".o", (M, D), bf); lse = self.buf(pre + ".lse", (B, H, S), f32)
        # long sequences (streaming kernels): second bf16 word of O, read back only by backward's delta = dO . O
        olo = self.buf(pre + ".olo", (M, D), bf) if S > _lib.ATTN_SINGLE_PASS_MAX_S else None
        _lib.attn_fwd(qkv, B, S, H, blk.attn.scale, o, lse, out_lo=olo)
        xm = self.buf(pre + ".xm", (M, D), f32)
        _lib.gemm_tn(o, wc[pre + ".proj"][0], _lib.EPI_BIAS_RESID_F32, xm, bias=blk.attn.proj.bias, resid=x,
                     row_scale=dp[0] if dp else None, group_rows=S if dp else 0)
        h2 = self.buf(pre + ".h2", (M, D), bf); mean2 = self.buf(pre + ".mean2", (M,), f32); rstd2 = self.buf(pre + ".rstd2", (M,), f32)
        _lib.layernorm_fwd(xm, blk.norm2.weight, blk.norm2.bias, h2, mean2, rstd2, M, D, blk.norm2.eps)
        Dh = blk.mlp.fc1.weight.shape[0]
        u = self.buf(pre + ".u", (M, Dh), bf); a = self.buf(pre + ".a", (M, Dh), bf)
        _lib.gemm_tn(h2, wc[pre + ".fc1"][0], _lib.EPI_BIAS_GELU_BF16, u, out1=a, bias=blk.mlp.fc1.bias)
        xo = self.buf(pre + ".xo", (M, D), f32)
        _lib.gemm_tn(a, wc[pre + ".fc2"][0], _lib.EPI_BIAS_RESID_F32, xo, bias=blk.mlp.fc2.bias, resid=xm,
                     row_scale=dp[1] if dp else None, group_rows=S if dp else 0)
        return xo

    def forward(self, x, vis_idx, msk_idx):
        m = self.m
        self._ensure_device(x.device)
        self.prepare_weights()
        bf, f32 = torch.bfloat16, torch.float32
        B = x.shape[0]
        Nv, Nm = vis_idx.shape[1], msk_idx.shape[1]
        N = Nv + Nm
        D, Dd = m.encoder.embed_dim, m.decoder.embed_dim
        wc = self.wcache
        self.shape = (B, Nv, Nm)
        # patch embed on the visible tubes only + bias + pos  (modeling_pretrain.py:85-90)
        A_pe = self.buf("A_pe", (B * Nv, 1536), bf)
        _lib.gather_tubes(x, vis_idx, A_pe)
        xe = self.buf("x0", (B * Nv, D), f32)
        pe = m.encoder.patch_embed.proj
        _lib.gemm_tn(A_pe, wc["pe"][0], _lib.EPI_BIAS_POS_F32, xe, bias=pe.bias, pos=self.pos_enc, row_idx=vis_idx,
                     group_rows=Nv, out_group_rows=Nv)
        for i, blk in enumerate(m.encoder.blocks):
            xe = self._block_fwd(f"enc{i}", blk, xe, B * Nv, Nv, B, D)
        self.x_enc_out = xe
        hn = self.buf("enc.hn", (B * Nv, D), bf); mean = self.buf("enc.mean", (B * Nv,), f32); rstd = self.buf("enc.rstd", (B * Nv,), f32)
        _lib.layernorm_fwd(xe, m.encoder.norm.weight, m.encoder.norm.bias, hn, mean, rstd, B * Nv, D, m.encoder.norm.eps)
        # encoder_to_decoder + pos_vis written into the first Nv rows of each clip; mask tokens fill the rest
        xf = self.buf("xfull", (B * N, Dd), f32)
        _lib.gemm_tn(hn, wc["e2d"][0], _lib.EPI_BIAS_POS_F32, xf, pos=self.pos_dec, row_idx=vis_idx, group_rows=Nv,
                     out_group_rows=N)
        _lib.decoder_assemble_fwd(m.mask_token, self.pos_dec, msk_idx, B, Nv, Nm, Dd, xf)
        xd = xf
        for i, blk in enumerate(m.decoder.blocks):
            xd = self._block_fwd(f"dec{i}", blk, xd, B * N, N, B, Dd)
        self.x_dec_out = xd
        hd = self.buf("dec.hd", (B * Nm, Dd), bf); mean = self.buf("dec.mean", (B * Nm,), f32); rstd = self.buf("dec.rstd", (B * Nm,), f32)
        _lib.layernorm_fwd(xd, m.decoder.norm.weight, m.decoder.norm.bias, hd, mean, rstd, B * Nm, Dd, m.decoder.norm.eps,
                           group_rows=Nm, in_group_rows=N, in_row_offset=Nv)
        pred = self.buf("pred", (B * Nm, m.decoder.num_classes), bf)
        _lib.gemm_tn(hd, wc["head"][0], _lib.EPI_BIAS_BF16, pred, bias=m.decoder.head.bias)
        return pred

    # ---- backward -----------------------------------------------------------------------------------------
    def _block_bwd(self, pre, blk, g, x_in, dxA, dxA16, dxB, dxB16, M, S, B, D, dp_attn=None, dp_prev_mlp=None):
        """dxA/dxA16 hold d(loss)/d(block output) (f32 + bf16).  On return they hold d(loss)/d(block input).
        DropPath: dxA16 must arrive already scaled by THIS block's MLP-branch scale (the producer applies it);
        ``dp_attn`` = this block's attention-branch scale (applied to dxB16 by norm2's backward), ``dp_prev_mlp`` = the
        MLP-branch scale of the block that runs next in backward (applied to the dxA16 this call leaves behind)."""
        H = blk.attn.num_heads
        bf, f32 = torch.bfloat16, torch.float32
        wc = self.wcache
        Dh = blk.mlp.fc1.weight.shape[0]
        name = blk._mofo_name
        h1 = self.buf(pre + ".h1", (M, D), bf); mean1 = self.buf(pre + ".mean1", (M,), f32); rstd1 = self.buf(pre + ".rstd1", (M,), f32)
        qkv = self.buf(pre + ".qkv", (M, 3 * D), bf); o = self.buf(pre + ".o", (M, D), bf); lse = self.buf(pre + ".lse", (B, H, S), f32)
        xm = self.buf(pre + ".xm", (M, D), f32)
        h2 = self.buf(pre + ".h2", (M, D), bf); mean2 = self.buf(pre + ".mean2", (M,), f32); rstd2 = self.buf(pre + ".rstd2", (M,), f32)
        u = self.buf(pre + ".u", (M, Dh), bf); a = self.buf(pre + ".a", (M, Dh), bf)
        grouped = self.grouped_wgrad
        wg = []                      # (dY, X, dW, dbias, skip) of this block, launched together below when grouped

        def wgrad(reads, dY, X, dW, dbias=None, skip=(0, 0)):
            if grouped:
                wg.append((dY, X, dW, dbias, skip))
            else:
                self._wgrad(reads, dY, X, dW, dbias=dbias, skip=skip)
        # fc2
        wgrad((dxA16,), dxA16, a, g[name + ".mlp.fc2.weight"], dbias=g[name + ".mlp.fc2.bias"])
        du = self.buf("bwd.du", (M, Dh), bf)
        self._before_write(du)                                   # previous block's fc1 wgrad reads it
        _lib.gemm_tn(dxA16, wc[pre + ".fc2"][1], _lib.EPI_GELU_BWD_BF16, du, aux=u)
        # fc1
        wgrad((du,), du, h2, g[name + ".mlp.fc1.weight"], dbias=g[name + ".mlp.fc1.bias"])
        dh = self.buf("bwd.dh", (M, D), bf)
        _lib.gemm_tn(du, wc[pre + ".fc1"][1], _lib.EPI_PLAIN_BF16, dh)
        # norm2 (+ residual gradient)
        self._before_write(dxB16)                                # previous block's proj wgrad
        _lib.layernorm_bwd(dh, xm, blk.norm2.weight, mean2, rstd2, dxA, M, D, dxB, dxB16, g[name + ".norm2.weight"],
                           g[name + ".norm2.bias"], group_rows=S if dp_attn is not None else 0,
                           in_group_rows=S if dp_attn is not None else 0, bf16_row_scale=dp_attn)
        # proj
        wgrad((dxB16,), dxB16, o, g[name + ".attn.proj.weight"], dbias=g[name + ".attn.proj.bias"])
        do = self.buf("bwd.do", (M, D), bf)
        _lib.gemm_tn(dxB16, wc[pre + ".proj"][1], _lib.EPI_PLAIN_BF16, do)
        # attention
        dqkv = self.buf("bwd.dqkv", (M, 3 * D), bf); delta = self.buf("bwd.delta", (B, H, S), f32)
        self._before_write(dqkv)                                 # previous block's qkv wgrad
        olo = self.buf(pre + ".olo", (M, D), bf) if S > _lib.ATTN_SINGLE_PASS_MAX_S else None
        _lib.attn_bwd(qkv, o, do, lse, B, S, H, blk.attn.scale, dqkv, delta, out_lo=olo)
        # qkv
        # q_bias / v_bias gradients: column sums of dqkv[:, :D] and dqkv[:, 2D:], written through a [3D] window whose
        # first D floats are q_bias.grad and whose last D floats are v_bias.grad (see _make_arena: a D-float gap sits
        # between them in the arena so the window is contiguous); the K third is skipped.
        wgrad((dqkv,), dqkv, h1, g[name + ".attn.qkv.weight"], dbias=g[name + ".attn.q_bias"], skip=(D, 2 * D))
        _lib.gemm_tn(dqkv, wc[pre + ".qkv"][1], _lib.EPI_PLAIN_BF16, dh)
        if grouped:
            # every operand of the four weight gradients is still alive here: dxA16 (this block's output gradient) is only
            # overwritten by the norm1 backward below, du / dxB16 / dqkv by the NEXT block's backward
            _lib.gemm_wgrad_grouped(wg, M)
        # norm1 (+ residual gradient)
        self._before_write(dxA16)                                # this block's fc2 wgrad
        _lib.layernorm_bwd(dh, x_in, blk.norm1.weight, mean1, rstd1, dxB, M, D, dxA, dxA16, g[name + ".norm1.weight"],
                           g[name + ".norm1.bias"], group_rows=S if dp_prev_mlp is not None else 0,
                           in_group_rows=S if dp_prev_mlp is not None else 0, bf16_row_scale=dp_prev_mlp)

    def backward(self, dpred, g, stage_done=None):
        """dpred bf16 [B*Nm, 1536]; g: name -> fp32 tensor that the parameter gradient is ACCUMULATED into.
        ``stage_done(k)`` is called as soon as every gradient of arena stage k has been enqueued (see
        ``backward_order``), so the caller can start that slice's all-reduce while backward continues."""
        m = self.m
        bf, f32 = torch.bfloat16, torch.float32
        B, Nv, Nm = self.shape
        N = Nv + Nm
        D, Dd = m.encoder.embed_dim, m.decoder.embed_dim
        wc = self.wcache
        C = m.decoder.num_classes
        # head
        hd = self.buf("dec.hd", (B * Nm, Dd), bf)
        self._wgrad((), dpred, hd, g["decoder.head.weight"], dbias=g["decoder.head.bias"])
        dhd = self.buf("bwd.dhd", (B * Nm, Dd), bf)
        _lib.gemm_tn(dpred, wc["head"][1], _lib.EPI_PLAIN_BF16, dhd)
        # decoder.norm on the masked rows; visible rows receive zero gradient from the head
        dxA = self.buf("bwd.dec.dxA", (B * N, Dd), f32); dxA16 = self.buf("bwd.dec.dxA16", (B * N, Dd), bf)
        dxB = self.buf("bwd.dec.dxB", (B * N, Dd), f32); dxB16 = self.buf("bwd.dec.dxB16", (B * N, Dd), bf)
        _lib.zero_rows(dxA, dxA16, B, N, Nv, Dd)          # visible rows: no gradient from the head (masked rows: LN-bwd below)
        _lib.layernorm_bwd(dhd, self.x_dec_out, m.decoder.norm.weight, self.buf("dec.mean", (B * Nm,), f32),
                           self.buf("dec.rstd", (B * Nm,), f32), None, B * Nm, Dd, dxA, dxA16, g["decoder.norm.weight"],
                           g["decoder.norm.bias"], group_rows=Nm, in_group_rows=N, in_row_offset=Nv)
        nd = len(m.decoder.blocks)
        for i in range(nd - 1, -1, -1):
            x_in = self.buf("xfull", (B * N, Dd), f32) if i == 0 else self.buf(f"dec{i - 1}.xo", (B * N, Dd), f32)
            self._block_bwd(f"dec{i}", m.decoder.blocks[i], g, x_in, dxA, dxA16, dxB, dxB16, B * N, N, B, Dd)
        # decoder input assembly: mask_token gradient + gradient entering encoder_to_decoder
        dvis = self.buf("bwd.dvis", (B * Nv, Dd), bf)
        _lib.decoder_assemble_bwd(dxA, B, Nv, Nm, Dd, g["mask_token"], dvis)
        hn = self.buf("enc.hn", (B * Nv, D), bf)
        self._wgrad((), dvis, hn, g["encoder_to_decoder.weight"])
        dhn = self.buf("bwd.dhn", (B * Nv, D), bf)
        _lib.gemm_tn(dvis, wc["e2d"][1], _lib.EPI_PLAIN_BF16, dhn)
        exA = self.buf("bwd.enc.dxA", (B * Nv, D), f32); exA16 = self.buf("bwd.enc.dxA16", (B * Nv, D), bf)
        exB = self.buf("bwd.enc.dxB", (B * Nv, D), f32); exB16 = self.buf("bwd.enc.dxB16", (B * Nv, D), bf)
        _lib.layernorm_bwd(dhn, self.x_enc_out, m.encoder.norm.weight, self.buf("enc.mean", (B * Nv,), f32),
                           self.buf("enc.rstd", (B * Nv,), f32), None, B * Nv, D, exA, exA16, g["encoder.norm.weight"],
                           g["encoder.norm.bias"])
        if stage_done is not None:
            self._join_side()
            stage_done(0)
        ne = len(m.encoder.blocks)
        table = self.enc_stage_table()
        last_stage = table[0]
        for i in range(ne - 1, -1, -1):
            x_in = self.buf("x0", (B * Nv, D), f32) if i == 0 else self.buf(f"enc{i - 1}.xo", (B * Nv, D), f32)
            self._block_bwd(f"enc{i}", m.encoder.blocks[i], g, x_in, exA, exA16, exB, exB16, B * Nv, Nv, B, D)
            stage = table[i]
            if stage_done is not None and stage != last_stage and (i == 0 or table[i - 1] != stage):
                self._join_side()
                stage_done(stage)
        # patch embedding (no input gradient)
        A_pe = self.buf("A_pe", (B * Nv, 1536), bf)
        self._wgrad((), exA16, A_pe, g["encoder.patch_embed.proj.weight"], dbias=g["encoder.patch_embed.proj.bias"])
        self._join_side()
        if stage_done is not None:
            stage_done(last_stage)


    # ---- fused step: eager launch sequence and its CUDA-graph replay ------------------------------------
    def step_eager(self, videos, vis_idx, msk_idx, normalize_target, grad_scale, zero_grad, stage_done):
        pred = self.forward(videos, vis_idx, msk_idx)
        arena = self.grad_arena()
        if zero_grad:
            arena.zero_()
        B, Nv, Nm = self.shape
        lp = self.buf("loss_partials", (B * Nm,), torch.float32)
        loss = self.buf("loss", (1,), torch.float32)
        dpred = self.buf("dpred", (B * Nm, self.m.decoder.num_classes), torch.bfloat16)
        _lib.target_mse(videos, msk_idx, pred, lp, loss, dpred, normalize_target, grad_scale)
        self.backward(dpred, self.arena_views, stage_done)
        return loss

    def static_inputs(self, B, Nv, Nm, frames, size):
        """Device buffers the captured graphs read their inputs from (the engine's prefetcher copies H2D straight
        into them; any other caller's tensors are copied in by ``step_graphed``)."""
        key = (B, Nv, Nm, frames, size)
        t = self.static_in.get(key)
        if t is None:
            t = (torch.empty(B, 3, frames, size, size, dtype=torch.float32, device=self.device),
                 torch.empty(B, Nv, dtype=torch.int32, device=self.device),
                 torch.empty(B, Nm, dtype=torch.int32, device=self.device))
            self.static_in[key] = t
        return t

    def step_graphed(self, videos, vis_idx, msk_idx, normalize_target, grad_scale, stage_done):
        """Replays the whole fused step (weight casts, forward, arena zeroing, target + MSE, backward) as CUDA graphs,
        one per gradient-sync stage, so a step costs a handful of launches instead of ~360.  ``stage_done(k)`` is
        invoked between the segments exactly as in the eager sequence (NCCL all-reduces stay eager launches)."""
        B, _, frames, size, _ = videos.shape
        Nv, Nm = vis_idx.shape[1], msk_idx.shape[1]
        sv, si, sm = self.static_inputs(B, Nv, Nm, frames, size)
        if videos.data_ptr() != sv.data_ptr():
            sv.copy_(videos, non_blocking=True)
        if vis_idx.data_ptr() != si.data_ptr():
            si.copy_(vis_idx, non_blocking=True)
        if msk_idx.data_ptr() != sm.data_ptr():
            sm.copy_(msk_idx, non_blocking=True)
        arena = self.grad_arena()                              # (re)attaches the p.grad views
        # nobody listens to stage boundaries (one process, optimizer after backward): ONE graph instead of one per stage
        whole = stage_done is None
        key = (B, Nv, Nm, frames, size, bool(normalize_target), float(grad_scale), arena.data_ptr(), whole)
        graphs = self.graphs.get(key, "missing")
        if graphs == "missing":
            self.step_eager(sv, si, sm, normalize_target, grad_scale, True, None)   # warm-up: allocations, attributes
            graphs = self._capture(sv, si, sm, normalize_target, grad_scale, whole)
            self.graphs[key] = graphs
        if graphs is None:                                     # capture unavailable: same kernels, launched one by one
            return self.step_eager(sv, si, sm, normalize_target, grad_scale, True, stage_done)
        if self.external_weights:
            # the optimizer kernel maintains the bf16 operand copies, so the graphs hold no casts; if something else
            # changed the fp32 parameters since (load_state_dict, a broadcast, manual edits) re-cast eagerly first
            self.prepare_weights()
        self.shape = (B, Nv, Nm)
        for k, (g, n_kernels) in enumerate(graphs):
            g.replay()
            _lib.launch_count += n_kernels                     # kernels of ours inside the replayed segment
            if stage_done is not None:
                stage_done(k)
        return self.buf("loss", (1,), torch.float32)

    def _capture(self, sv, si, sm, normalize_target, grad_scale, whole=False):
        dev = self.device
        graphs = []
        cur = {"g": None, "n0": 0}
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        n_stages = 1 if whole else len(self.stage_end)

        def begin():
            cur["g"] = torch.cuda.CUDAGraph()
            cur["n0"] = _lib.launch_count
            cur["g"].capture_begin()       # private pool per segment: nothing is allocated while capturing

        def boundary(k):
            if whole and k + 1 < len(self.stage_end):        # single-graph capture: only the last boundary closes it
                return
            cur["g"].capture_end()
            graphs.append((cur["g"], _lib.launch_count - cur["n0"]))
            cur["g"] = None
            if not whole and k + 1 < n_stages:
                begin()

        self.force_cast = True
        try:
            with torch.cuda.stream(stream):
                begin()
                self.step_eager(sv, si, sm, normalize_target, grad_scale, True, boundary)
            torch.cuda.current_stream(dev).wait_stream(stream)
            assert cur["g"] is None and len(graphs) == n_stages
            return graphs
        except Exception as e:                                  # pragma: no cover - depends on driver / torch build
            import warnings
            warnings.warn(f"mofo_b200: CUDA graph capture failed ({e!r}); launching kernels individually")
            try:
                if cur["g"] is not None:
                    cur["g"].capture_end()
            except Exception:
                pass
            torch.cuda.synchronize(dev)
            return None
        finally:
            self.force_cast = False


class _ForwardFn(torch.autograd.Function):
    """pred = model(x, mask) with the manual kernel backward (drop-in autograd path)."""

    @staticmethod
    def forward(ctx, runner, x, vis_idx, msk_idx, *params):
        pred = runner.forward(x, vis_idx, msk_idx)
        ctx.runner = runner
        B, Nv, Nm = runner.shape
        return pred.view(B, Nm, -1).clone()

    @staticmethod
    def backward(ctx, dpred):
        r = ctx.runner
        B, Nv, Nm = r.shape
        if r.scratch_arena is None:
            r.scratch_arena = r._make_arena()
        arena, views = r.scratch_arena
        arena.zero_()
        dp = dpred.reshape(B * Nm, -1).to(torch.bfloat16).contiguous()
        r.backward(dp, views)
        grads = tuple(views[n] for n, _ in r.m.named_parameters())
        return (None, None, None, None) + grads


class PretrainVisionTransformer(nn.Module):
    """See module docstring; constructor signature of modeling_pretrain.py:166-190."""

    def __init__(self, img_size=224, patch_size=16, encoder_in_chans=3, encoder_num_classes=0, encoder_embed_dim=768,
                 encoder_depth=12, encoder_num_heads=12, decoder_num_classes=1536, decoder_embed_dim=512,
                 decoder_depth=8, decoder_num_heads=8, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0.,
                 use_learnable_pos_emb=False, tubelet_size=2, num_classes=0, in_chans=0):
        super().__init__()
        unsupported = dict(encoder_num_classes=(encoder_num_classes, 0), qk_scale=(qk_scale, None),
                           drop_rate=(drop_rate, 0.), attn_drop_rate=(attn_drop_rate, 0.),
                           drop_path_rate=(drop_path_rate, 0.), init_values=(init_values, 0.),
                           use_learnable_pos_emb=(use_learnable_pos_emb, False), tubelet_size=(tubelet_size, 2),
                           patch_size=(patch_size, 16), encoder_in_chans=(encoder_in_chans, 3))
        for k, (v, want) in unsupported.items():
            if v != want and not (v is None and want is None):
                raise NotImplementedError(f"mofo_b200 pretraining path supports {k}={want!r} only (got {v!r}); "
                                          "these are the values run_mae_pretraining_BB.py uses")
        if encoder_embed_dim // encoder_num_heads != 64 or decoder_embed_dim // decoder_num_heads != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64 (all registry models)")
        self.encoder = _Encoder(img_size, patch_size, encoder_in_chans, encoder_embed_dim, encoder_depth,
                                encoder_num_heads, mlp_ratio, qkv_bias, norm_layer, tubelet_size)
        self.decoder = _Decoder(patch_size, decoder_num_classes, decoder_embed_dim, decoder_depth, decoder_num_heads,
                                mlp_ratio, qkv_bias, norm_layer, tubelet_size)
        self.encoder_to_decoder = nn.Linear(encoder_embed_dim, decoder_embed_dim, bias=False)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        # plain tensor attributes, NOT buffers: absent from state_dict like the reference (modeling_pretrain.py:42,232)
        self.encoder_pos_embed = get_sinusoid_encoding_table(self.encoder.patch_embed.num_patches, encoder_embed_dim)
        self.pos_embed = get_sinusoid_encoding_table(self.encoder.patch_embed.num_patches, decoder_embed_dim)
        self.encoder.pos_embed = self.encoder_pos_embed
        _trunc_normal_(self.mask_token, std=.02)
        for i, blk in enumerate(self.encoder.blocks):
            blk._mofo_name = f"encoder.blocks.{i}"
        for i, blk in enumerate(self.decoder.blocks):
            blk._mofo_name = f"decoder.blocks.{i}"
        self._runner = _Runner(self)
        self.use_cuda_graph = True      # replay the fused step as CUDA graphs (pretrain_step only)
        self._n_msk = None
        self._bad_rows = None

    def get_num_layers(self):
        return len(self.encoder.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'mask_token'}       # modeling_pretrain.py:250-251

    # ---- helpers ------------------------------------------------------------------------------------------
    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError("mofo_b200.PretrainVisionTransformer runs on CUDA (sm_100a) only; there is no CPU path")
        if self.mask_token.device != x.device:
            raise RuntimeError("model parameters and input must be on the same CUDA device (call model.to(device))")

    def indices_from_mask(self, mask):
        """bool [B,N] -> ascending (vis_idx, msk_idx) int32 on device; the masked count per row is read once
        (first call) and verified on device afterwards (``check_mask_rows``)."""
        mask = mask.reshape(mask.shape[0], -1)
        if mask.dtype != torch.bool and mask.dtype != torch.uint8:
            mask = mask.to(torch.bool)
        mask = mask.contiguous()
        if self._n_msk is None or self._n_msk[0] != mask.shape[1]:
            self._n_msk = (mask.shape[1], int(mask[0].sum().item()))   # per token count N; rows are verified on device
        if self._bad_rows is None or self._bad_rows.device != mask.device:
            self._bad_rows = torch.zeros(1, dtype=torch.int32, device=mask.device)
        return _lib.mask_indices(mask, self._n_msk[1], self._bad_rows)

    def check_mask_rows(self, count=None):
        """``count``: a host copy of the device counter (the engine reads it with each step's loss); None reads it now."""
        if count is None:
            count = int(self._bad_rows.item()) if self._bad_rows is not None else 0
        if count != 0:
            raise RuntimeError("mask rows with unequal masked-token counts (the reference's x[~mask].reshape(B,-1,C) "
                               "raises here too, modeling_pretrain.py:90)")

    def zero_grad_arena(self):
        arena = self._runner.grad_arena() if self._runner.device is not None else None
        if arena is not None:
            arena.zero_()
        return arena

    # ---- reference API ------------------------------------------------------------------------------------
    def forward(self, x, mask):
        """x: CUDA f32 [B,3,16,224,224]; mask: bool [B,N] (True = masked) -> [B, N_mask, 1536] bf16."""
        self._require_cuda(x)
        x = x.float().contiguous()
        vis_idx, msk_idx = self.indices_from_mask(mask)
        r = self._runner
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _ForwardFn.apply(r, x, vis_idx, msk_idx, *self.parameters())
        with torch.no_grad():
            pred = r.forward(x, vis_idx, msk_idx)
            B, Nv, Nm = r.shape
            return pred.view(B, Nm, -1).clone()

    # ---- fused training step (used by mofo_b200.engine_for_pretraining) --------------------------------
    def pretrain_step(self, videos, mask=None, vis_idx=None, msk_idx=None, normalize_target=True, grad_scale=1.0,
                      zero_grad=True, stage_done=None):
        """One fused pass: forward, target + MSE (engine_for_pretraining.py:258-304) and backward.  Gradients are
        accumulated into the flat arena (``p.grad`` views).  Returns the loss as a 1-element CUDA tensor (no sync)."""
        self._require_cuda(videos)
        r = self._runner
        r._ensure_device(videos.device)
        if vis_idx is None:
            vis_idx, msk_idx = self.indices_from_mask(mask)
        if videos.dtype == torch.uint8:
            # raw uint8 clip (NCTHW): ToTorchFormatTensor + GroupNormalize on the GPU, written straight into the buffer
            # the step's graphs read (mofo_normalize_u8); the host ships 1 byte per sample instead of 4
            B, _, frames, size, _ = videos.shape
            sv = r.static_inputs(B, vis_idx.shape[1], msk_idx.shape[1], frames, size)[0]
            videos = _lib.normalize_u8(videos.contiguous(), sv)
        else:
            videos = videos.float().contiguous()
        with torch.no_grad():
            if self.use_cuda_graph and zero_grad:
                return r.step_graphed(videos, vis_idx, msk_idx, normalize_target, grad_scale, stage_done)
            return r.step_eager(videos, vis_idx, msk_idx, normalize_target, grad_scale, zero_grad, stage_done)


# ----------------------------------------------------------------------------------------------------------
# registry (modeling_pretrain.py:268-338)
# ----------------------------------------------------------------------------------------------------------
_REGISTRY = {}


def _register(fn):
    _REGISTRY[fn.__name__] = fn
    try:                                          # make timm.create_model(name, ...) work when timm is present
        from timm.models.registry import register_model
        register_model(fn)
    except Exception:
        pass
    return fn


def _finish(model, pretrained, kwargs):
    model.default_cfg = {'url': '', 'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
                         'interpolation': 'bicubic', 'mean': (0.5, 0.5, 0.5), 'std': (0.5, 0.5, 0.5)}
    if pretrained:
        checkpoint = torch.load(kwargs["init_ckpt"], map_location="cpu")
        model.load_state_dict(checkpoint["model"])
    return model


@_register
def pretrain_mae_small_patch16_224(pretrained=False, **kwargs):
    init_ckpt = {k: kwargs.pop(k) for k in ("init_ckpt",) if k in kwargs}
    model = PretrainVisionTransformer(img_size=224, patch_size=16, encoder_embed_dim=384, encoder_depth=12,
                                      encoder_num_heads=6, encoder_num_classes=0, decoder_num_classes=1536,
                                      decoder_embed_dim=192, decoder_num_heads=3, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    return _finish(model, pretrained, init_ckpt)


@_register
def pretrain_videomae_base_patch16_224(pretrained=False, **kwargs):
    init_ckpt = {k: kwargs.pop(k) for k in ("init_ckpt",) if k in kwargs}
    model = PretrainVisionTransformer(img_size=224, patch_size=16, encoder_embed_dim=768, encoder_depth=12,
                                      encoder_num_heads=12, encoder_num_classes=0, decoder_num_classes=1536,
                                      decoder_embed_dim=384, decoder_num_heads=6, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    return _finish(model, pretrained, init_ckpt)


@_register
def pretrain_videomae_large_patch16_224(pretrained=False, **kwargs):
    init_ckpt = {k: kwargs.pop(k) for k in ("init_ckpt",) if k in kwargs}
    model = PretrainVisionTransformer(img_size=224, patch_size=16, encoder_embed_dim=1024, encoder_depth=24,
                                      encoder_num_heads=16, encoder_num_classes=0, decoder_num_classes=1536,
                                      decoder_embed_dim=512, decoder_num_heads=8, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    return _finish(model, pretrained, init_ckpt)


def create_model(name, pretrained=False, **kwargs):
    """timm.create_model look-alike (drops None-valued kwargs like timm 0.4.12; run_mae_pretraining_BB.py:141-147)."""
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _REGISTRY[name](pretrained=pretrained, **kwargs)
