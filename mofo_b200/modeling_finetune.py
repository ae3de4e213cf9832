"""B200-native look-alike of the reference's ``modeling_finetune.VisionTransformer`` (BASELINE configs[4]).

Same public surface as /root/reference/modeling_finetune.py:305-409 for the plain classifier: constructor keywords,
``forward(x) -> logits [B, num_classes]``, ``forward_features``, ``get_num_layers``, ``no_weight_decay``,
``get_classifier`` / ``reset_classifier`` and the reference's ``state_dict`` names (``patch_embed.proj.*``,
``blocks.i.{norm1,attn.{q_bias,v_bias,qkv,proj},norm2,mlp.{fc1,fc2}}``, ``fc_norm.*``, ``head.*``), so a pretraining
checkpoint whose ``encoder.`` prefix was stripped (run_class_finetuning.py) loads unchanged; registry entry points
``vit_small_patch16_224``, ``vit_base_patch16_224``, ``vit_large_patch16_224``.

All 1568 tokens of a clip run through the encoder - the dense-attention stress of the hot path: tubelet gather + patch
embedding GEMM with the position table fused, 12 pre-LN blocks on the same kernels as the pretraining step
(``_Runner._block_fwd/_block_bwd``: tcgen05 GEMMs with fused epilogues, streaming attention at S = 1568, LayerNorm),
token mean pooling (``mofo_token_mean_fwd/bwd``), ``fc_norm`` and the classifier head (a padded-N GEMM).  ``forward`` is
one ``autograd.Function`` whose backward is the manual kernel backward, so the reference's ``engine_for_finetuning``
(criterion on the logits, ``loss_scaler(loss, optimizer, ...)``) drives it unchanged.  No CPU / eager fallback.

DropPath (the finetuning recipe's 0.1, modeling_finetune.py:20-31): the per-sample keep / scale factor of each residual
branch rides in the residual GEMM's epilogue (``row_scale``) and, in backward, on the bf16 gradient copy the branch's
GEMMs read (``bf16_row_scale`` of the LayerNorm backward / token-mean backward).
``VisionTransformer_BB_focused`` (:422-635, ``forward(x, BB)``) pools the same encoder by the motion box, with every fusing
method the reference runs ('org', 'weighted_mean', 'soft_attn', 'MCA' - see the class).
Not implemented (raises): dropout, ``init_values`` > 0, learnable position embedding, ``use_mean_pooling=False``.
"""
from __future__ import annotations

import os
from functools import partial

import torch
import torch.nn as nn

from . import _lib
from .modeling_pretrain import _Block, _PatchEmbed, _Runner, _trunc_normal_, get_sinusoid_encoding_table

__all__ = ["VisionTransformer", "VisionTransformer_BB_focused", "vit_small_patch16_224", "vit_base_patch16_224",
           "vit_base_patch16_224_BB_focused", "vit_large_patch16_224", "create_model"]


class _FtRunner(_Runner):
    """Kernel orchestration of the classifier: reuses the block forward / backward, buffers, weight cache and arena
    machinery of the pretraining runner; only the ends of the network differ."""

    def __init__(self, model):
        super().__init__(model)
        self.blk_group = int(os.environ.get("MOFO_ENC_GROUP", "4"))

    def _ensure_device(self, device):
        if self.device != device:
            self.device = device
            self.bufs.clear(); self.wcache.clear(); self.wversion = None
            self.arena = None; self.arena_views = None; self.scratch_arena = None
            self.pos = self.m.pos_embed[0].to(device).contiguous()

    def _named(self):
        skip = getattr(self.m, "_unused_prefixes", ())
        return [(n, p) for n, p in self.m.named_parameters() if not n.startswith(skip)] if skip else list(self.m.named_parameters())

    def backward_order(self):
        m = self.m
        names = dict(self._named())
        order, stage_of = [], {}

        def add(prefix, stage):
            for n in names:
                if (n == prefix or n.startswith(prefix + ".")) and n not in stage_of:
                    order.append(n); stage_of[n] = stage
        add("head", 0); add("fc_norm", 0); add("local_MCA", 0)
        nb = len(m.blocks)
        for i in range(nb - 1, -1, -1):
            add(f"blocks.{i}", (nb - 1 - i) // self.blk_group)
        last = (nb - 1) // self.blk_group
        add("patch_embed", last)
        assert len(order) == len(names)
        return order, stage_of, last + 1

    def prepare_weights(self):
        m = self.m
        version = tuple(p._version for p in m.parameters())
        if version == self.wversion:
            return
        self.wversion = version
        wc = self.wcache

        def cast(name, W, need_t=True, pad_rows=0):
            R = W.shape[0]
            C = W.numel() // R
            Rp = R + pad_rows
            if name not in wc:
                wb = torch.zeros(Rp, C, dtype=torch.bfloat16, device=self.device)
                wt = torch.zeros(C, Rp, dtype=torch.bfloat16, device=self.device) if need_t else None
                wc[name] = (wb, wt)
            wb, wt = wc[name]
            if pad_rows == 0:
                _lib.cast_weight(W.detach(), wb, wt)
            else:                   # classifier head: rows padded to a multiple of 8 with zeros (GEMM N / K granularity)
                tmp = torch.zeros(Rp, C, dtype=torch.float32, device=self.device)
                tmp[:R].copy_(W.detach())
                _lib.cast_weight(tmp, wb, wt)

        cast("pe", m.patch_embed.proj.weight, need_t=False)
        for i, blk in enumerate(m.blocks):
            pre = f"blk{i}"
            cast(pre + ".qkv", blk.attn.qkv.weight)
            cast(pre + ".proj", blk.attn.proj.weight)
            cast(pre + ".fc1", blk.mlp.fc1.weight)
            cast(pre + ".fc2", blk.mlp.fc2.weight)
            qb = self.buf(pre + ".qkvbias", (3 * blk.attn.q_bias.numel(),), torch.float32)
            _lib.pack_qkv_bias(blk.attn.q_bias.detach(), blk.attn.v_bias.detach(), qb)
        if getattr(m, "fusing_method", None) == "MCA":          # cross-attention block of the box-focused classifier
            mca = m.local_MCA[0]
            D = m.embed_dim
            # [q.weight; kv.weight] = the [3D, D] layout of a fused qkv projection (kv rows: K heads, then V heads, :140-143)
            cast("mca.qkv", torch.cat((mca.attn.q.weight.detach(), mca.attn.kv.weight.detach()), 0))
            cast("mca.proj", mca.attn.proj.weight)
            cast("mca.fc1", mca.mlp.fc1.weight)
            cast("mca.fc2", mca.mlp.fc2.weight)
            qb = self.buf("mca.qkvbias", (3 * D,), torch.float32)   # q_bias on Q; v_bias is added once, after P @ V (rows of P sum to 1)
            if not getattr(self, "_mca_bias_zeroed", False):
                qb.zero_(); self._mca_bias_zeroed = True
            qb[:D].copy_(mca.attn.q_bias.detach())
        C = m.num_classes
        self.c_pad = (C + 7) // 8 * 8
        cast("head", m.head.weight, pad_rows=self.c_pad - C)
        hb = self.buf("head.bias_pad", (self.c_pad,), torch.float32)
        hb.zero_()
        hb[:C].copy_(m.head.bias.detach())

    # ---- forward ------------------------------------------------------------------------------------------
    def forward(self, x, pool_weights=None, mca=None):
        """``pool_weights``: None (mean over tokens, :400) or f32 [B, N] per-token weights of the box-focused pooling.
        ``mca``: None, or (key_allowed u8 [B, N], w_plain f32 [B, N]) for fusing 'MCA': the cross-attention block runs on the
        encoder output, ``pool_weights`` (1 / n_in on the box tokens) pool ITS output and ``w_plain`` (1 / N on clips with no
        token in the box, else 0) pools the encoder output itself (:560-562)."""
        m = self.m
        self._ensure_device(x.device)
        self.prepare_weights()
        self.pool_weights = pool_weights
        self.mca_ctx = mca
        bf, f32 = torch.bfloat16, torch.float32
        B = x.shape[0]
        N = m.patch_embed.num_patches
        D = m.embed_dim
        wc = self.wcache
        self.shape = (B, N)
        idx = self.bufs.get(("idx_all", B, N))
        if idx is None:
            idx = torch.arange(N, dtype=torch.int32, device=self.device).repeat(B, 1).contiguous()
            self.bufs[("idx_all", B, N)] = idx
        # patch embedding of every tube + bias + position table (modeling_finetune.py:390-394)
        A_pe = self.buf("A_pe", (B * N, 1536), bf)
        _lib.gather_tubes(x, idx, A_pe)
        xe = self.buf("x0", (B * N, D), f32)
        pe = m.patch_embed.proj
        _lib.gemm_tn(A_pe, wc["pe"][0], _lib.EPI_BIAS_POS_F32, xe, bias=pe.bias, pos=self.pos, row_idx=idx, group_rows=N,
                     out_group_rows=N)
        # DropPath (modeling_finetune.py:20-31, 343-349): per block i and branch, sample b keeps its residual branch with
        # probability 1 - dpr[i] and the kept branches are scaled by 1 / keep_prob; one [2*depth, B] draw per forward
        self.dp = None
        if m.training and m.drop_path_rate > 0.0:
            keep = 1.0 - torch.tensor(m.dpr, dtype=f32, device=self.device).repeat_interleave(2)[:, None]     # [2*depth, 1]
            u = m._drop_path_uniform(2 * len(m.blocks), B, self.device)
            self.dp = ((keep + u).floor_() / keep).contiguous()        # timm drop_path: floor(keep + U[0,1)) / keep
        for i, blk in enumerate(m.blocks):
            dp = (self.dp[2 * i], self.dp[2 * i + 1]) if self.dp is not None else None
            xe = self._block_fwd(f"blk{i}", blk, xe, B * N, N, B, D, dp=dp)
        self.x_out = xe
        # norm = Identity, fc_norm(x.mean(1)), head  (:398-401, 405-406)
        pooled = self.buf("pooled", (B, D), f32)
        if mca is not None:
            allowed, w_plain = mca
            xo = self._mca_fwd(m.local_MCA[0], xe, allowed, B, N, D)
            plain = self.buf("pooled.plain", (B, D), f32)
            _lib.token_mean_fwd(xo, B, N, D, pooled, weights=pool_weights)
            _lib.token_mean_fwd(xe, B, N, D, plain, weights=w_plain)
            pooled.add_(plain)
        else:
            _lib.token_mean_fwd(xe, B, N, D, pooled, weights=pool_weights)
        hn = self.buf("fc.hn", (B, D), bf); mean = self.buf("fc.mean", (B,), f32); rstd = self.buf("fc.rstd", (B,), f32)
        _lib.layernorm_fwd(pooled, m.fc_norm.weight, m.fc_norm.bias, hn, mean, rstd, B, D, m.fc_norm.eps)
        logits = self.buf("logits", (B, self.c_pad), f32)
        zero = self.buf("logits.zero", (B, self.c_pad), f32)
        if not getattr(self, "_zeroed", False):
            zero.zero_(); self._zeroed = True
        _lib.gemm_tn(hn, wc["head"][0], _lib.EPI_BIAS_RESID_F32, logits, bias=self.buf("head.bias_pad", (self.c_pad,), f32),
                     resid=zero)
        return logits

    # ---- cross-attention block of fusing 'MCA' (modeling_finetune.py:162-191 over CrossAttention :100-160) ---------------
    # The reference slices each clip into its in-box tokens x (queries) and out-of-box tokens y (keys / values) and runs
    # x = x + attn(norm1(x), norm1(y)); x = x + mlp(norm2(x)) on the ragged pair, then averages x.  Here the block runs on ALL
    # N tokens of every clip with the keys restricted by a mask: rows of tokens outside the box are computed and then get
    # pooling weight 0, so their values never reach the logits and their gradient is exactly 0 - the same function of the
    # parameters, on dense [B, N] shapes.  Heads are D / 3 = 256 wide, so scores, P @ V and their gradients are tcgen05 GEMM
    # calls on per-clip, per-head views (K^T / V^T come from a second, operand-swapped GEMM; dK / dV from the MN-major
    # weight-gradient kernel) with the masked softmax kernels between them.
    def _mca_fwd(self, mca, x, allowed, B, N, D):
        bf, f32 = torch.bfloat16, torch.float32
        wc = self.wcache
        M, H = B * N, mca.attn.num_heads
        hd = D // H
        scale = mca.attn.scale
        h1 = self.buf("mca.h1", (M, D), bf); mean1 = self.buf("mca.mean1", (M,), f32); rstd1 = self.buf("mca.rstd1", (M,), f32)
        _lib.layernorm_fwd(x, mca.norm1.weight, mca.norm1.bias, h1, mean1, rstd1, M, D, mca.norm1.eps)
        qkv = self.buf("mca.qkv", (M, 3 * D), bf)
        _lib.gemm_tn(h1, wc["mca.qkv"][0], _lib.EPI_BIAS_BF16, qkv, bias=self.buf("mca.qkvbias", (3 * D,), f32))
        kvT = self.buf("mca.kvT", (B, 2 * D, N), bf)                 # K^T, V^T per clip: W_kv @ norm1(x_b)^T
        w_kv = wc["mca.qkv"][0][D:]
        for b in range(B):
            _lib.gemm_tn(w_kv, h1[b * N:(b + 1) * N], _lib.EPI_PLAIN_BF16, kvT[b])
        S = self.buf("mca.S", (B, H, N, N), f32)
        zero = self.buf("mca.zero", (N, N), f32)
        if not getattr(self, "_mca_zeroed", False):
            zero.zero_(); self._mca_zeroed = True
        for b in range(B):
            rows = qkv[b * N:(b + 1) * N]
            for h in range(H):
                _lib.gemm_tn(rows[:, h * hd:(h + 1) * hd], rows[:, D + h * hd:D + (h + 1) * hd], _lib.EPI_BIAS_RESID_F32, S[b, h],
                             resid=zero)
        P = self.buf("mca.P", (B, H, N, N), bf)
        _lib.masked_softmax_fwd(S, allowed, scale, P)
        o = self.buf("mca.o", (M, D), bf)
        v_bias = mca.attn.v_bias.detach()
        for b in range(B):
            for h in range(H):
                _lib.gemm_tn(P[b, h], kvT[b, D + h * hd:D + (h + 1) * hd], _lib.EPI_BIAS_BF16, o[b * N:(b + 1) * N, h * hd:(h + 1) * hd],
                             bias=v_bias[h * hd:(h + 1) * hd])
        xm = self.buf("mca.xm", (M, D), f32)
        _lib.gemm_tn(o, wc["mca.proj"][0], _lib.EPI_BIAS_RESID_F32, xm, bias=mca.attn.proj.bias, resid=x)
        h2 = self.buf("mca.h2", (M, D), bf); mean2 = self.buf("mca.mean2", (M,), f32); rstd2 = self.buf("mca.rstd2", (M,), f32)
        _lib.layernorm_fwd(xm, mca.norm2.weight, mca.norm2.bias, h2, mean2, rstd2, M, D, mca.norm2.eps)
        Dh = mca.mlp.fc1.weight.shape[0]
        u = self.buf("mca.u", (M, Dh), bf); a = self.buf("mca.a", (M, Dh), bf)
        _lib.gemm_tn(h2, wc["mca.fc1"][0], _lib.EPI_BIAS_GELU_BF16, u, out1=a, bias=mca.mlp.fc1.bias)
        xo = self.buf("mca.xo", (M, D), f32)
        _lib.gemm_tn(a, wc["mca.fc2"][0], _lib.EPI_BIAS_RESID_F32, xo, bias=mca.mlp.fc2.bias, resid=xm)
        return xo

    def _mca_bwd(self, mca, g, x_in, dxA, dxA16, dxB, dxB16, dpooled, B, N, D, dp_prev_mlp):
        """dxA / dxA16: gradient of the block's output on entry, of its input (the encoder output) on return."""
        bf, f32 = torch.bfloat16, torch.float32
        wc = self.wcache
        M, H = B * N, mca.attn.num_heads
        hd = D // H
        scale = mca.attn.scale
        Dh = mca.mlp.fc1.weight.shape[0]
        name = "local_MCA.0"
        h1 = self.buf("mca.h1", (M, D), bf); mean1 = self.buf("mca.mean1", (M,), f32); rstd1 = self.buf("mca.rstd1", (M,), f32)
        qkv = self.buf("mca.qkv", (M, 3 * D), bf); kvT = self.buf("mca.kvT", (B, 2 * D, N), bf)
        P = self.buf("mca.P", (B, H, N, N), bf); o = self.buf("mca.o", (M, D), bf); xm = self.buf("mca.xm", (M, D), f32)
        h2 = self.buf("mca.h2", (M, D), bf); mean2 = self.buf("mca.mean2", (M,), f32); rstd2 = self.buf("mca.rstd2", (M,), f32)
        u = self.buf("mca.u", (M, Dh), bf); a = self.buf("mca.a", (M, Dh), bf)
        # mlp, norm2
        _lib.gemm_wgrad(dxA16, a, g[name + ".mlp.fc2.weight"], dbias=g[name + ".mlp.fc2.bias"])
        du = self.buf("mca.du", (M, Dh), bf)
        _lib.gemm_tn(dxA16, wc["mca.fc2"][1], _lib.EPI_GELU_BWD_BF16, du, aux=u)
        _lib.gemm_wgrad(du, h2, g[name + ".mlp.fc1.weight"], dbias=g[name + ".mlp.fc1.bias"])
        dh = self.buf("mca.dh", (M, D), bf)
        _lib.gemm_tn(du, wc["mca.fc1"][1], _lib.EPI_PLAIN_BF16, dh)
        _lib.layernorm_bwd(dh, xm, mca.norm2.weight, mean2, rstd2, dxA, M, D, dxB, dxB16, g[name + ".norm2.weight"],
                           g[name + ".norm2.bias"])
        # proj
        _lib.gemm_wgrad(dxB16, o, g[name + ".attn.proj.weight"], dbias=g[name + ".attn.proj.bias"])
        do = self.buf("mca.do", (M, D), bf)
        _lib.gemm_tn(dxB16, wc["mca.proj"][1], _lib.EPI_PLAIN_BF16, do)
        _lib.colsum_bf16(do, M, D, g[name + ".attn.v_bias"])         # o = P @ V + v_bias
        # attention: dP = dO V^T, dS = scale * P (dP - sum P dP), dQ = dS K, dK = dS^T Q, dV = P^T dO
        dP = self.buf("mca.S", (B, H, N, N), f32)                    # the score buffer is free again
        zero = self.buf("mca.zero", (N, N), f32)
        for b in range(B):
            for h in range(H):
                _lib.gemm_tn(do[b * N:(b + 1) * N, h * hd:(h + 1) * hd], qkv[b * N:(b + 1) * N, 2 * D + h * hd:2 * D + (h + 1) * hd],
                             _lib.EPI_BIAS_RESID_F32, dP[b, h], resid=zero)
        dS = self.buf("mca.dS", (B, H, N, N), bf)
        _lib.masked_softmax_bwd(P, dP, scale, dS)
        dqkv = self.buf("mca.dqkv", (M, 3 * D), bf); dkv = self.buf("mca.dkv", (M, 2 * D), f32)
        dkv.zero_()
        for b in range(B):
            r0, r1 = b * N, (b + 1) * N
            for h in range(H):
                c0, c1 = h * hd, (h + 1) * hd
                _lib.gemm_tn(dS[b, h], kvT[b, c0:c1], _lib.EPI_PLAIN_BF16, dqkv[r0:r1, c0:c1])
                _lib.gemm_wgrad(dS[b, h], qkv[r0:r1, c0:c1], dkv[r0:r1, c0:c1])
                _lib.gemm_wgrad(P[b, h], do[r0:r1, c0:c1], dkv[r0:r1, D + c0:D + c1])
        _lib.cast_f32_bf16(dkv, dqkv[:, D:])
        # q / kv projections, norm1
        _lib.colsum_bf16(dqkv[:, :D], M, D, g[name + ".attn.q_bias"])
        _lib.gemm_wgrad(dqkv[:, :D], h1, g[name + ".attn.q.weight"])
        _lib.gemm_wgrad(dqkv[:, D:], h1, g[name + ".attn.kv.weight"])
        _lib.gemm_tn(dqkv, wc["mca.qkv"][1], _lib.EPI_PLAIN_BF16, dh)
        # clips without a token in the box pool the encoder output directly: + w_plain[b, n] * dpooled[b]
        dxB.view(B, N, D).addcmul_(self.mca_ctx[1][:, :, None], dpooled[:, None, :])
        _lib.layernorm_bwd(dh, x_in, mca.norm1.weight, mean1, rstd1, dxB, M, D, dxA, dxA16, g[name + ".norm1.weight"],
                           g[name + ".norm1.bias"], group_rows=N if dp_prev_mlp is not None else 0,
                           in_group_rows=N if dp_prev_mlp is not None else 0, bf16_row_scale=dp_prev_mlp)

    # ---- backward -----------------------------------------------------------------------------------------
    def backward(self, dlogits, g):
        """dlogits f32 [B, num_classes]; g: name -> fp32 tensor the parameter gradient is ACCUMULATED into."""
        m = self.m
        bf, f32 = torch.bfloat16, torch.float32
        B, N = self.shape
        D = m.embed_dim
        C = m.num_classes
        wc = self.wcache
        dl = self.buf("bwd.dl", (B, self.c_pad), bf)
        dl.zero_()
        dl[:, :C].copy_(dlogits)
        hn = self.buf("fc.hn", (B, D), bf)
        dWp = self.buf("bwd.dWhead", (self.c_pad, D), f32); dbp = self.buf("bwd.dbhead", (self.c_pad,), f32)
        dWp.zero_(); dbp.zero_()
        _lib.gemm_wgrad(dl, hn, dWp, dbias=dbp)
        g["head.weight"].add_(dWp[:C]); g["head.bias"].add_(dbp[:C])
        dhn = self.buf("bwd.dhn", (B, D), bf)
        _lib.gemm_tn(dl, wc["head"][1], _lib.EPI_PLAIN_BF16, dhn)
        dpooled = self.buf("bwd.dpooled", (B, D), f32)
        _lib.layernorm_bwd(dhn, self.buf("pooled", (B, D), f32), m.fc_norm.weight, self.buf("fc.mean", (B,), f32),
                           self.buf("fc.rstd", (B,), f32), None, B, D, dpooled, None, g["fc_norm.weight"], g["fc_norm.bias"])
        dxA = self.buf("bwd.dxA", (B * N, D), f32); dxA16 = self.buf("bwd.dxA16", (B * N, D), bf)
        dxB = self.buf("bwd.dxB", (B * N, D), f32); dxB16 = self.buf("bwd.dxB16", (B * N, D), bf)
        nb = len(m.blocks)
        dp = self.dp
        last_scale = dp[2 * nb - 1] if dp is not None else None       # MLP-branch DropPath scale of the last block
        if self.mca_ctx is not None:
            _lib.token_mean_bwd(dpooled, B, N, D, dxA, dxA16, weights=self.pool_weights)
            self._mca_bwd(m.local_MCA[0], g, self.x_out, dxA, dxA16, dxB, dxB16, dpooled, B, N, D, last_scale)
        else:
            _lib.token_mean_bwd(dpooled, B, N, D, dxA, dxA16, bf16_row_scale=last_scale, weights=self.pool_weights)
        for i in range(nb - 1, -1, -1):
            x_in = self.buf("x0", (B * N, D), f32) if i == 0 else self.buf(f"blk{i - 1}.xo", (B * N, D), f32)
            self._block_bwd(f"blk{i}", m.blocks[i], g, x_in, dxA, dxA16, dxB, dxB16, B * N, N, B, D,
                            dp_attn=dp[2 * i] if dp is not None else None,
                            dp_prev_mlp=dp[2 * i - 1] if (dp is not None and i > 0) else None)
        A_pe = self.buf("A_pe", (B * N, 1536), bf)
        self._wgrad((), dxA16, A_pe, g["patch_embed.proj.weight"], dbias=g["patch_embed.proj.bias"])
        self._join_side()


class _FtForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, x, pool_weights, mca, *params):
        logits = runner.forward(x, pool_weights, mca)
        ctx.runner = runner
        return logits[:, :runner.m.num_classes].clone()

    @staticmethod
    def backward(ctx, dlogits):
        r = ctx.runner
        if r.scratch_arena is None:
            r.scratch_arena = r._make_arena()
        arena, views = r.scratch_arena
        arena.zero_()
        r.backward(dlogits.float().contiguous(), views)
        return (None, None, None, None) + tuple(views[n] for n, _ in r._named())


class VisionTransformer(nn.Module):
    """Constructor signature of modeling_finetune.py:308-328."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, init_values=0., use_learnable_pos_emb=False, init_scale=0., all_frames=16,
                 tubelet_size=2, use_mean_pooling=True):
        super().__init__()
        unsupported = dict(patch_size=(patch_size, 16), in_chans=(in_chans, 3), qk_scale=(qk_scale, None), drop_rate=(drop_rate, 0.),
                           attn_drop_rate=(attn_drop_rate, 0.), init_values=(init_values, 0.),
                           use_learnable_pos_emb=(use_learnable_pos_emb, False), tubelet_size=(tubelet_size, 2),
                           use_mean_pooling=(use_mean_pooling, True))
        for k, (v, want) in unsupported.items():
            if v != want and not (v is None and want is None):
                raise NotImplementedError(f"mofo_b200 finetuning classifier supports {k}={want!r} only (got {v!r})")
        if embed_dim // num_heads != 64 or num_classes <= 0:
            raise NotImplementedError("attention kernels are specialised for head_dim 64; num_classes must be positive")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.drop_path_rate = float(drop_path_rate)
        self.dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]      # stochastic depth decay rule (:343)
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim, num_frames=all_frames, tubelet_size=tubelet_size)
        self.pos_embed = get_sinusoid_encoding_table(self.patch_embed.num_patches, embed_dim)   # attribute, not a buffer (:338-340)
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = nn.Identity()
        self.fc_norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        _trunc_normal_(self.head.weight, std=.02)
        self.apply(self._init_weights)
        self.head.weight.data.mul_(init_scale)
        self.head.bias.data.mul_(init_scale)
        for i, blk in enumerate(self.blocks):
            blk._mofo_name = f"blocks.{i}"
        self._runner = _FtRunner(self)

    def _init_weights(self, m):                       # modeling_finetune.py:365-372
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    def _drop_path_uniform(self, n, B, device):
        """U[0,1) draws behind the DropPath masks, [n, B] (tests substitute a fixed tensor to compare with the reference)."""
        return torch.rand(n, B, dtype=torch.float32, device=device)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes).to(self.head.weight.device)
        self._runner.wcache.pop("head", None); self._runner.wversion = None
        self._runner.arena = None; self._runner.arena_views = None; self._runner.scratch_arena = None

    def forward(self, x):
        """x: CUDA f32 [B,3,all_frames,224,224] -> logits f32 [B, num_classes] (autograd flows to every parameter)."""
        if not x.is_cuda:
            raise RuntimeError("mofo_b200.VisionTransformer runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.float().contiguous()
        return self._run(x, None)

    def _run(self, x, pool_weights, mca=None):
        r = self._runner
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _FtForwardFn.apply(r, x, pool_weights, mca, *[p for _, p in r._named()])
        with torch.no_grad():
            return r.forward(x, pool_weights, mca)[:, :self.num_classes].clone()


# ---- parameter holders of the box-focused classifier's fusing modules (names / shapes / init of modeling_finetune.py:100-191,
# 264-303), so that state_dict round-trips with the reference.  local_MCA[0] is run by _FtRunner._mca_fwd / _mca_bwd;
# SoftAttention reduces to a plain mean (see VisionTransformer_BB_focused) and global_MCA is never called by the reference
class _SoftAttention(nn.Module):
    def __init__(self, feature_dim, step_dim, bias=True):
        super().__init__()
        weight = torch.zeros(feature_dim, 1)
        nn.init.kaiming_uniform_(weight)
        self.weight = nn.Parameter(weight)
        if bias:
            self.b = nn.Parameter(torch.zeros(step_dim))


class _CrossAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=False)
        self.kv = nn.Linear(dim, dim * 2, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)


class _MCA(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        from .modeling_pretrain import _Mlp
        self.norm1 = norm_layer(dim)
        self.attn = _CrossAttention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class VisionTransformer_BB_focused(VisionTransformer):
    """Box-focused classifier (modeling_finetune.py:422-635), ``forward(x, BB)`` with ``BB`` int [B, frames, 4].

    Built: the token-in-box predicate (closed form of the reference's all-ones ``patch_yab`` Conv3d over a painted clip,
    ``mofo_box_tokens``, integer-exact) and the fusing methods ``'org'``, ``'weighted_mean'`` (the constructor default) and
    ``'soft_attn'`` as a per-token weighted pooling inside the same kernel pipeline.  (``'soft_attn'`` as the reference WRITES
    it reduces to mean_in + mean_out: SoftAttention's [n,c] * [n,1,1] broadcast followed by sum(1).mean(0) leaves
    (sum_i a_i) * mean(x) with sum_i a_i = 1, so its own parameters get a mathematically zero gradient - none here.)
    ``'MCA'`` (the finetuning script's default, :575-583): ``local_MCA[0]``, a pre-LN block whose attention takes its queries
    from the tokens in the box and its keys / values from the tokens outside it (3 heads of embed_dim / 3), runs on the
    encoder output as a key-masked dense block (``_FtRunner._mca_fwd``) and its box tokens are averaged; ``global_MCA`` is
    commented out in the reference and stays unused.  Requires qkv_bias=True (what every registered model passes)."""

    _FUSING = {"org": 0, "weighted_mean": 1, "soft_attn": 2, "MCA": 0}     # -> mofo_box_tokens mode

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, init_values=0., use_learnable_pos_emb=False, init_scale=0., all_frames=16,
                 tubelet_size=2, use_mean_pooling=True, fusing_method='weighted_mean'):
        if fusing_method not in self._FUSING:
            raise NotImplementedError(f"mofo_b200 box-focused classifier implements fusing_method {sorted(self._FUSING)} (got {fusing_method!r})")
        super().__init__(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads, mlp_ratio, qkv_bias, qk_scale,
                         drop_rate, attn_drop_rate, drop_path_rate, norm_layer, init_values, use_learnable_pos_emb, 1.0,
                         all_frames, tubelet_size, use_mean_pooling)
        if fusing_method == "MCA" and (not qkv_bias or embed_dim % 3 != 0 or (embed_dim // 3) % 64 != 0):
            raise NotImplementedError("fusing_method 'MCA' needs qkv_bias=True and embed_dim / 3 a multiple of 64")
        self.fusing_method = fusing_method
        # parameters the forward never touches (no gradient in the reference either)
        self._unused_prefixes = ("soft_att_local", "soft_att_global", "global_MCA", "patch_yab") + \
            (() if fusing_method == "MCA" else ("local_MCA",))
        self.in_chans = in_chans
        self.soft_att_local = _SoftAttention(embed_dim, 1)
        self.soft_att_global = _SoftAttention(embed_dim, 1)
        self.local_MCA = nn.ModuleList([_MCA(embed_dim, 3, mlp_ratio, qkv_bias, norm_layer)])
        self.global_MCA = nn.ModuleList([_MCA(embed_dim, 3, mlp_ratio, qkv_bias, norm_layer)])
        for mod in (self.local_MCA, self.global_MCA):
            mod.apply(self._init_weights)
        self.head.weight.data.mul_(init_scale)          # the base constructor ran with init_scale 1 (:504-505 scale AFTER all inits)
        self.head.bias.data.mul_(init_scale)
        self.patch_yab = nn.Conv3d(in_chans, embed_dim, kernel_size=(tubelet_size, patch_size, patch_size),
                                   stride=(tubelet_size, patch_size, patch_size))
        self.patch_yab.weight.data.fill_(1)
        self.patch_yab.bias.data.fill_(0)
        self._runner = _FtRunner(self)

    def tokens_in_box(self, BB, frames, size):
        """bool [B, N]: what the reference's ``x_patch_yabide`` holds (:589-630)."""
        bb = torch.as_tensor(BB).to(device=self.head.weight.device, dtype=torch.int64).contiguous()
        return _lib.box_tokens(bb, frames, size, 0, want_weights=False)[0].bool()

    def forward(self, x, BB):
        if not x.is_cuda:
            raise RuntimeError("mofo_b200.VisionTransformer_BB_focused runs on CUDA (sm_100a) only; there is no CPU path")
        x = x.float().contiguous()
        bb = torch.as_tensor(BB).to(device=x.device, dtype=torch.int64).contiguous()
        if self.fusing_method != "MCA":
            _, weights = _lib.box_tokens(bb, x.shape[2], x.shape[3], self._FUSING[self.fusing_method])
            return self._run(x, weights)
        inbox, _ = _lib.box_tokens(bb, x.shape[2], x.shape[3], 0, want_weights=False)
        inb = inbox.bool()
        N = inb.shape[1]
        n_in = inb.sum(1, keepdim=True)
        w_in = (inb.float() / n_in.clamp(min=1)).contiguous()                 # in_bbx.mean(0) of the block's output (:583)
        w_plain = ((n_in == 0).float() / N).expand(-1, N).contiguous()         # no token in the box: x[i].mean(0) (:560-562)
        key_allowed = (~inb | (n_in == N)).to(torch.uint8).contiguous()        # keys = tokens outside the box; none -> y = x (:131-133)
        return self._run(x, w_in, (key_allowed, w_plain))


_REGISTRY = {}


def _register(fn):
    _REGISTRY[fn.__name__] = fn
    try:
        from timm.models.registry import register_model
        register_model(fn)
    except Exception:
        pass
    return fn


@_register
def vit_small_patch16_224(pretrained=False, **kwargs):
    return VisionTransformer(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


@_register
def vit_base_patch16_224(pretrained=False, **kwargs):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


@_register
def vit_base_patch16_224_BB_focused(pretrained=False, **kwargs):
    return VisionTransformer_BB_focused(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                        norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


@_register
def vit_large_patch16_224(pretrained=False, **kwargs):
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def create_model(name, pretrained=False, **kwargs):
    """timm.create_model look-alike (drops None-valued kwargs like timm 0.4.12; run_class_finetuning.py)."""
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _REGISTRY[name](pretrained=pretrained, **kwargs)
