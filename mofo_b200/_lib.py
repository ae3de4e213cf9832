"""ctypes binding of libmofo_sm100.so (the C ABI declared in include/mofo_b200.h).

The library is the only compute path: there is no CPU or PyTorch fallback.  Loading fails loudly if
the shared object is missing (run ``python -m mofo_b200.build``), and every call raises ``MofoError``
with the library's message on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOFO_B200_LIB", os.path.join(HERE, "libmofo_sm100.so"))

EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_PLAIN_BF16, EPI_GELU_BWD_BF16, EPI_BIAS_POS_F32 = range(6)

# symbol -> argtypes ; must list every function declared in include/mofo_b200.h
_P, _I, _F, _D, _L = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_int64
SIGNATURES = {
    "mofo_version": ([], C.c_int),
    "mofo_last_error": ([], C.c_char_p),
    "mofo_sm_count": ([], C.c_int),
    "mofo_tube_mask_bb": ([_P, _P, _I, _I, _I, _I, _I, _I, _D, _P, _P, _P, _P, _P], C.c_int),
    "mofo_tube_mask_plain": ([_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P], C.c_int),
    "mofo_mask_indices": ([_P, _I, _I, _I, _P, _P, _P, _P], C.c_int),
    "mofo_gather_tubes": ([_P, _P, _I, _I, _I, _I, _P, _P], C.c_int),
    "mofo_gemm_tn": ([_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _I, _I, _P, _I, _P, _I, _P, _P], C.c_int),
    "mofo_gemm_wgrad": ([_P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P], C.c_int),
    "mofo_gemm_wgrad_grouped": ([_I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "mofo_attn_fwd": ([_P, _I, _I, _I, _F, _P, _P, _P, _P], C.c_int),
    "mofo_attn_bwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P], C.c_int),
    "mofo_layernorm_fwd": ([_P, _P, _P, _I, _I, _F, _I, _I, _I, _P, _P, _P, _P], C.c_int),
    "mofo_layernorm_bwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], C.c_int),
    "mofo_decoder_assemble_fwd": ([_P, _P, _P, _I, _I, _I, _I, _P, _P], C.c_int),
    "mofo_decoder_assemble_bwd": ([_P, _I, _I, _I, _I, _P, _P, _P], C.c_int),
    "mofo_zero_rows": ([_P, _P, _I, _I, _I, _I, _P], C.c_int),
    "mofo_token_mean_fwd": ([_P, _P, _I, _I, _I, _P, _P], C.c_int),
    "mofo_token_mean_bwd": ([_P, _P, _I, _I, _I, _P, _P, _P, _P], C.c_int),
    "mofo_box_tokens": ([_P, _I, _I, _I, _I, _P, _P, _P], C.c_int),
    "mofo_masked_softmax_fwd": ([_P, _P, _I, _I, _I, _F, _P, _P], C.c_int),
    "mofo_masked_softmax_bwd": ([_P, _P, C.c_int64, _I, _F, _P, _P], C.c_int),
    "mofo_cast_f32_bf16": ([_P, _I, _I, _I, _P, _I, _P], C.c_int),
    "mofo_target_mse": ([_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P], C.c_int),
    "mofo_cast_weight": ([_P, _I, _I, _P, _P, _P], C.c_int),
    "mofo_pack_qkv_bias": ([_P, _P, _I, _P, _P], C.c_int),
    "mofo_colsum_bf16": ([_P, _I, _I, _I, _P, _P], C.c_int),
    "mofo_sq_norm_f32": ([_P, _L, _P, _P], C.c_int),
    "mofo_normalize_u8": ([_P, _I, _I, _I, _P, _P], C.c_int),
    "mofo_clip_preprocess": ([_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P], C.c_int),
    "mofo_adamw_step": ([_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P], C.c_int),
    "mofo_motion_map": ([_P, _I, _I, _I, _I, _I, _I, _P, _I, _P], C.c_int),
    "mofo_motion_box_filter": ([_P, _I, _I, _I, _P, _I, _P, _I, _D, _D, _D, _P, _P, _P, _P, _P], C.c_int),
}


class MofoError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MofoError(f"{LIB_PATH} not found: build it with `python -m mofo_b200.build` "
                        "(there is no CPU/PyTorch fallback for the MOFO hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


# kernels launched per successful C-ABI call (bench.py reports the count it observed as "gpu_launches")
_LAUNCHES = {"mofo_target_mse": 2, "mofo_attn_bwd": 3, "mofo_motion_box_filter": 8}
launch_count = 0


class KernelTimer:
    """Optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline leg):
    ``with _lib.timing("mofo_gemm_tn") as t: ...`` then ``t.summary()`` -> (launches, total_ms, total_flops)."""

    def __init__(self, name):
        self.name, self.events, self.flops = name, [], 0.0

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.events)
        return len(self.events), ms, self.flops


_timer = None


def timing(name):
    return KernelTimer(name)


def _check(rc: int, what: str):
    global launch_count
    if rc != 0:
        msg = load().mofo_last_error()
        raise MofoError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
    launch_count += _LAUNCHES.get(what, 1)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "mofo_b200 kernels take CUDA tensors only"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
# thin tensor-level wrappers (shape bookkeeping only; all arithmetic happens in the library)
# ------------------------------------------------------------------------------------------------
def tube_mask_bb(bb_first, rng_words, grid, n_mask_per_frame, ratio_bb):
    """bb_first f64 [B,4], rng_words u32-as-int32/uint32 [B,W] (CUDA) -> mask u8 [B,N], vis_idx, msk_idx, words_used."""
    T, H, Wd = grid
    B, W = rng_words.shape
    npf = H * Wd
    dev = rng_words.device
    mask = torch.empty(B, T * npf, dtype=torch.uint8, device=dev)
    vis = torch.empty(B, T * (npf - n_mask_per_frame), dtype=torch.int32, device=dev)
    msk = torch.empty(B, T * n_mask_per_frame, dtype=torch.int32, device=dev)
    used = torch.empty(B, dtype=torch.int32, device=dev)
    if bb_first is None:
        _check(load().mofo_tube_mask_plain(_ptr(rng_words), B, W, T, H, Wd, n_mask_per_frame, _ptr(mask), _ptr(vis),
                                           _ptr(msk), _ptr(used), _stream()), "mofo_tube_mask_plain")
    else:
        assert bb_first.dtype == torch.float64 and bb_first.shape == (B, 4) and bb_first.is_contiguous()
        _check(load().mofo_tube_mask_bb(_ptr(bb_first), _ptr(rng_words), B, W, T, H, Wd, n_mask_per_frame,
                                        float(ratio_bb), _ptr(mask), _ptr(vis), _ptr(msk), _ptr(used), _stream()),
               "mofo_tube_mask_bb")
    return mask, vis, msk, used


def mask_indices(mask, n_msk, bad_rows):
    """mask bool/uint8 [B,N] (CUDA) -> ascending (vis_idx [B,N-n_msk], msk_idx [B,n_msk]) int32."""
    B, N = mask.shape
    assert mask.element_size() == 1 and mask.is_contiguous()
    vis = torch.empty(B, N - n_msk, dtype=torch.int32, device=mask.device)
    msk = torch.empty(B, n_msk, dtype=torch.int32, device=mask.device)
    _check(load().mofo_mask_indices(_ptr(mask), B, N, n_msk, _ptr(vis), _ptr(msk), _ptr(bad_rows), _stream()),
           "mofo_mask_indices")
    return vis, msk


def gather_tubes(video, idx, out=None):
    B, Cc, frames, size, _ = video.shape
    assert Cc == 3 and video.dtype == torch.float32 and video.is_contiguous() and idx.dtype == torch.int32
    n = idx.shape[1]
    if out is None:
        out = torch.empty(B * n, 1536, dtype=torch.bfloat16, device=video.device)
    _check(load().mofo_gather_tubes(_ptr(video), _ptr(idx), B, n, frames, size, _ptr(out), _stream()), "mofo_gather_tubes")
    return out


def gemm_tn(A, Bm, epilogue, out0, out1=None, bias=None, resid=None, aux=None, pos=None, row_idx=None,
            group_rows=0, out_group_rows=0, M=None, row_scale=None):
    """out0 = epi(A[M,K] @ Bm[N,K]^T).  A/Bm bf16 row-major (last dim contiguous)."""
    M = A.shape[0] if M is None else M
    K = A.shape[1]
    N = Bm.shape[0]
    assert A.dtype == torch.bfloat16 and Bm.dtype == torch.bfloat16 and Bm.shape[1] == K
    assert A.stride(1) == 1 and Bm.stride(1) == 1 and out0.stride(-1) == 1
    t = _timer if (_timer is not None and _timer.name == "mofo_gemm_tn") else None
    if t is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    _check(load().mofo_gemm_tn(_ptr(A), A.stride(0), _ptr(Bm), Bm.stride(0), M, N, K, epilogue, _ptr(bias), _ptr(resid),
                               resid.stride(0) if resid is not None else 0, _ptr(aux),
                               aux.stride(0) if aux is not None else 0, _ptr(pos), _ptr(row_idx), group_rows,
                               out_group_rows, _ptr(out0), out0.stride(0), _ptr(out1),
                               out1.stride(0) if out1 is not None else 0, _ptr(row_scale), _stream()), "mofo_gemm_tn")
    if t is not None:
        e1.record()
        t.events.append((e0, e1)); t.flops += 2.0 * M * N * K
    return out0


def gemm_wgrad(dY, X, dW, M=None, dbias=None, skip=(0, 0)):
    """dW[N,K] += dY[M,N]^T @ X[M,K]  (dW f32, accumulated); optionally dbias[N] += colsum(dY) outside skip=[lo,hi)."""
    M = dY.shape[0] if M is None else M
    N, K = dY.shape[1], X.shape[1]
    assert dY.dtype == torch.bfloat16 and X.dtype == torch.bfloat16 and dW.dtype == torch.float32
    assert dY.stride(1) == 1 and X.stride(1) == 1 and dW.stride(-1) == 1
    ldw = K
    if dW.dim() == 2 and dW.shape == (N, K):           # may be a column slice of a wider f32 matrix
        ldw = dW.stride(0)
    else:
        assert dW.numel() == N * K and dW.is_contiguous()
    _check(load().mofo_gemm_wgrad(_ptr(dY), dY.stride(0), _ptr(X), X.stride(0), M, N, K, _ptr(dW), ldw, _ptr(dbias),
                                  skip[0], skip[1], _stream()), "mofo_gemm_wgrad")
    return dW


ATTN_SINGLE_PASS_MAX_S = 192      # sequences up to this length run the single-pass kernels (no out_lo, no delta pass)


def gemm_wgrad_grouped(problems, M):
    """problems: list (<= 4) of (dY, X, dW, dbias | None, (skip_lo, skip_hi)); every problem reduces over the same M rows.
    One launch; falls back to individual mofo_gemm_wgrad calls when a shape does not fit the grouped kernel."""
    n = len(problems)
    if n == 0:
        return
    ks = [X.shape[1] for _, X, _, _, _ in problems]
    if n > 4 or not (all(k % 192 == 0 for k in ks) or all(k % 256 == 0 for k in ks)):
        for dY, X, dW, dbias, skip in problems:
            gemm_wgrad(dY, X, dW, M=M, dbias=dbias, skip=skip)
        return
    P_, I_ = C.c_void_p * n, C.c_int * n
    for dY, X, dW, _, _ in problems:
        assert dY.dtype == torch.bfloat16 and X.dtype == torch.bfloat16 and dW.dtype == torch.float32
        assert dY.stride(1) == 1 and X.stride(1) == 1 and dW.stride(-1) == 1 and dW.numel() == dY.shape[1] * X.shape[1]
    _check(load().mofo_gemm_wgrad_grouped(
        n, P_(*[_ptr(p[0]) for p in problems]), I_(*[p[0].stride(0) for p in problems]),
        P_(*[_ptr(p[1]) for p in problems]), I_(*[p[1].stride(0) for p in problems]), M,
        I_(*[p[0].shape[1] for p in problems]), I_(*[p[1].shape[1] for p in problems]),
        P_(*[_ptr(p[2]) for p in problems]), I_(*[p[1].shape[1] for p in problems]),
        P_(*[_ptr(p[3]) for p in problems]), I_(*[p[4][0] for p in problems]), I_(*[p[4][1] for p in problems]), _stream()),
        "mofo_gemm_wgrad_grouped")


def attn_fwd(qkv, B, S, H, scale, out, lse, out_lo=None):
    _check(load().mofo_attn_fwd(_ptr(qkv), B, S, H, float(scale), _ptr(out), _ptr(out_lo), _ptr(lse), _stream()), "mofo_attn_fwd")
    return out, lse


def attn_bwd(qkv, out, dout, lse, B, S, H, scale, dqkv, delta, out_lo=None):
    global launch_count
    _check(load().mofo_attn_bwd(_ptr(qkv), _ptr(out), _ptr(out_lo), _ptr(dout), _ptr(lse), B, S, H, float(scale), _ptr(dqkv),
                                _ptr(delta), _stream()), "mofo_attn_bwd")
    if S <= ATTN_SINGLE_PASS_MAX_S and os.environ.get("MOFO_ATTN_SMALL", "1") != "0":
        launch_count -= 2                      # one kernel instead of delta + dQ + dK/dV
    return dqkv


def layernorm_fwd(x, gamma, beta, y, mean, rstd, M, D, eps=1e-6, group_rows=0, in_group_rows=0, in_row_offset=0):
    g = group_rows if group_rows > 0 else M
    ig = in_group_rows if in_group_rows > 0 else M
    _check(load().mofo_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), M, D, eps, g, ig, in_row_offset, _ptr(y),
                                     _ptr(mean), _ptr(rstd), _stream()), "mofo_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, gamma, mean, rstd, dres, M, D, dx_f32, dx_bf16, dgamma, dbeta, group_rows=0,
                  in_group_rows=0, in_row_offset=0, bf16_row_scale=None):
    g = group_rows if group_rows > 0 else M
    ig = in_group_rows if in_group_rows > 0 else M
    _check(load().mofo_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), M, D, g, ig,
                                     in_row_offset, _ptr(dx_f32), _ptr(dx_bf16), _ptr(dgamma), _ptr(dbeta), _ptr(bf16_row_scale),
                                     _stream()),
           "mofo_layernorm_bwd")


def decoder_assemble_fwd(mask_token, pos, msk_idx, B, n_vis, n_msk, Dd, x_full):
    _check(load().mofo_decoder_assemble_fwd(_ptr(mask_token), _ptr(pos), _ptr(msk_idx), B, n_vis, n_msk, Dd,
                                            _ptr(x_full), _stream()), "mofo_decoder_assemble_fwd")


def decoder_assemble_bwd(dx_full, B, n_vis, n_msk, Dd, dmask_token, dvis):
    _check(load().mofo_decoder_assemble_bwd(_ptr(dx_full), B, n_vis, n_msk, Dd, _ptr(dmask_token), _ptr(dvis),
                                            _stream()), "mofo_decoder_assemble_bwd")


def zero_rows(x_f32, x_bf16, groups, group_rows, n_zero, D):
    _check(load().mofo_zero_rows(_ptr(x_f32), _ptr(x_bf16), groups, group_rows, n_zero, D, _stream()), "mofo_zero_rows")


def token_mean_fwd(x, B, N, D, pooled, weights=None):
    """pooled f32 [B, D] = sum_n weights[b, n] * x[b*N + n] (weights None: the plain mean over tokens)."""
    _check(load().mofo_token_mean_fwd(_ptr(x), _ptr(weights), B, N, D, _ptr(pooled), _stream()), "mofo_token_mean_fwd")
    return pooled


def token_mean_bwd(dpooled, B, N, D, dx_f32, dx_bf16, bf16_row_scale=None, weights=None):
    _check(load().mofo_token_mean_bwd(_ptr(dpooled), _ptr(weights), B, N, D, _ptr(dx_f32), _ptr(dx_bf16), _ptr(bf16_row_scale),
                                      _stream()), "mofo_token_mean_bwd")


def box_tokens(boxes, frames, size, mode, want_weights=True):
    """boxes int64 [B, frames, 4] (CUDA) -> (inbox u8 [B, N], weights f32 [B, N] | None)."""
    assert boxes.dtype == torch.int64 and boxes.is_cuda and boxes.is_contiguous() and boxes.shape[1:] == (frames, 4)
    B = boxes.shape[0]
    N = (frames // 2) * (size // 16) ** 2
    inbox = torch.empty(B, N, dtype=torch.uint8, device=boxes.device)
    weights = torch.empty(B, N, dtype=torch.float32, device=boxes.device) if want_weights else None
    _check(load().mofo_box_tokens(_ptr(boxes), B, frames, size, mode, _ptr(inbox), _ptr(weights), _stream()), "mofo_box_tokens")
    return inbox, weights


def masked_softmax_fwd(S, key_allowed, scale, P):
    """P bf16 [B, ..., Nq, Nk] = softmax_k(scale * S) over the keys with key_allowed u8 [B, Nk] != 0 (0 elsewhere); S f32, same shape."""
    B, Nk = key_allowed.shape
    assert S.dtype == torch.float32 and P.dtype == torch.bfloat16 and key_allowed.dtype == torch.uint8
    assert S.is_contiguous() and P.is_contiguous() and key_allowed.is_contiguous() and S.shape == P.shape and S.shape[0] == B and S.shape[-1] == Nk
    _check(load().mofo_masked_softmax_fwd(_ptr(S), _ptr(key_allowed), B, S.numel() // (B * Nk), Nk, float(scale), _ptr(P), _stream()),
           "mofo_masked_softmax_fwd")
    return P


def masked_softmax_bwd(P, dP, scale, dS):
    """dS bf16 = scale * P * (dP - sum_k P * dP) row by row; P bf16, dP f32, all [..., Nk] contiguous."""
    assert P.dtype == torch.bfloat16 and dP.dtype == torch.float32 and dS.dtype == torch.bfloat16
    assert P.is_contiguous() and dP.is_contiguous() and dS.is_contiguous() and P.shape == dP.shape == dS.shape
    Nk = P.shape[-1]
    _check(load().mofo_masked_softmax_bwd(_ptr(P), _ptr(dP), P.numel() // Nk, Nk, float(scale), _ptr(dS), _stream()), "mofo_masked_softmax_bwd")
    return dS


def cast_f32_bf16(src, dst):
    """dst bf16 [M, N] = src f32 [M, N]; either may be a column slice (row stride > N)."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.shape == dst.shape and src.dim() == 2
    assert src.stride(1) == 1 and dst.stride(1) == 1
    _check(load().mofo_cast_f32_bf16(_ptr(src), src.stride(0), src.shape[0], src.shape[1], _ptr(dst), dst.stride(0), _stream()),
           "mofo_cast_f32_bf16")
    return dst


def target_mse(video, msk_idx, pred, loss_partials, loss, dpred, normalize_target=True, grad_scale=1.0,
               labels_out=None):
    B, _, frames, size, _ = video.shape
    n_msk = msk_idx.shape[1]
    _check(load().mofo_target_mse(_ptr(video), _ptr(msk_idx), _ptr(pred), B, n_msk, frames, size,
                                  1 if normalize_target else 0, float(grad_scale), _ptr(loss_partials), _ptr(loss),
                                  _ptr(dpred), _ptr(labels_out), _stream()), "mofo_target_mse")


def cast_weight(W, W_bf16=None, Wt_bf16=None):
    R = W.shape[0]
    Cc = W.numel() // R
    assert W.dtype == torch.float32 and W.is_contiguous()
    _check(load().mofo_cast_weight(_ptr(W), R, Cc, _ptr(W_bf16), _ptr(Wt_bf16), _stream()), "mofo_cast_weight")


def pack_qkv_bias(q_bias, v_bias, out):
    _check(load().mofo_pack_qkv_bias(_ptr(q_bias), _ptr(v_bias), q_bias.numel(), _ptr(out), _stream()), "mofo_pack_qkv_bias")


def colsum_bf16(X, M, N, out):
    """out[N] += column sums of X[:M, :N] (X may be a column slice view; stride(0) is the leading dimension)."""
    assert X.dtype == torch.bfloat16 and X.stride(1) == 1 and out.dtype == torch.float32
    _check(load().mofo_colsum_bf16(_ptr(X), X.stride(0), M, N, _ptr(out), _stream()), "mofo_colsum_bf16")


def sq_norm_f32(x, out):
    _check(load().mofo_sq_norm_f32(_ptr(x), x.numel(), _ptr(out), _stream()), "mofo_sq_norm_f32")


ADAMW_TILE = (32, 128)      # rows x cols of a 2-D tile of mofo_adamw_step
ADAMW_RUN = 4096            # elements of a 1-D tile


def adamw_step(params, grads, exp_avg, exp_avg_sq, w16, segs, tiles, hyper, clip_coef=None, loss_guard=None, sq_norm_out=None):
    _check(load().mofo_adamw_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(w16), _ptr(segs),
                                  _ptr(tiles), tiles.shape[0], _ptr(hyper), _ptr(clip_coef), _ptr(loss_guard),
                                  _ptr(sq_norm_out), _stream()), "mofo_adamw_step")


def clip_preprocess(frames_u8, crops, boxes_in, out_size, clip_out, boxes_out):
    """frames_u8 uint8 [B,T,H,W,3], crops int32 [B,4], boxes_in f64 [B,T,4] | None -> clip_out f32 [B,3,T,S,S], boxes_out f64 [B,T,4]."""
    B, T, H, W, Cc = frames_u8.shape
    assert Cc == 3 and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous()
    assert crops.dtype == torch.int32 and crops.shape == (B, 4) and crops.is_contiguous()
    assert boxes_in is None or (boxes_in.dtype == torch.float64 and boxes_in.shape == (B, T, 4) and boxes_in.is_contiguous())
    _check(load().mofo_clip_preprocess(_ptr(frames_u8), B, T, H, W, _ptr(crops), _ptr(boxes_in), out_size, _ptr(clip_out),
                                       _ptr(boxes_out), _stream()), "mofo_clip_preprocess")
    return clip_out, boxes_out


def normalize_u8(clip_u8, out):
    """uint8 [B,3,T,H,W] -> ImageNet-normalised f32 (same shape), on the current stream."""
    B, Cc, frames, size, _ = clip_u8.shape
    assert Cc == 3 and clip_u8.dtype == torch.uint8 and clip_u8.is_contiguous() and out.dtype == torch.float32
    _check(load().mofo_normalize_u8(_ptr(clip_u8), B, frames, size, _ptr(out), _stream()), "mofo_normalize_u8")
    return out


def motion_map(flows_u8, ws, border, out):
    """flows_u8 uint8 [T,H,W,C>=2] -> out uint8 [T,H,W,OC] (motion-boundary magnitude map, value replicated over OC)."""
    T, H, W, Cc = flows_u8.shape
    assert flows_u8.dtype == torch.uint8 and flows_u8.is_contiguous() and flows_u8.is_cuda
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.shape[:3] == (T, H, W) and out.dim() == 4
    _check(load().mofo_motion_map(_ptr(flows_u8), T, H, W, Cc, ws, border, _ptr(out), out.shape[3], _stream()), "mofo_motion_map")
    return out


def motion_box_filter(frames_u8, w_before, w_after, remove_thrd, std_k, std_eps, work, stats, filtered, gray):
    """frames_u8 uint8 [T,H,W,3]; w_* f64 [r+1] weights by distance -> filtered uint8 [T,H,W,3], gray uint8 [T,H,W]."""
    T, H, W, Cc = frames_u8.shape
    assert Cc == 3 and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.is_cuda
    assert w_before.dtype == torch.float64 and w_after.dtype == torch.float64 and w_before.is_cuda and w_after.is_cuda
    assert work.dtype == torch.uint8 and work.numel() >= 2 * frames_u8.numel() and stats.dtype == torch.int64 and stats.numel() >= 4 * T
    assert filtered.shape == frames_u8.shape and filtered.dtype == torch.uint8 and filtered.is_contiguous()
    assert gray.shape == (T, H, W) and gray.dtype == torch.uint8 and gray.is_contiguous()
    _check(load().mofo_motion_box_filter(_ptr(frames_u8), T, H, W, _ptr(w_before), w_before.numel() - 1, _ptr(w_after),
                                         w_after.numel() - 1, remove_thrd, std_k, std_eps, _ptr(work), _ptr(stats), _ptr(filtered),
                                         _ptr(gray), _stream()), "mofo_motion_box_filter")
    return filtered, gray
