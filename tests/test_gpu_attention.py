"""GPU parity of the fused attention kernels (fwd + bwd) against the unfused fp32 formulation of
modeling_finetune.py:85-95 evaluated by torch autograd on the same bf16-rounded qkv."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from mofo_b200 import _lib
    _lib.load()
    return _lib


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def ref_attention(qkv, B, S, H):
    q, k, v = qkv.float().reshape(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    q = q * (64 ** -0.5)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B * S, H * 64)


@pytest.mark.parametrize("path", ["default", "streaming"])
@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (2, 160, 3), (1, 1568, 2), (3, 16, 2), (1, 129, 1), (2, 384, 6), (1, 255, 2),
                                   (32, 160, 12), (2, 192, 2), (3, 64, 1), (2, 161, 2), (5, 40, 3)])
def test_attention_fwd_bwd(lib, B, S, H, path, monkeypatch):
    """``default``: S <= 192 runs the single-pass kernels (attention_small.cu), longer sequences the streaming kernels;
    ``streaming``: MOFO_ATTN_SMALL=0 forces the streaming kernels for every S (both paths stay covered)."""
    if path == "streaming":
        if S > 192:
            pytest.skip("already the streaming path")
        monkeypatch.setenv("MOFO_ATTN_SMALL", "0")
    torch.manual_seed(B * 1000 + S + H)
    qkv = (torch.randn(B * S, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    out = torch.empty(B * S, H * 64, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(B, H, S, device="cuda")
    lib.attn_fwd(qkv, B, S, H, 64 ** -0.5, out, lse)
    qr = qkv.float().requires_grad_(True)
    ref = ref_attention(qr, B, S, H)
    e = rel(out, ref)
    assert e < 1e-2, f"attention fwd rel err {e}"
    # lse (log2 domain)
    q, k, _ = qkv.float().reshape(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((q * 64 ** -0.5) @ k.transpose(-2, -1), dim=-1) * 1.4426950408889634
    assert (lse - lse_ref).abs().max().item() < 2e-2
    dout = torch.randn(B * S, H * 64, device="cuda").bfloat16()
    ref.backward(dout.float())
    dqkv = torch.full((B * S, 3 * H * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    delta = torch.empty(B, H, S, device="cuda")
    lib.attn_bwd(qkv, out, dout, lse, B, S, H, 64 ** -0.5, dqkv, delta)
    g = qr.grad
    D = H * 64
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        e = rel(dqkv[:, sl], g[:, sl])
        assert e < 2e-2, f"attention bwd {name} rel err {e}"


def test_small_attention_row_sums_of_dS_cancel(lib, monkeypatch):
    """Keys with a large common component k0: dQ_i = sum_j dS_ij k_j must not pick up k0, because sum_j dS_ij = 0.  The
    single-pass kernels form delta from the recomputed P (fp32) and dP, so the k0 component of sum_i dQ_i (what the q_bias
    gradient accumulates, modeling_finetune.py:82-84) stays at the level of the fp32 reference; a delta taken from the
    bf16 O of the forward pass (streaming kernels, FlashAttention-style) leaks O(2^-9) of it."""
    B, S, H = 8, 160, 2
    torch.manual_seed(5)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda")
    k0 = torch.randn(H * 64, device="cuda") * 6.0
    qkv[:, H * 64:2 * H * 64] += k0
    qkv = qkv.bfloat16()
    dout = torch.randn(B * S, H * 64, device="cuda").bfloat16()
    qr = qkv.float().requires_grad_(True)
    ref_attention(qr, B, S, H).backward(dout.float())
    want = qr.grad[:, :H * 64].sum(0)

    def run(with_lo=False):
        out = torch.empty(B * S, H * 64, dtype=torch.bfloat16, device="cuda")
        out_lo = torch.empty_like(out) if with_lo else None
        lse = torch.empty(B, H, S, device="cuda")
        lib.attn_fwd(qkv, B, S, H, 64 ** -0.5, out, lse, out_lo=out_lo)
        dqkv = torch.empty(B * S, 3 * H * 64, dtype=torch.bfloat16, device="cuda")
        lib.attn_bwd(qkv, out, dout, lse, B, S, H, 64 ** -0.5, dqkv, torch.empty(B, H, S, device="cuda"), out_lo=out_lo)
        if with_lo:      # out + out_lo carries the attention output to ~2^-16
            full = ref_attention(qkv.float(), B, S, H)
            assert rel(out.float() + out_lo.float(), full) < 0.8 * rel(out.float(), full)   # the rest is P's bf16 rounding
        return dqkv[:, :H * 64].float().sum(0)
    e_small = rel(run(), want)
    monkeypatch.setenv("MOFO_ATTN_SMALL", "0")
    e_stream = rel(run(), want)
    e_stream_lo = rel(run(with_lo=True), want)
    print(f"sum_i dQ_i rel err: single-pass {e_small:.3e}, streaming {e_stream:.3e}, streaming + out_lo {e_stream_lo:.3e}")
    assert e_small < 3e-2, e_small
    assert e_stream_lo < 3e-2 and e_stream_lo < e_stream, (e_stream_lo, e_stream)
