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


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (2, 160, 3), (1, 1568, 2), (3, 16, 2), (1, 129, 1), (2, 384, 6), (1, 255, 2)])
def test_attention_fwd_bwd(lib, B, S, H):
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
