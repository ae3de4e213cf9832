"""GPU parity of the tcgen05 GEMM kernels (all fused epilogues, wgrad) against torch fp32 matmul on the same
bf16-rounded inputs."""
import math

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


SHAPES = [(128, 128, 64), (256, 256, 128), (5120, 768, 768), (300, 384, 192), (128, 1152, 384), (640, 64, 128),
          (1000, 2304, 768), (4096, 1536, 384), (513, 192, 1536), (5120, 3072, 768), (32, 128, 1536),
          (5120, 2304, 768), (50176, 384, 384), (384, 1152, 384)]   # incl. the CTA-pair tile widths 224 / 192 / 256


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tn_bias_and_plain(lib, M, N, K):
    torch.manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    lib.gemm_tn(A, B, lib.EPI_PLAIN_BF16, out)
    e = rel(out, ref)
    assert e < 6e-3, f"plain rel err {e}"
    lib.gemm_tn(A, B, lib.EPI_BIAS_BF16, out, bias=bias)
    e = rel(out, ref + bias)
    assert e < 6e-3, f"bias rel err {e}"


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1000, 1536, 384), (5120, 768, 3072), (5120, 3072, 768), (6272, 384, 1536)])
def test_gemm_tn_fused_epilogues(lib, M, N, K):
    torch.manual_seed(1)
    A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t()
    # bias + GELU: out1 = gelu(u), out0 = gelu'(u) with u = bf16(acc + bias)
    dg = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"); a = torch.empty_like(dg)
    lib.gemm_tn(A, B, lib.EPI_BIAS_GELU_BF16, dg, out1=a, bias=bias)
    u = (ref + bias).bfloat16().float().requires_grad_(True)
    act = torch.nn.functional.gelu(u)
    act.sum().backward()
    assert rel(a, act.detach()) < 6e-3
    assert rel(dg, u.grad) < 6e-3
    # bias + residual, f32 out
    res = torch.randn(M, N, device="cuda"); o32 = torch.empty(M, N, device="cuda")
    lib.gemm_tn(A, B, lib.EPI_BIAS_RESID_F32, o32, bias=bias, resid=res)
    assert rel(o32, ref + bias + res) < 1e-5
    # GELU backward: multiply by the saved derivative
    g = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    lib.gemm_tn(A, B, lib.EPI_GELU_BWD_BF16, g, aux=dg)
    assert rel(g, ref * dg.float()) < 6e-3
    # bias + pos + row remap (patch-embed / encoder_to_decoder epilogue)
    group, out_group = 50, 80
    Mg = (M // group) * group
    ntok = 300
    pos = torch.randn(ntok, N, device="cuda")
    ridx = torch.randint(0, ntok, (Mg,), device="cuda").int()
    outp = torch.zeros((Mg // group) * out_group, N, device="cuda")
    lib.gemm_tn(A[:Mg], B, lib.EPI_BIAS_POS_F32, outp, bias=bias, pos=pos, row_idx=ridx, group_rows=group, out_group_rows=out_group)
    want = (ref[:Mg] + bias + pos[ridx.long()]).reshape(-1, group, N)
    got = outp.reshape(-1, out_group, N)
    assert rel(got[:, :group], want) < 1e-5
    assert got[:, group:].abs().max().item() == 0


def test_gemm_tn_strided_views(lib):
    # A operand as a column slice (lda > K), as used for per-head / sliced activations
    torch.manual_seed(5)
    big = torch.randn(700, 1152, device="cuda").bfloat16()
    A = big[:, 384:768]
    B = (torch.randn(256, 384, device="cuda") / 20).bfloat16()
    out = torch.empty(700, 256, dtype=torch.bfloat16, device="cuda")
    lib.gemm_tn(A, B, lib.EPI_PLAIN_BF16, out)
    assert rel(out, A.float() @ B.float().t()) < 6e-3


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (5120, 768, 1536), (1000, 384, 192), (6272, 1536, 384),
                                   (5120, 2304, 768), (333, 64, 128), (4096, 384, 1536), (64, 1536, 384),
                                   (6272, 1152, 384), (1000, 384, 384), (5120, 3072, 768), (777, 768, 3072), (300, 256, 192),
                                   (2000, 1536, 1536), (500, 1152, 128)])
def test_gemm_wgrad(lib, M, N, K):
    """Covers every tile shape of mofo_gemm_wgrad: 2 n-tiles x 192 (N % 256 == 0, K % 192 == 0), 3 n-tiles x 128
    (N % 384 == 0), and the single-accumulator 256 / 192 / 128 tiles."""
    torch.manual_seed(M + N + K)
    dY = torch.randn(M, N, device="cuda").bfloat16(); X = torch.randn(M, K, device="cuda").bfloat16()
    dW = torch.zeros(N, K, device="cuda")
    lib.gemm_wgrad(dY, X, dW)
    ref = dY.float().t() @ X.float()
    e = rel(dW, ref)
    assert e < 2e-5, f"wgrad rel err {e}"
    lib.gemm_wgrad(dY, X, dW)          # accumulates
    assert rel(dW, 2 * ref) < 2e-5
    # fused bias gradient (column sums of dY), with a skipped column window like the qkv bias
    dW.zero_()
    db = torch.zeros(N, device="cuda")
    lo, hi = (N // 3, 2 * N // 3) if N % 3 == 0 else (0, 0)
    lib.gemm_wgrad(dY, X, dW, dbias=db, skip=(lo, hi))
    want = dY.float().sum(0)
    want[lo:hi] = 0
    assert rel(dW, ref) < 2e-5
    assert rel(db, want) < 2e-5, f"dbias rel err {rel(db, want)}"
    assert db[lo:hi].abs().max().item() == 0 if hi > lo else True


def test_gemm_rejects_bad_args(lib):
    A = torch.randn(64, 60, device="cuda").bfloat16(); B = torch.randn(64, 60, device="cuda").bfloat16()
    out = torch.empty(64, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(lib.MofoError):
        lib.gemm_tn(A, B, lib.EPI_PLAIN_BF16, out)      # K % 8 != 0


@pytest.mark.parametrize("M,D", [(5120, 768), (6272, 384), (320, 192), (2560, 1024), (300, 512)])      # D = 1024 / 512: the 256-wide k-tile group
def test_gemm_wgrad_grouped_matches_individual_calls(lib, M, D):
    """The four weight gradients of a transformer block in one grouped launch (fc2, fc1, proj, qkv with its skipped bias window)."""
    torch.manual_seed(M + D)
    t = lambda *s: torch.randn(*s, device="cuda").bfloat16()
    dx, a, du, h2, dxb, o, dqkv, h1 = t(M, D), t(M, 4 * D), t(M, 4 * D), t(M, D), t(M, D), t(M, D), t(M, 3 * D), t(M, D)
    shapes = [(D, 4 * D), (4 * D, D), (D, D), (3 * D, D)]
    skips = [(0, 0), (0, 0), (0, 0), (D, 2 * D)]
    pairs = [(dx, a), (du, h2), (dxb, o), (dqkv, h1)]
    want_w = [torch.zeros(n, k, device="cuda") for n, k in shapes]; want_b = [torch.zeros(n, device="cuda") for n, _ in shapes]
    for (dY, X), w, b, sk in zip(pairs, want_w, want_b, skips):
        lib.gemm_wgrad(dY, X, w, dbias=b, skip=sk)
    got_w = [torch.zeros(n, k, device="cuda") for n, k in shapes]; got_b = [torch.zeros(n, device="cuda") for n, _ in shapes]
    lib.gemm_wgrad_grouped([(dY, X, w, b, sk) for (dY, X), w, b, sk in zip(pairs, got_w, got_b, skips)], M)
    for i in range(4):
        assert rel(got_w[i], want_w[i]) < 1e-5, i
        assert rel(got_b[i], want_b[i]) < 1e-5, i
        ref = pairs[i][0].float().t() @ pairs[i][1].float()
        assert rel(got_w[i], ref) < 2e-5
    assert got_b[3][D:2 * D].abs().max().item() == 0
    # a shape the grouped kernel does not take (K % 192 != 0 and K % 256 != 0) falls back to individual launches
    dY2, X2 = t(256, 128), t(256, 128)
    w2 = torch.zeros(128, 128, device="cuda")
    lib.gemm_wgrad_grouped([(dY2, X2, w2, None, (0, 0))], 256)
    assert rel(w2, dY2.float().t() @ X2.float()) < 2e-5


def test_gemm_on_per_head_views_with_ragged_n_and_k(lib):
    """The products of the box-focused classifier's cross attention (3 heads of 256 over 1568 tokens): operands are column /
    row slices of wider matrices, N = 1568 and K = 1568 are not multiples of the tile sizes."""
    torch.manual_seed(9)
    N, D, hd = 1568, 768, 256
    qkv = torch.randn(2 * N, 3 * D, device="cuda").bfloat16()
    rows = qkv[N:2 * N]
    q, k = rows[:, hd:2 * hd], rows[:, D + hd:D + 2 * hd]
    zero = torch.zeros(N, N, device="cuda")
    S = torch.empty(N, N, device="cuda")
    lib.gemm_tn(q, k, lib.EPI_BIAS_RESID_F32, S, resid=zero)                 # f32 scores, M = N = 1568, K = 256
    ref = q.float() @ k.float().t()
    assert rel(S, ref) < 1e-5
    P = (torch.rand(N, N, device="cuda") / N).bfloat16()
    kvT = (torch.randn(2 * D, N, device="cuda")).bfloat16()
    o = torch.zeros(N, D, dtype=torch.bfloat16, device="cuda")
    bias = torch.randn(D, device="cuda")
    lib.gemm_tn(P, kvT[D + hd:D + 2 * hd], lib.EPI_BIAS_BF16, o[:, hd:2 * hd], bias=bias[hd:2 * hd])     # K = 1568
    ref = P.float() @ kvT[D + hd:D + 2 * hd].float().t() + bias[hd:2 * hd]
    assert rel(o[:, hd:2 * hd], ref) < 4e-3
    assert o[:, :hd].abs().max().item() == 0 and o[:, 2 * hd:].abs().max().item() == 0
    # operand-swapped projection: K^T, V^T [2D, N] = W_kv @ h^T  (N = 1568 output columns)
    h = torch.randn(N, D, device="cuda").bfloat16(); w = (torch.randn(2 * D, D, device="cuda") / math.sqrt(D)).bfloat16()
    out = torch.empty(2 * D, N, dtype=torch.bfloat16, device="cuda")
    lib.gemm_tn(w, h, lib.EPI_PLAIN_BF16, out)
    assert rel(out, w.float() @ h.float().t()) < 4e-3
    # dK = dS^T Q into a column slice of an f32 [N, 2D] matrix (ldw > K), accumulated
    dS = torch.randn(N, N, device="cuda").bfloat16()
    dkv = torch.zeros(N, 2 * D, device="cuda")
    lib.gemm_wgrad(dS, q, dkv[:, hd:2 * hd])
    ref = dS.float().t() @ q.float()
    assert rel(dkv[:, hd:2 * hd], ref) < 1e-4
    assert dkv[:, :hd].abs().max().item() == 0 and dkv[:, 2 * hd:].abs().max().item() == 0
