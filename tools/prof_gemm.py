"""Small driver for ncu / timing: the step's main GEMM shapes (ViT-B, B=32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mofo_b200 import _lib
torch.manual_seed(0)
dev = "cuda"
def t(*s, dt=torch.bfloat16): return (torch.randn(*s, device=dev) * 0.5).to(dt)
Md, Dd, Me, D = 50176, 384, 5120, 768
cases = []
# name, A, B, epi, kwargs, out shapes
h2 = t(Md, Dd); W1 = t(4 * Dd, Dd); b1 = t(4 * Dd, dt=torch.float32)
u = torch.empty(Md, 4 * Dd, dtype=torch.bfloat16, device=dev); a = torch.empty_like(u)
cases.append(("dec_fc1_gelu", lambda: _lib.gemm_tn(h2, W1, _lib.EPI_BIAS_GELU_BF16, u, out1=a, bias=b1), 2.0 * Md * 4 * Dd * Dd))
W2 = t(Dd, 4 * Dd); b2 = t(Dd, dt=torch.float32); xm = t(Md, Dd, dt=torch.float32); xo = torch.empty_like(xm)
cases.append(("dec_fc2_resid", lambda: _lib.gemm_tn(a, W2, _lib.EPI_BIAS_RESID_F32, xo, bias=b2, resid=xm), 2.0 * Md * 4 * Dd * Dd))
dx = t(Md, Dd); W2t = t(4 * Dd, Dd); du = torch.empty_like(u)
cases.append(("dec_fc2_dgrad_gelubwd", lambda: _lib.gemm_tn(dx, W2t, _lib.EPI_GELU_BWD_BF16, du, aux=u), 2.0 * Md * 4 * Dd * Dd))
W1t = t(Dd, 4 * Dd); dh = torch.empty(Md, Dd, dtype=torch.bfloat16, device=dev)
cases.append(("dec_fc1_dgrad_plain", lambda: _lib.gemm_tn(du, W1t, _lib.EPI_PLAIN_BF16, dh), 2.0 * Md * 4 * Dd * Dd))
Wq = t(3 * Dd, Dd); bq = t(3 * Dd, dt=torch.float32); qkv = torch.empty(Md, 3 * Dd, dtype=torch.bfloat16, device=dev)
cases.append(("dec_qkv_bias", lambda: _lib.gemm_tn(h2, Wq, _lib.EPI_BIAS_BF16, qkv, bias=bq), 2.0 * Md * 3 * Dd * Dd))
he = t(Me, D); We1 = t(4 * D, D); be1 = t(4 * D, dt=torch.float32)
ue = torch.empty(Me, 4 * D, dtype=torch.bfloat16, device=dev); ae = torch.empty_like(ue)
cases.append(("enc_fc1_gelu", lambda: _lib.gemm_tn(he, We1, _lib.EPI_BIAS_GELU_BF16, ue, out1=ae, bias=be1), 2.0 * Me * 4 * D * D))
We2 = t(D, 4 * D); be2 = t(D, dt=torch.float32); xe = t(Me, D, dt=torch.float32); xeo = torch.empty_like(xe)
cases.append(("enc_fc2_resid", lambda: _lib.gemm_tn(ae, We2, _lib.EPI_BIAS_RESID_F32, xeo, bias=be2, resid=xe), 2.0 * Me * 4 * D * D))
dW1 = torch.zeros(4 * Dd, Dd, device=dev); db1 = torch.zeros(4 * Dd, device=dev)
cases.append(("dec_fc1_wgrad", lambda: _lib.gemm_wgrad(du, h2, dW1, dbias=db1), 2.0 * Md * 4 * Dd * Dd))
cases.append(("dec_fc1_wgrad_nobias", lambda: _lib.gemm_wgrad(du, h2, dW1), 2.0 * Md * 4 * Dd * Dd))
dW2 = torch.zeros(Dd, 4 * Dd, device=dev); db2 = torch.zeros(Dd, device=dev)
cases.append(("dec_fc2_wgrad", lambda: _lib.gemm_wgrad(dx, a, dW2, dbias=db2), 2.0 * Md * 4 * Dd * Dd))
cases.append(("dec_fc2_wgrad_nobias", lambda: _lib.gemm_wgrad(dx, a, dW2), 2.0 * Md * 4 * Dd * Dd))
dWe1 = torch.zeros(4 * D, D, device=dev); dbe1 = torch.zeros(4 * D, device=dev)
cases.append(("enc_fc1_wgrad", lambda: _lib.gemm_wgrad(ue, he, dWe1, dbias=dbe1), 2.0 * Me * 4 * D * D))
cases.append(("enc_fc1_wgrad_nobias", lambda: _lib.gemm_wgrad(ue, he, dWe1), 2.0 * Me * 4 * D * D))
x32 = t(Md, Dd, dt=torch.float32); g = torch.ones(Dd, device=dev); bb = torch.zeros(Dd, device=dev)
y = torch.empty(Md, Dd, dtype=torch.bfloat16, device=dev); mean = torch.empty(Md, device=dev); rstd = torch.empty(Md, device=dev)
cases.append(("dec_ln_fwd", lambda: _lib.layernorm_fwd(x32, g, bb, y, mean, rstd, Md, Dd), 0))
dxo = torch.empty_like(x32); dxb = torch.empty_like(y); dg = torch.zeros(Dd, device=dev); dbt = torch.zeros(Dd, device=dev)
cases.append(("dec_ln_bwd", lambda: _lib.layernorm_bwd(y, x32, g, mean, rstd, xm, Md, Dd, dxo, dxb, dg, dbt), 0))
xe32 = t(Me, D, dt=torch.float32); ge = torch.ones(D, device=dev); bbe = torch.zeros(D, device=dev)
ye = torch.empty(Me, D, dtype=torch.bfloat16, device=dev); me_ = torch.empty(Me, device=dev); re_ = torch.empty(Me, device=dev)
_lib.layernorm_fwd(xe32, ge, bbe, ye, me_, re_, Me, D)
dxe = torch.empty_like(xe32); dxeb = torch.empty_like(ye); dge = torch.zeros(D, device=dev); dbe = torch.zeros(D, device=dev)
cases.append(("enc_ln_bwd", lambda: _lib.layernorm_bwd(ye, xe32, ge, me_, re_, xe, Me, D, dxe, dxeb, dge, dbe), 0))
only = os.environ.get("ONLY")
for name, fn, fl in cases:
    if only and only not in name:
        continue
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get("N", 5))
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:26s} {ms * 1e3:8.1f} us" + (f"  {fl / ms / 1e9:7.0f} TF/s" if fl else ""))
