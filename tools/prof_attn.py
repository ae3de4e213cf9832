"""Small driver for ncu: attention fwd + bwd at the decoder shape (S=1568, H=6) on a few clips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mofo_b200 import _lib
B, S, H = int(os.environ.get("PB", 8)), int(os.environ.get("PS", 1568)), int(os.environ.get("PH", 6))
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * H * 64, device="cuda")).bfloat16()
out = torch.empty(B * S, H * 64, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(B, H, S, device="cuda")
dout = torch.randn(B * S, H * 64, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device="cuda")
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    _lib.attn_fwd(qkv, B, S, H, 0.125, out, lse)
    e[1].record()
    _lib.attn_bwd(qkv, out, dout, lse, B, S, H, 0.125, dqkv, delta)
    e[2].record()
    torch.cuda.synchronize()
    fl = 4.0 * B * H * S * S * 64
    print(f"fwd {e[0].elapsed_time(e[1]):.3f} ms ({fl / e[0].elapsed_time(e[1]) / 1e9:.0f} TF/s)  bwd {e[1].elapsed_time(e[2]):.3f} ms ({2.5 * fl / e[1].elapsed_time(e[2]) / 1e9:.0f} TF/s)")
