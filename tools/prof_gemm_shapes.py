"""gemm_tn timing over arbitrary shapes: SHAPES="M,N,K;M,N,K" (PLAIN_BF16 epilogue).  Compare MOFO_GEMM_2CTA=0/1, MOFO_FORCE_BN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mofo_b200 import _lib
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("SHAPES", "8192,8192,8192").split(";")]
for M, N, K in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    fn = lambda: _lib.gemm_tn(A, B, _lib.EPI_PLAIN_BF16, out)
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"M={M} N={N} K={K}: {ms * 1e3:8.1f} us {2.0 * M * N * K / ms / 1e9:7.0f} TF/s")
