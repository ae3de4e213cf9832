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

if os.environ.get("TRACE"):      # -DMOFO_ATTN_TRACE build: per-iteration phase timestamps of one mid-grid CTA
    import ctypes, numpy as np
    n_it = (S + 63) // 64
    buf = (ctypes.c_longlong * (64 * 8 * 2))()
    assert _lib.load().mofo_debug_read_trace(buf, 64 * 8 * 2) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(64, 8, 2)[:n_it]
    names = os.environ["TRACE"].split(",")
    print("iteration period (thread 0):", np.diff(t[:, 0, 0])[2:-1].mean())
    for th in (0, 1):
        d = np.diff(t[2:-1, :, th], axis=1)
        print(f"thread {'0' if th == 0 else '255'} phase deltas (clk, mean over iterations):",
              {names[k] if k < len(names) else k: round(float(d[:, k].mean())) for k in range(d.shape[1])})
    print("per-iteration slot 0 (thread 0):", (t[1:, 0, 0] - t[:-1, 0, 0]).tolist())

if os.environ.get("STRACE"):     # -DMOFO_ATTN_TRACE=4 build: phase timestamps of the single-pass forward kernel
    import ctypes, numpy as np
    buf = (ctypes.c_longlong * 64)()
    assert _lib.load().mofo_debug_read_strace(buf, 64) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(32, 2)
    names = ["start", "prologue done", "loads landed (t0 only)", "S0 ready", "-", "softmax0 done", "sync0", "S1+O0 ready", "epilogue0 done",
             "softmax1 done", "sync1", "O1 ready", "epilogue1 done", "dealloc"]
    for th in (0, 1):
        base = t[0, th]
        print(f"thread {'0' if th == 0 else '255'}:", {names[k]: int(t[k, th] - base) for k in range(14) if t[k, th] > 0})
