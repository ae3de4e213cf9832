"""One training step of the UNMODIFIED reference (baseline/_ref, bf16 autocast, B=32 ViT-B) for an ncu launch list:
python tools/ref_step.py [amp] [batch] [steps]  -- prints the CUDA-event time of the last step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baseline import refrun
amp = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = refrun.create_model("pretrain_videomae_base_patch16_224").to(dev)
opt = refrun.create_optimizer(model, lr=1.5e-4 * B / 256)
scaler = refrun.make_scaler(dev)
batches = refrun.synthetic_batches(B, 1, 1234, dev)
for i in range(steps):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push(f"step{i}")
    refrun.train_epoch(model, opt, scaler, batches, 1, dev, amp)
    torch.cuda.nvtx.range_pop()
    e1.record(); torch.cuda.synchronize()
    print(f"reference step {i}: {e0.elapsed_time(e1):.2f} ms ({amp}, B={B})")
