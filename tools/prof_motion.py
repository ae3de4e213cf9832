"""Small driver for ncu: one SSv2-shaped flow video through the motion-box pixel stages (9 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mofo_b200 import motion_boxes as mb
g = torch.Generator().manual_seed(7)
flows = torch.randint(96, 160, (48, 240, 320, 3), dtype=torch.uint8, generator=g)
flows[:, 60:150, 80:200, :2] += 70
flows = flows.cuda()
flt = mb.MotionMapFilter()
for _ in range(3):
    filt, gray = flt.filter(mb.motion_map(flows))
torch.cuda.synchronize()
print(int(gray.max()))
