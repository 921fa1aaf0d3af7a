"""One tcgen05 dgrad (nine launches) + wgrad call at P3, batch 8 -- the target of the ncu captures in profiles/."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200.conv_tc import deform_conv_dgrad_tc, deform_conv_wgrad_tc
dev = "cuda:0"
B = 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 256, 128, 128, generator=g).to(dev).bfloat16().contiguous(memory_format=torch.channels_last)
w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).to(dev).bfloat16()
off = (torch.randn(B, 18, 128, 128, generator=g) * 1.2).to(dev)
gy = torch.randn(B, 256, 128, 128, generator=g).to(dev).bfloat16().contiguous(memory_format=torch.channels_last)
for _ in range(2):
    gi = deform_conv_dgrad_tc(gy, off, w)
    gw = deform_conv_wgrad_tc(x, off, gy)
torch.cuda.synchronize()
print(float(gw.abs().max()))
