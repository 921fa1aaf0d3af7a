"""One fp32 AlignConv launch at P3, batch 8 (conv_tf32x3_kernel) -- the target of the ncu capture in profiles/."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
dev = "cuda:0"
B, H = 8, 128
x = torch.randn(B, 256, H, H, device=dev).contiguous(memory_format=torch.channels_last)
anc = torch.from_numpy(synth.refined_anchors(B, H, H, 8, 1)).to(dev)
w = torch.randn(256, 256, 3, 3, device=dev) * 0.01
for _ in range(3):
    y = alignconv_forward(x, anc, w, 8)
torch.cuda.synchronize()
print(float(y.abs().max()))
