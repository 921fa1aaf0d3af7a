"""One AlignConv + one ORConv (batch 8, P3) + one narrow 256->15 prediction conv (batch 8, all five levels) per
iteration, three iterations (for ncu --set full -k regex:conv_tc_kernel --launch-skip 6 --launch-count 3)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
from s2anet_b200.orn import orconv_forward
from s2anet_b200.conv_tc import conv2d_forward_tc_multi
from oracle import oracle as O
dev = "cuda:0"
B, s, H = 8, 8, 128
dt = torch.bfloat16
x = torch.randn(B, 256, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
anc = torch.from_numpy(synth.refined_anchors(B, H, H, s, 1)).to(dev)
w = (torch.randn(256, 256, 3, 3, device=dev) * 0.01).to(dt)
wo = (torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01).to(dt)
idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
xs = [torch.randn(B, 256, 1024 // t, 1024 // t, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for t in (8, 16, 32, 64, 128)]
wn = (torch.randn(15, 256, 3, 3, device=dev) * 0.02).to(dt)
bn = torch.zeros(15, device=dev)
for _ in range(3):
    y = alignconv_forward(x, anc, w, s)
    z = orconv_forward(y, wo, idx, None, with_pool=True)
    n = conv2d_forward_tc_multi(xs, wn, bn, relu=False)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(z[0].float().abs().mean()))
