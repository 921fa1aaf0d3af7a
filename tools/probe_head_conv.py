"""Time the prediction convs (256 -> 15, 3x3 and 1x1) and a tower conv at the bench shape (all 5 levels, batch 8)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200.conv_tc import conv2d_forward_tc_multi
dev = "cuda:0"
B = 8
xs = [torch.randn(B, 256, 1024 // s, 1024 // s, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for s in (8, 16, 32, 64, 128)]
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for Co, ks in ((256, 3), (15, 3), (15, 1)):
    w = (torch.randn(Co, 256, ks, ks, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(Co, device=dev)
    us = t(lambda: conv2d_forward_tc_multi(xs, w, bias, relu=True))
    print("conv %dx%d 256->%d: %.1f us" % (ks, ks, Co, us), flush=True)
    if os.environ.get("S2A_TC_DEBUG"):
        conv2d_forward_tc_multi(xs, w, bias, relu=True); torch.cuda.synchronize()
