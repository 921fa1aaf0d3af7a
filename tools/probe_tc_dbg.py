"""Timing experiments on the tcgen05 conv kernel (S2A_TC_DEBUG bit flags; results are WRONG by design)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
from s2anet_b200.orn import orconv_forward
from oracle import oracle as O
dev = "cuda:0"
B, s, H = 8, 8, 128
dt = torch.bfloat16
x = torch.randn(B, 256, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
anc = torch.from_numpy(synth.refined_anchors(B, H, H, s, 1)).to(dev)
w = (torch.randn(256, 256, 3, 3, device=dev) * 0.01).to(dt)
wo = (torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01).to(dt)
idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
fl = 2.0 * B * H * H * 256 * 2304
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for dbg in [int(v) for v in sys.argv[1:]] or [0]:
    os.environ["S2A_TC_DEBUG"] = str(dbg)
    ta = t(lambda: alignconv_forward(x, anc, w, s))
    to = t(lambda: orconv_forward(x, wo, idx, None, with_pool=True))
    print("debug=%2d  alignconv %.3f ms %5.0f TF/s | orconv %.3f ms %5.0f TF/s" % (dbg, ta, fl / ta / 1e9, to, fl / to / 1e9), flush=True)
os.environ["S2A_TC_DEBUG"] = "8"
alignconv_forward(x, anc, w, s); torch.cuda.synchronize()
for _ in range(30): orconv_forward(x, wo, idx, None, with_pool=True)
torch.cuda.synchronize()
