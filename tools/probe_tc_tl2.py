"""MMA-warp timeline probe.  Needs a library built with the clock probes compiled in:
    S2A_NVCC_EXTRA=-DS2A_TC_TIMELINE python -m s2anet_b200.build --force
then run with the S2A_TC_DEBUG values to try (8 = timeline; +1 no weight TMA, +2 no gather, +32 no epilogue)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
dev = "cuda:0"
B, s, H = 8, 8, 128
dt = torch.bfloat16
x = torch.randn(B, 256, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
anc = torch.from_numpy(synth.refined_anchors(B, H, H, s, 1)).to(dev)
w = (torch.randn(256, 256, 3, 3, device=dev) * 0.01).to(dt)
for _ in range(3): alignconv_forward(x, anc, w, s)
torch.cuda.synchronize()
for dbg in sys.argv[1:]:
    os.environ["S2A_TC_DEBUG"] = dbg
    print("debug", dbg, flush=True)
    alignconv_forward(x, anc, w, s)
    torch.cuda.synchronize()
