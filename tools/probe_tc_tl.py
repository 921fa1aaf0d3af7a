"""MMA-warp timeline probe.  Needs a library built with the clock probes compiled in:
    S2A_NVCC_EXTRA=-DS2A_TC_TIMELINE python -m s2anet_b200.build --force
then run with the S2A_TC_DEBUG values to try (8 = timeline; +1 no weight TMA, +2 no gather, +32 no epilogue)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200.orn import orconv_forward
from oracle import oracle as O
dev = "cuda:0"
B, H = 8, 128
dt = torch.bfloat16
x = torch.randn(B, 256, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
wo = (torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01).to(dt)
idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
for _ in range(3): orconv_forward(x, wo, idx, None, with_pool=True)
torch.cuda.synchronize()
for dbg in sys.argv[1:]:
    os.environ["S2A_TC_DEBUG"] = dbg
    print("debug", dbg, flush=True)
    for _ in range(2): orconv_forward(x, wo, idx, None, with_pool=True)
    torch.cuda.synchronize()
