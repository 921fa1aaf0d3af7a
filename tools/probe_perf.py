"""Quick per-kernel timing probe (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
from s2anet_b200.orn import orconv_forward
from oracle import oracle as O
dev = "cuda:0"
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
for dt in (torch.bfloat16, torch.float16):
    for B, s in ((1, 8), (8, 8), (8, 16), (8, 32), (8, 64), (8, 128)):
        H = 1024 // s
        x = torch.randn(B, 256, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        anc = torch.from_numpy(synth.refined_anchors(B, H, H, s, 1)).to(dev)
        w = (torch.randn(256, 256, 3, 3, device=dev) * 0.01).to(dt)
        wo = (torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01).to(dt)
        fl = 2.0 * B * H * H * 256 * 2304
        ms = timeit(lambda: alignconv_forward(x, anc, w, s))
        ms2 = timeit(lambda: orconv_forward(x, wo, idx, None, with_pool=True))
        # regular anchors (theta=0, 4s square): the friendliest gather
        anc0 = torch.from_numpy(np.broadcast_to(synth.grid_anchors(H, H, s), (B, H, H, 5)).copy()).to(dev)
        ms3 = timeit(lambda: alignconv_forward(x, anc0, w, s))
        print("%s B=%d stride=%3d H=%3d: alignconv %.3f ms %.0f TF/s | grid-anchors %.3f ms %.0f TF/s | orconv %.3f ms %.0f TF/s"
              % (str(dt)[6:], B, s, H, ms, fl / ms / 1e9, ms3, fl / ms3 / 1e9, ms2, fl / ms2 / 1e9))
    conv = torch.nn.Conv2d(256, 256, 3, padding=1).to(dev).to(dt).to(memory_format=torch.channels_last)
    x = torch.randn(8, 256, 128, 128, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        ms = timeit(lambda: conv(x))
    print("cuDNN conv3x3 256->256 B=8 P3 %s: %.3f ms %.0f TF/s" % (str(dt)[6:], ms, 2.0 * 8 * 16384 * 256 * 2304 / ms / 1e9))
