"""Time the fp32 conv paths (AlignConv / ORConv2d at P3): conv_tf32x3_kernel (tcgen05, 3 x TF32) against the SIMT kernel
of csrc/conv_f32.cu (alignconv._FORCE_SIMT_F32)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.alignconv import alignconv_forward
from s2anet_b200.orn import orconv_forward
from oracle import oracle as O
dev = "cuda:0"
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
for B in (1, 8):
    H = 128
    x = torch.randn(B, 256, H, H, device=dev)
    anc = torch.from_numpy(synth.refined_anchors(B, H, H, 8, 1)).to(dev)
    w = torch.randn(256, 256, 3, 3, device=dev) * 0.01
    wo = torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01
    fl = 2.0 * B * H * H * 256 * 2304
    from s2anet_b200 import alignconv
    xcl = x.contiguous(memory_format=torch.channels_last)
    for simt in (False, True):
        alignconv._FORCE_SIMT_F32 = simt
        ms = t(lambda: alignconv_forward(x, anc, w, 8))
        ms1 = t(lambda: alignconv_forward(xcl, anc, w, 8))
        ms2 = t(lambda: orconv_forward(x, wo, idx, None, with_pool=True))
        print("fp32 P3 B=%d %s: alignconv %.3f ms %.1f TF/s (channels_last input %.3f ms) | orconv %.3f ms %.1f TF/s" % (
            B, "simt  " if simt else "tf32x3", ms, fl / ms / 1e9, ms1, ms2, fl / ms2 / 1e9))
    alignconv._FORCE_SIMT_F32 = False
