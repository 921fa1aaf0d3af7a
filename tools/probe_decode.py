"""Time select_decode / fam_decode at the bench shape (batch 8, 1024^2, bf16 channels_last, Co padded to 32)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200.decode import fam_decode, select_decode
dev = "cuda:0"
B = 8
g = torch.Generator().manual_seed(0)
cls, reg = [], []
for s in (8, 16, 32, 64, 128):
    h = 1024 // s
    cls.append((torch.randn(B, 32, h, h, generator=g) * 2 - 4).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)[:, :15])
    reg.append((torch.randn(B, 32, h, h, generator=g) * 0.3).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)[:, :5])
ref = fam_decode(reg, (8, 16, 32, 64, 128))
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print("fam_decode %.1f us | select_decode %.1f us" % (t(lambda: fam_decode(reg, (8, 16, 32, 64, 128))), t(lambda: select_decode(cls, reg, ref, 2000))))
