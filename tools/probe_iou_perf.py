"""Rotated IoU pairs/s at BASELINE config 4 shape (16 images x 21,824 x 500) and a square 8,000^2 self-matrix."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.box_iou_rotated import box_iou_rotated, box_iou_rotated_batched
dev = "cuda:0"
B = 16
an = torch.from_numpy(synth.all_level_anchors(B, 3)).to(dev)
gt = torch.from_numpy(np.stack([synth.dota_like_gt(500, 100 + i) for i in range(B)])).to(dev)
out = torch.empty((B, an.shape[1], 500), device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: box_iou_rotated_batched(an, gt, out=out))
print("anchor x GT: %.3f ms, %.1f G pairs/s" % (ms, B * an.shape[1] * 500 / ms / 1e6))
bx, _, _ = synth.clustered_boxes(n_seed=1600, rep=5, seed=0)
tb = torch.from_numpy(bx).to(dev)
ms = t(lambda: box_iou_rotated(tb, tb))
print("clustered %d^2: %.3f ms, %.1f G pairs/s" % (tb.shape[0], ms, tb.shape[0] ** 2 / ms / 1e6))
