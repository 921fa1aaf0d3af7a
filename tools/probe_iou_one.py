"""Config-4-shaped batched IoU (16 images x 21,824 x 500) for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.box_iou_rotated import box_iou_rotated_batched
dev = "cuda:0"
B = 16
an = torch.from_numpy(synth.all_level_anchors(B, 3)).to(dev)
gt = torch.from_numpy(np.stack([synth.dota_like_gt(500, 100 + i) for i in range(B)])).to(dev)
out = torch.empty((B, an.shape[1], 500), device=dev)
for _ in range(3):
    box_iou_rotated_batched(an, gt, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
