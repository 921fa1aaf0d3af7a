"""Per-rank kernel time of the sharded anchor x GT IoU (config 4) emulated on ONE GPU: rank r of `world` computes its
cyclically dealt 32-row tiles (box_iou_rotated_tiles, compact)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from s2anet_b200 import synth
from s2anet_b200.box_iou_rotated import box_iou_rotated_tiles
dev = "cuda:0"
B = 64
anc = torch.from_numpy(synth.all_level_anchors(B, 3)).to(dev)
gt = torch.from_numpy(np.stack([synth.dota_like_gt(500, 100 + i) for i in range(B)])).to(dev)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for world in (1, 2, 4, 8):
    ms = [t(lambda: box_iou_rotated_tiles(anc, gt, r, world, compact=True, tile_rows=32)) for r in range(world)]
    print("world %d per-rank ms: max %.3f min %.3f mean %.3f sum %.3f" % (world, max(ms), min(ms), sum(ms) / len(ms), sum(ms)), flush=True)
