"""Rotated IoU pairs/s at BASELINE config 4 (64 images x 21,824 x 500), per tile height, with and without the reject
tests, plus a square clustered self-matrix and the 8-way cyclic / contiguous shard times (load balance)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import synth
from s2anet_b200.box_iou_rotated import box_iou_rotated, box_iou_rotated_batched, box_iou_rotated_tiles
dev = "cuda:0"
B = int(os.environ.get("IOU_B", "64"))
an = torch.from_numpy(synth.all_level_anchors(B, 3)).to(dev)
gt = torch.from_numpy(np.stack([synth.dota_like_gt(500, 100 + i) for i in range(B)])).to(dev)
N = an.shape[1]
out = torch.empty((B, N, 500), device=dev)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
pairs = B * N * 500
for tr in (256, 128, 64):
    ms = t(lambda: box_iou_rotated_tiles(an, gt, 0, 1, compact=False, out=out, tile_rows=tr))
    print("anchor x GT, tile_rows %3d: %.3f ms, %.1f G pairs/s" % (tr, ms, pairs / ms / 1e6))
ms = t(lambda: box_iou_rotated_batched(an, gt, out=out, _flags=1), reps=2)
print("anchor x GT, no reject (every pair clipped): %.3f ms, %.1f G pairs/s" % (ms, pairs / ms / 1e6))
for w in (8,):
    cyc = [t(lambda r=r: box_iou_rotated_tiles(an, gt, r, w, compact=True)) for r in range(w)]
    cyc256 = [t(lambda r=r: box_iou_rotated_tiles(an, gt, r, w, compact=True, tile_rows=256)) for r in range(w)]
    per = -(-N // w); per = -(-per // 64) * 64
    blk = [t(lambda r=r: box_iou_rotated_batched(an, gt, min(N, r * per), min(N, (r + 1) * per), out=out)) for r in range(w)]
    print("8-way shards, ms per rank: cyclic 256-row tiles", " ".join("%.3f" % x for x in cyc256))
    print("8-way shards, ms per rank: cyclic 64-row tiles", " ".join("%.3f" % x for x in cyc), "| contiguous blocks", " ".join("%.3f" % x for x in blk))
bx, _, _ = synth.clustered_boxes(n_seed=1600, rep=5, seed=0)
tb = torch.from_numpy(bx).to(dev)
ms = t(lambda: box_iou_rotated(tb, tb))
print("clustered %d^2: %.3f ms, %.1f G pairs/s" % (tb.shape[0], ms, tb.shape[0] ** 2 / ms / 1e6))
