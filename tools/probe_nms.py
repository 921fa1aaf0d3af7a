"""Time the fused multiclass NMS on the bench's candidate set (batch 8, 5,344 boxes x 15 classes per image)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from s2anet_b200.nms_rotated import multiclass_nms_rotated_batched
dev = torch.device("cuda", 0)
head = bench.build_head(torch, dev, torch.bfloat16, seed=0)
feats = bench.make_feats(torch, 8, 4, dev, torch.bfloat16)
n = head.calibrate_scores(feats, 3000)
outs = head.forward_levels(feats)
bboxes, scores = head.select_and_decode(outs)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
us = t(lambda: multiclass_nms_rotated_batched(bboxes, scores, 0.05, 0.5, 2000))
per_class = (scores > 0.05).sum(dim=1)
print("candidates/image %.0f, largest class segment %d, multiclass NMS %.1f us" % (n, int(per_class.max()), us))
