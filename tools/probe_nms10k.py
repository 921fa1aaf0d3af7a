"""The fused multiclass NMS on the bench's candidate sets (3 k and 10 k candidates per image, batch 8): per-kernel times
under `ncu --metrics gpu__time_duration.sum`, or CUDA-event totals when run bare."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from s2anet_b200.nms_rotated import multiclass_nms_rotated_batched
dev = torch.device("cuda", 0)
head = bench.build_head(torch, dev, torch.bfloat16, seed=0, bias=bench.CALIBRATED_BIAS[3000])
feats = bench.make_feats(torch, 8, 4, dev, torch.bfloat16)
for target in (3000, 10000):
    n = head.calibrate_scores(feats, target)
    bboxes, scores = head.select_and_decode(head.forward_levels(feats))
    per_class = (scores > 0.05).sum(dim=1)            # [B, C]
    print("target %d: %d candidates/image, largest class segment %d, mean %.0f" % (target, n, int(per_class.max()), float(per_class.float().mean())), flush=True)
    for _ in range(3):
        multiclass_nms_rotated_batched(bboxes, scores, 0.05, 0.5, 2000)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        multiclass_nms_rotated_batched(bboxes, scores, 0.05, 0.5, 2000)
    b.record(); torch.cuda.synchronize()
    print("  fused multiclass NMS: %.3f ms per call" % (a.elapsed_time(b) / 10), flush=True)
