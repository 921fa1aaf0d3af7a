"""Golden vectors for label assignment, produced by the REFERENCE's own assign_labels
(models/utils.py:33-147, imported unmodified from /root/reference) running on the reference's own
box_iou_rotated CPU extension (oracle/_ref/ext_cpu, built in place from the reference sources):

    python tests/golden/make_golden_assign.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def main():
    from oracle import build_oracle as bo
    from s2anet_b200 import dropin, synth
    dropin.install()
    sys.path.insert(0, REF)
    import models.utils as mu
    ext = bo.load_ref_extension("box_iou_rotated_cuda", "cpu")
    captured = {}

    def iou_cpu(a, b):                       # utils/metrics.py:85-107 on the reference's CPU kernel
        captured["iou"] = ext.box_iou_rotated(a[..., :5].float().contiguous(), b[..., :5].float().contiguous())
        return captured["iou"]
    mu.bbox_iou_rotated = iou_cpu

    out = {}
    cases = []
    # case 0: all-level anchors of one 256x256 image vs 40 DOTA-like GTs (some anchors invalid: w >= image)
    anchors = synth.all_level_anchors(1, 3)[0]
    anchors = anchors[(anchors[:, 0] < 256) & (anchors[:, 1] < 256)]
    gts = synth.dota_like_gt(40, 5)
    gts[:, :2] = gts[:, :2] / 4.0
    cases.append((anchors, gts, (256, 256), dict()))
    # case 1: GTs that coincide with anchors (exact ties for the per-GT maximum), duplicated GTs, a far-away GT
    a1 = synth.all_level_anchors(1, 9)[0][:3000].copy()
    g1 = np.concatenate([a1[[10, 10, 500, 777]], np.array([[5000., 5000., 30., 10., 0.3]], np.float32),
                         synth.dota_like_gt(12, 9)]).astype(np.float32)
    cases.append((a1, g1, (1024, 1024), dict()))
    # case 2: other thresholds, gt_max_assign_all=False, no invalid-anchor filter
    cases.append((a1[:1500], g1, (1024, 1024), dict(pos_iou_thr=0.6, neg_iou_thr=0.3, min_pos_iou_thr=0.1,
                                                     gt_max_assign_all=False, filter_invalid_anchors=False)))
    # case 3: no GT at all
    cases.append((a1[:200], np.zeros((0, 5), np.float32), (1024, 1024), dict()))
    for i, (a, g, size, kw) in enumerate(cases):
        captured.clear()
        res = mu.assign_labels(torch.from_numpy(a), torch.from_numpy(g), imgs_size=size, **kw)
        out["anchors_%d" % i] = a
        out["gts_%d" % i] = g
        out["size_%d" % i] = np.asarray(size, np.int32)
        out["assign_%d" % i] = res.numpy()
        out["iou_ref_ext_cpu_%d" % i] = captured["iou"].numpy() if "iou" in captured else np.zeros((a.shape[0], 0), np.float32)
        for k, v in kw.items():
            out["kw_%d_%s" % (i, k)] = np.asarray(v)
        r = res.numpy()
        print("case", i, a.shape, g.shape, "pos", int((r >= 0).sum()), "neg", int((r == -1).sum()), "ign", int((r == -2).sum()))
    np.savez_compressed(os.path.join(HERE, "assign_small.npz"), **out)


if __name__ == "__main__":
    main()
