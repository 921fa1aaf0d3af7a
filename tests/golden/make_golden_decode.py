"""Golden vectors for the box-decode stages, produced by the REFERENCE's own Python (run in the
authoring container):

    python tests/golden/make_golden_decode.py

  * AnchorGeneratorRotated.gen_grid_anchors  (models/anchors.py:75-126)
  * fam_bbox_decode                          (models/head.py:27-52)
  * S2ANetHead.get_bboxes_single_img         (models/head.py:684-717), with the final
    multiclass_nms_rotated call intercepted so the (bboxes, scores) it would receive are recorded.
All imported unmodified from /root/reference on top of s2anet_b200.dropin's extension-module shims.
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
    from s2anet_b200 import dropin
    dropin.install()
    sys.path.insert(0, REF)
    import models.head as ref_head
    from models.anchors import AnchorGeneratorRotated

    g = torch.Generator().manual_seed(7)
    B, C, topk = 2, 15, 50
    strides = [8, 16, 32]
    sizes = [(12, 10), (6, 5), (3, 3)]
    head = ref_head.S2ANetHead(num_classes=C)
    head.max_before_nms_per_level = topk
    captured = []
    ref_head.multiclass_nms_rotated = lambda bboxes, scores, **kw: (captured.append((bboxes, scores)) or (bboxes, scores))

    out = {"strides": np.asarray(strides, np.int32), "sizes": np.asarray(sizes, np.int32), "topk": np.int32(topk)}
    fam, cls, reg, refine = [], [], [], []
    for l, ((H, W), s) in enumerate(zip(sizes, strides)):
        gen = AnchorGeneratorRotated(s, [4.0], [1.0], angles=[0, ])
        anchors = gen.gen_grid_anchors((H, W), s).reshape(-1, 5)
        fam_pred = torch.randn(B, 5, H, W, generator=g) * torch.tensor([0.3, 0.3, 0.6, 0.6, 0.4]).view(1, 5, 1, 1)
        fam_pred[0, 2, 0, 0] = 20.0            # exercises the wh_ratio_clip=1e-6 clamp (max_ratio 13.8)
        fam_pred[0, 3, 0, 1] = -20.0
        rf = ref_head.fam_bbox_decode(fam_pred, anchors)                       # [B,H,W,5]
        rf16 = ref_head.fam_bbox_decode(fam_pred.half(), anchors)              # the fp16 validation path (A.7)
        cl = torch.randn(B, C, H, W, generator=g) * 2.0 - 2.0
        rg = torch.randn(B, 5, H, W, generator=g) * torch.tensor([0.3, 0.3, 0.6, 0.6, 0.4]).view(1, 5, 1, 1)
        rg[1, 2, 0, 0] = 9.0                   # default clip 16/1000 -> max_ratio 4.135
        out["grid_anchors_%d" % l] = anchors.numpy()
        out["fam_pred_%d" % l] = fam_pred.numpy()
        out["refine_%d" % l] = rf.numpy()
        out["refine_f16in_%d" % l] = rf16.numpy()
        out["cls_%d" % l] = cl.numpy()
        out["reg_%d" % l] = rg.numpy()
        fam.append(fam_pred); cls.append(cl); reg.append(rg); refine.append(rf)
    for b in range(B):
        sc = [c[b].permute(1, 2, 0).reshape(-1, C) for c in cls]
        bp = [r[b].permute(1, 2, 0).reshape(-1, 5) for r in reg]
        an = [r[b].reshape(-1, 5) for r in refine]
        head.get_bboxes_single_img(sc, bp, an)
        out["bboxes_%d" % b] = captured[-1][0].numpy()
        out["scores_%d" % b] = captured[-1][1].numpy()
        # fp16 network outputs, fp32 anchors (val.py half mode); no top-k level so the order is defined
        head.max_before_nms_per_level = 0
        head.get_bboxes_single_img([s_.half() for s_ in sc], [b_.half() for b_ in bp], an)
        head.max_before_nms_per_level = topk
        out["bboxes_f16in_%d" % b] = captured[-1][0].float().numpy()
        out["scores_f16in_%d" % b] = captured[-1][1].float().numpy()
    np.savez_compressed(os.path.join(HERE, "decode_small.npz"), **out)
    print("decode_small.npz written:", {k: v.shape for k, v in out.items() if k.startswith(("bboxes", "refine_0"))})


if __name__ == "__main__":
    main()
