"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the authoring container, where
/root/reference is mounted and oracle/_ref has been built by oracle/build_oracle.py).

    python tests/golden/make_golden.py

Every vector records which reference artefact produced it:
  * "ref_ext_cpu"   : the reference's unmodified torch extension (CPU kernels) built in place;
  * "ref_hdr_cuda"  : the reference's box_iou_rotated_utils.h compiled by nvcc as host code
                      (__CUDACC__ defined => the hull ordering its CUDA kernels execute);
  * "ref_py"        : the reference's Python (AlignConv.get_offset, ORConv2d.get_indices) imported
                      from /root/reference with its CUDA-extension imports stubbed;
  * "torchvision"   : torchvision.ops.deform_conv2d (CPU) -- the deform-conv CPU implementation
                      BASELINE.json names, since the reference has none (deform_conv.py:58-59).
The reference publishes no vectors of its own.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_oracle as bo      # noqa: E402
from oracle import oracle as O             # noqa: E402
from s2anet_b200 import synth              # noqa: E402

REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def greedy(iou, order, thr, labels=None, ge=False):
    """the reference's sweep (nms_rotated_cuda.cu:104-127) on a precomputed IoU matrix."""
    keep, dead = [], np.zeros(len(order), bool)
    for a, i in enumerate(order):
        if dead[a]:
            continue
        keep.append(i)
        for b in range(a + 1, len(order)):
            j = order[b]
            if dead[b] or (labels is not None and labels[i] != labels[j]):
                continue
            v = iou[i, j]
            if (v >= thr) if ge else (v > thr):
                dead[b] = True
    return np.asarray(keep, np.int64)


def import_reference_python():
    """models/alignconv.py and models/orn/modules/ORConv.py with the .so imports stubbed."""
    for name in ("models", "models.dcn", "models.init_weights", "models.orn", "models.orn.functions",
                 "models.orn.modules"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["models.dcn"].DeformConv = lambda *a, **k: torch.nn.Identity()
    sys.modules["models.init_weights"].normal_init = lambda *a, **k: None
    sys.modules["models.orn.functions"].active_rotating_filter = None
    import importlib.util

    def load(modname, path):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    ac = load("models.alignconv", "models/alignconv.py")
    orc = load("models.orn.modules.ORConv", "models/orn/modules/ORConv.py")
    return ac, orc


def main():
    bo.build_oracle()
    bo.build_ref_shims()
    bo.build_ref_extensions("cpu")
    iou_ext = bo.load_ref_extension("box_iou_rotated_cuda")
    nms_ext = bo.load_ref_extension("nms_rotated_cuda")
    ml_ext = bo.load_ref_extension("ml_nms_rotated_cuda")

    # ---- config 1: 2,000 clustered boxes -------------------------------------------------------
    boxes, scores, labels = synth.clustered_boxes(seed=0)
    full_cuda = O.ref_pairwise("iou", "cudasem", boxes, boxes)            # ref_hdr_cuda
    full_cpu = iou_ext.box_iou_rotated(torch.from_numpy(boxes[:200]), torch.from_numpy(boxes)).numpy()   # ref_ext_cpu
    order = np.argsort(-scores, kind="stable")
    keep_cuda = greedy(full_cuda, order, 0.5)
    keep_ml_cuda = greedy(full_cuda, order, 0.5, labels=labels)
    tb, ts, tl = torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(labels)
    keep_cpu = nms_ext.nms_rotated(tb, ts, 0.5).numpy()                   # ref_ext_cpu (>=, std::sort hull)
    keep_ml_cpu = ml_ext.ml_nms_rotated(tb, ts, tl, 0.5).numpy()
    np.savez_compressed(
        os.path.join(HERE, "cfg1_iou_nms.npz"), boxes=boxes, scores=scores, labels=labels,
        iou_rows_0_200_ref_hdr_cuda=full_cuda[:200], iou_rows_0_200_ref_ext_cpu=full_cpu,
        iou_full_sha256_ref_hdr_cuda=np.frombuffer(sha(full_cuda).encode(), np.uint8),
        iou_full_rowsum_ref_hdr_cuda=full_cuda.astype(np.float64).sum(1),
        keep_nms_thr05_cuda_semantics=keep_cuda, keep_mlnms_thr05_cuda_semantics=keep_ml_cuda,
        keep_nms_thr05_ref_ext_cpu=keep_cpu, keep_mlnms_thr05_ref_ext_cpu=keep_ml_cpu)
    print("cfg1: kept", len(keep_cuda), len(keep_ml_cuda), "cpu ext", len(keep_cpu), len(keep_ml_cpu))

    # ---- adversarial pairs -------------------------------------------------------------------
    adv = synth.adversarial_boxes()
    np.savez_compressed(os.path.join(HERE, "adversarial_iou.npz"), boxes=adv,
                        iou_ref_hdr_cuda=O.ref_pairwise("iou", "cudasem", adv, adv),
                        iou_ref_ext_cpu=iou_ext.box_iou_rotated(torch.from_numpy(adv), torch.from_numpy(adv)).numpy())

    # ---- anchors x GT (config 4, one image, every 16th anchor) -----------------------------------
    anchors = synth.all_level_anchors(1, 3)[0][::16]
    gts = synth.dota_like_gt(500, 3)
    np.savez_compressed(os.path.join(HERE, "cfg4_anchor_gt_iou.npz"), anchors=anchors, gts=gts,
                        iou_ref_hdr_cuda=O.ref_pairwise("iou", "cudasem", anchors, gts))

    # ---- AlignConv / deform conv / ORConv (small shapes) -----------------------------------------
    ac, orc = import_reference_python()
    import torchvision
    g = torch.Generator().manual_seed(1)
    B, C, H, W, Co, stride = 2, 16, 12, 10, 24, 8
    x = torch.randn(B, C, H, W, generator=g)
    anc = torch.from_numpy(synth.refined_anchors(B, H, W, stride, seed=1))
    w = torch.randn(Co, C, 3, 3, generator=g) * 0.05
    mod = ac.AlignConv.__new__(ac.AlignConv)
    torch.nn.Module.__init__(mod)
    mod.kernel_size = (3, 3)
    with torch.no_grad():
        import warnings
        warnings.simplefilter("ignore")
        off = torch.stack([mod.get_offset(anc[i].reshape(-1, 5), (H, W), stride) for i in range(B)])   # ref_py
    y = torch.relu(torchvision.ops.deform_conv2d(x, off, w, padding=1))                                  # torchvision
    # a generic deform conv: 5 in-groups... stride 2, dilation 2, pad 2, 2 deformable groups
    off2 = torch.randn(B, 2 * 2 * 9, 6, 5, generator=g) * 2.0
    y2 = torchvision.ops.deform_conv2d(x, off2, w, stride=2, padding=2, dilation=2)
    np.savez_compressed(os.path.join(HERE, "alignconv_small.npz"), x=x.numpy(), anchors=anc.numpy(), weight=w.numpy(),
                        stride=np.float32(stride), offset_ref_py=off.numpy(), out_torchvision=y.numpy(),
                        offset2=off2.numpy(), out2_torchvision=y2.numpy())

    conv = orc.ORConv2d.__new__(orc.ORConv2d)
    conv.nOrientation, conv.nRotation, conv.kernel_size = 1, 8, (3, 3)
    idx18 = orc.ORConv2d.get_indices(conv)                                                               # ref_py
    conv.nOrientation = 8
    idx88 = orc.ORConv2d.get_indices(conv)
    O_, I_ = 4, 6
    wo = torch.randn(O_, I_, 1, 3, 3, generator=g) * 0.1
    # ARF by literal index scatter (ActiveRotatingFilter_cuda.cu:36-43), on torch
    wr = torch.zeros(O_ * 8, I_, 3, 3)
    flat = wo.reshape(O_, I_, 9)
    for l in range(9):
        for k in range(8):
            dst = int(idx18.reshape(9, 8)[l, k]) - 1
            wr.reshape(O_, 8, I_, 9)[:, k, :, dst] = flat[:, :, l]
    xo = torch.randn(2, I_, 9, 7, generator=g)
    bias = torch.randn(O_ * 8, generator=g) * 0.1
    yo = torch.nn.functional.conv2d(xo, wr, bias, padding=1)
    yp = yo.view(2, -1, 8, 9, 7).max(dim=2)[0]
    np.savez_compressed(os.path.join(HERE, "orconv_small.npz"), indices_1_8=idx18.numpy(), indices_8_8=idx88.numpy(),
                        weight=wo.numpy(), rotated=wr.numpy(), x=xo.numpy(), bias=bias.numpy(), out=yo.numpy(),
                        pooled=yp.numpy())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
