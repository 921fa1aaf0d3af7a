"""The reference's UNMODIFIED Python callers on the B200 ops, on the GPU (-m gpu).

north_star: "the Python operator surface stays unchanged so it drops into models/detector.py and val.py".  These
tests import the reference's own `models/head.py` (forward / forward_single :262-348, get_bboxes /
get_bboxes_single_img :648-725), `models/detector.py` (:28-37) and `utils/bbox_nms_rotated.py` from the staged
checkout (baseline/_ref on the GPU box, /root/reference in the authoring container; never edited) on top of the
shim modules of `s2anet_b200.dropin.install()`, and run them

  * in fp32 and in fp16 (val.py:126,196,246 runs half precision by default: `model.half()`, `imgs.half()`),
  * with `install()` alone (the reference's own AlignConv / DeformConvFunction / ORConv2d classes calling the
    extension-level entries `deform_conv_forward_cuda`, `arf_forward`, `ml_nms_rotated`) and after
    `accelerate(model)` (the fused kernels swapped in),

against `s2anet_b200.head.S2ANetHead` loaded with the same state dict.

Stated tolerances.  Raw head outputs: fp32 max-abs <= 2e-3 * max|ref| (the stock conv towers run cuDNN with different
algorithms / TF32 settings in the two callers; the custom ops themselves are held to 1e-4 in test_gpu_conv.py);
fp16: relative L2 <= 3e-2 against the fp32 run of the same head (eight stacked half-precision layers).  Post-processing from IDENTICAL head outputs:
same number of detections, identical labels, coordinates within 1e-3 px, scores within 1e-6.
"""
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
IMG = 512
STRIDES = (8, 16, 32, 64, 128)


@pytest.fixture(scope="module")
def ref():
    """Reference tree on sys.path + shims installed; yields the reference's `models.head` module."""
    from oracle import build_oracle
    root = build_oracle.reference_root()
    if root is None:
        pytest.skip("reference tree neither mounted (/root/reference) nor staged (baseline/_ref)")
    saved = dict(sys.modules)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # both callers in plain fp32
    sys.path.insert(0, root)
    from s2anet_b200 import dropin
    dropin.install()
    import models.head as ref_head
    assert ref_head.__file__.startswith(root)
    yield ref_head
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    sys.path.remove(root)
    for k in list(sys.modules):
        if k not in saved and k.split(".")[0] in ("models", "utils", "matplotlib", "DOTA_devkit"):
            del sys.modules[k]


def make_feats(batch, dtype, seed=7):
    """FPN-like maps: spatially smooth (low-resolution noise, bilinearly upsampled 4x) rather than white noise, whose
    pixel-to-pixel jumps would turn the 1/32-pixel position rounding of ANY half-precision deformable conv into
    percent-level output noise and make the fp16-vs-fp32 comparison below meaningless."""
    g = torch.Generator().manual_seed(seed)
    feats = []
    for s in STRIDES:
        h = IMG // s
        if h >= 8:
            t = torch.nn.functional.interpolate(torch.randn(batch, 256, h // 4, h // 4, generator=g), size=(h, h),
                                                mode="bilinear", align_corners=False) * 1.5
        else:
            t = torch.randn(batch, 256, h, h, generator=g)
        feats.append(t.to(DEV).to(dtype))
    return feats


@pytest.fixture(scope="module")
def weights():
    """A synthetic, calibrated state dict (rotated refined anchors, a few thousand scores above 0.05)."""
    from s2anet_b200.head import S2ANetHead
    mine = S2ANetHead(15).eval()
    mine.init_synthetic(3)
    mine = mine.to(DEV)
    n = mine.calibrate_scores(make_feats(2, torch.float32), target_candidates=1500)
    assert n > 200
    return {k: v.detach().clone() for k, v in mine.state_dict().items()}


def build_pair(ref_head, weights, dtype, accelerate):
    from s2anet_b200 import dropin
    from s2anet_b200.head import S2ANetHead
    theirs = ref_head.S2ANetHead(num_classes=15).eval().to(DEV)
    theirs.load_state_dict(weights)
    mine = S2ANetHead(15).eval().to(DEV)
    mine.load_state_dict(weights)
    if dtype != torch.float32:
        theirs, mine = theirs.to(dtype), mine.to(dtype)
    if accelerate:
        assert dropin.accelerate(theirs) == 3
    else:
        assert type(theirs.align_conv).__module__ == "models.alignconv"          # the reference's own classes
        assert type(theirs.or_conv).__module__.startswith("models.orn")
    return theirs, mine


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("accelerate", [False, True], ids=["install-only", "accelerated"])
def test_reference_head_fp32(ref, weights, accelerate):
    theirs, mine = build_pair(ref, weights, torch.float32, accelerate)
    feats = make_feats(2, torch.float32)
    before = _launches()
    with torch.no_grad():
        res = theirs([f.clone() for f in feats], post_process=False)
        p = res["pred"]                       # (fam_cls, fam_reg, odm_cls, odm_reg, init_anchors, refine) x levels
        outs = mine.forward_levels(feats)
    assert _launches() > before               # the B200 library really ran underneath the reference code
    for name, k in (("fam_cls", 0), ("fam_reg", 1), ("odm_cls", 2), ("odm_reg", 3), ("refine", 5)):
        for l in range(len(STRIDES)):
            a, b = p[k][l].float(), outs[l][k].float()
            tol = 2e-3 * float(b.abs().max()) + 1e-5
            assert float((a - b).abs().max()) <= tol, (name, l, float((a - b).abs().max()), tol)
    with torch.no_grad():
        dets_ref = theirs.get_bboxes(p)                                          # the reference's own post-processing
        mine_outs = [(p[0][l], p[1][l], p[2][l], p[3][l], p[4][l], p[5][l]) for l in range(len(STRIDES))]
        bboxes, scores = mine.select_and_decode(mine_outs)
        from s2anet_b200.nms_rotated import multiclass_nms_rotated_batched
        d, lb, cnt = multiclass_nms_rotated_batched(bboxes, scores, mine.score_thres_before_nms, mine.iou_thres_nms,
                                                    mine.max_per_img)
    total = 0
    for i, (db, dl) in enumerate(dets_ref):
        k = int(cnt[i])
        assert db.shape[0] == k, (db.shape, k)
        total += k
        if k:
            assert torch.equal(dl.reshape(-1).float(), lb[i, :k])
            assert float((db[:, :5] - d[i, :k, :5]).abs().max()) <= 1e-3
            assert float((db[:, 5] - d[i, :k, 5]).abs().max()) <= 1e-6
    assert total > 50                          # NMS really had work to do


@pytest.mark.parametrize("accelerate", [False, True], ids=["install-only", "accelerated"])
def test_reference_head_fp16(ref, weights, accelerate):
    """val.py's default: the whole model in half precision.  `install()` alone must be enough: the
    reference's DeformConvFunction casts offset/weight to half and calls deform_conv_forward_cuda with half tensors."""
    theirs16, mine16 = build_pair(ref, weights, torch.float16, accelerate)
    theirs32, _ = build_pair(ref, weights, torch.float32, accelerate)
    f32 = make_feats(2, torch.float32)
    f16 = [f.half() for f in f32]
    with torch.no_grad():
        p16 = theirs16([f.clone() for f in f16], post_process=False)["pred"]
        p32 = theirs32([f.half().float() for f in f32], post_process=False)["pred"]
        outs16 = mine16.forward_levels(f16)
    for k in (1, 2, 3):
        for l in range(len(STRIDES)):
            assert p16[k][l].dtype == torch.float16 and bool(torch.isfinite(p16[k][l]).all())
            assert rel_l2(p16[k][l], p32[k][l]) <= 3e-2, (k, l, rel_l2(p16[k][l], p32[k][l]))
            assert rel_l2(outs16[l][k], p32[k][l]) <= 3e-2, (k, l, rel_l2(outs16[l][k], p32[k][l]))
    with torch.no_grad():
        res = theirs16([f.clone() for f in f16], post_process=True)              # forward + get_bboxes, as val.py calls it
    dets = res["boxes_ls"]
    assert len(dets) == 2 and sum(d.shape[0] for d, _ in dets) > 50
    for d, l in dets:
        assert d.shape[1] == 6 and bool(torch.isfinite(d.float()).all()) and l.shape[0] == d.shape[0]
        assert bool((d[1:, 5] <= d[:-1, 5]).all())                               # score-sorted, as the reference returns them
    # the same post-processing through this repo's batched path, from the same half-precision predictions
    mine_outs = [tuple(res_l) for res_l in zip(*[p16[i] for i in range(6)])]
    with torch.no_grad():
        got = mine16.get_bboxes_from_outs(mine_outs)
    for (d, l), (gd, gl) in zip(dets, got):
        assert d.shape[0] == gd.shape[0]
        assert torch.equal(l.reshape(-1).float(), gl.reshape(-1).float())
        assert float((d.float() - gd.float()).abs().max()) <= 2e-2


def test_reference_detector_half_precision_end_to_end(ref, monkeypatch):
    """models/detector.py:28-37 unmodified: ResNet-50 + FPN (stock PyTorch, random init: there is no network for the
    checkpoint) + the head on the shims, `model.half()` + `imgs.half()` + post_process=True like val.py:196,246."""
    import models.backbone as bb
    monkeypatch.setattr(bb, "load_checkpoint", lambda name: {})                  # no download
    monkeypatch.setattr(bb, "load_state_dict", lambda model, sd: model)          # random init
    from models.detector import S2ANet
    torch.manual_seed(0)
    model = S2ANet(backbone_name="resnet50", num_classes=15).to(DEV).eval().half()
    with torch.no_grad():
        torch.nn.init.constant_(model.head.odm_cls_head.bias, -2.0)              # a few hundred scores above 0.05
    imgs = torch.rand(2, 3, 256, 256, device=DEV).half()
    before = _launches()
    with torch.no_grad():
        res = model(imgs, post_process=True)
    assert _launches() > before
    assert len(res["boxes_ls"]) == 2
    for d, l in res["boxes_ls"]:
        assert d.dim() == 2 and d.shape[1] == 6 and d.shape[0] <= 2000 and bool(torch.isfinite(d.float()).all())


def test_reference_training_step_backward_through_the_shims(ref, weights):
    """train.py's path: the reference head in training mode, loss-free proxy (sum of the ODM outputs), backward through
    the reference's DeformConvFunction / _ActiveRotatingFilter autograd Functions on the shim entries; and the same
    through this repo's classes after accelerate() (ORConv2d must not cut the graph)."""
    from s2anet_b200 import dropin
    feats = make_feats(1, torch.float32, seed=11)
    grads = []
    for accelerate in (False, True):
        theirs, _ = build_pair(ref, weights, torch.float32, accelerate)
        theirs.train()
        p = theirs([f.clone().requires_grad_(True) for f in feats], post_process=False)["pred"]
        loss = sum(o.float().square().mean() for o in p[2]) + sum(o.float().square().mean() for o in p[3])
        loss.backward()
        g_or = theirs.or_conv.weight.grad
        g_al = theirs.align_conv.deform_conv.weight.grad
        assert g_or is not None and g_al is not None and float(g_or.abs().sum()) > 0 and float(g_al.abs().sum()) > 0
        grads.append((g_or.clone(), g_al.clone()))
    for a, b in zip(grads[0], grads[1]):
        assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max()) + 1e-7


def _launches():
    from s2anet_b200 import _lib
    return _lib.launches
