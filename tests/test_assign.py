"""Fused label assignment: s2a_assign_labels.

CPU part: the oracle's numpy restatement against golden vectors produced by the reference's own
assign_labels (models/utils.py:33-147) on the reference's CPU IoU kernel
(tests/golden/make_golden_assign.py).  GPU part: the two-pass sm_100a kernel against the golden vectors and
the oracle; results are integer labels, so the bar is exact equality."""
import numpy as np
import pytest
import torch

from s2anet_b200 import synth

DEV = "cuda:0"


def _cases(g):
    i = 0
    while "anchors_%d" % i in g:
        kw = {k[len("kw_%d_" % i):]: g[k].item() for k in g.files if k.startswith("kw_%d_" % i)}
        yield i, g["anchors_%d" % i], g["gts_%d" % i], tuple(int(v) for v in g["size_%d" % i]), kw, g["assign_%d" % i]
        i += 1


def test_oracle_matches_reference_assign_labels(oracle, golden):
    g = golden("assign_small.npz")
    n = 0
    for i, a, gt, size, kw, want in _cases(g):
        # on the reference's own IoU matrix: the assignment logic alone
        got = oracle.assign_labels(a, gt, imgs_size=size, ious=g["iou_ref_ext_cpu_%d" % i] if gt.shape[0] else None, **kw)
        assert np.array_equal(got, want), i
        # and on the oracle's IoU (CUDA-build semantics): identical here, no degenerate pair in these sets
        assert np.array_equal(oracle.assign_labels(a, gt, imgs_size=size, **kw), want), i
        n += 1
    assert n == 4
    # the special cases really are in the set: anchor 10 lies outside the image (ignored), so the GT that was copied
    # from it -- present twice -- goes to its best valid anchor, and the LATER duplicate (index 1) wins
    a1 = g["assign_1"]
    assert a1[10] == -2 and a1[13] == 1 and a1[500] == 2 and a1[777] == 3


@pytest.mark.gpu
def test_kernel_matches_reference_golden(golden):
    from s2anet_b200.assign import assign_labels
    g = golden("assign_small.npz")
    for i, a, gt, size, kw, want in _cases(g):
        got = assign_labels(torch.from_numpy(a).to(DEV), torch.from_numpy(gt).to(DEV), imgs_size=size, **kw)
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want), i


@pytest.mark.gpu
def test_batched_ragged_against_oracle_and_materialised_matrix(oracle):
    """Full anchor set of a 1024^2 image (21,824) x up to 300 GTs, 4 images with different GT counts; compared
    with the oracle and with the reference's formulation run in PyTorch on this library's IoU matrix."""
    from s2anet_b200.assign import assign_labels_batched
    from s2anet_b200.box_iou_rotated import box_iou_rotated
    B, counts = 4, [300, 1, 0, 137]
    anchors = synth.all_level_anchors(B, 21)
    gts = np.zeros((B, 300, 5), np.float32)
    for b, c in enumerate(counts):
        if c:
            gts[b, :c] = synth.dota_like_gt(c, 40 + b)
    gts[0, 7] = anchors[0, 1234]            # exact hit and a duplicated GT
    gts[0, 9] = anchors[0, 1234]
    gts[3, 5, 2:4] = 0.0                    # zero-area GT: IoU 0 with everything
    ta, tg = torch.from_numpy(anchors).to(DEV), torch.from_numpy(gts).to(DEV)
    tc = torch.tensor(counts, dtype=torch.int32, device=DEV)
    got = assign_labels_batched(ta, tg, tc).cpu().numpy()
    for b, c in enumerate(counts):
        want = oracle.assign_labels(anchors[b], gts[b, :c]) if b in (1, 2) else None
        ious = box_iou_rotated(ta[b], tg[b, :c]).cpu().numpy() if c else None
        ref = oracle.assign_labels(anchors[b], gts[b, :c], ious=ious)
        assert np.array_equal(got[b], ref), b
        if want is not None:
            assert np.array_equal(got[b], want), b
    assert got[0, 1234] == 9 and (got[2] >= 0).sum() == 0 and (got[0] >= 0).sum() >= 250
    # padding rows beyond gt_counts never matter
    tg2 = tg.clone()
    tg2[1, 1:] = torch.from_numpy(synth.dota_like_gt(299, 99)).to(DEV)
    assert torch.equal(assign_labels_batched(ta, tg2, tc), torch.from_numpy(got).to(DEV))
    # gt_max_assign_all=False and other thresholds
    got2 = assign_labels_batched(ta[:1], tg[:1], tc[:1], pos_iou_thr=0.6, neg_iou_thr=0.3, min_pos_iou_thr=0.05,
                                 gt_max_assign_all=False, filter_invalid_anchors=False).cpu().numpy()
    ious = box_iou_rotated(ta[0], tg[0, :300]).cpu().numpy()
    ref2 = oracle.assign_labels(anchors[0], gts[0, :300], pos_iou_thr=0.6, neg_iou_thr=0.3, min_pos_iou_thr=0.05,
                                gt_max_assign_all=False, filter_invalid_anchors=False, ious=ious)
    assert np.array_equal(got2[0], ref2)


@pytest.mark.gpu
def test_argument_checks():
    from s2anet_b200.assign import assign_labels
    a, g = torch.rand(10, 5, device=DEV) * 50 + 1, torch.rand(3, 5, device=DEV) * 50 + 1
    with pytest.raises(RuntimeError):
        assign_labels(a, g, min_pos_iou_thr=-0.1)
    with pytest.raises(NotImplementedError):
        assign_labels(a.cpu(), g.cpu())
    assert assign_labels(a[:0], g).shape == (0,)
