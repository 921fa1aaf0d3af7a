"""Fused label assignment for rotated anchors.

`assign_labels` keeps the reference's name, argument order and return value (models/utils.py:33-147):
anchors [M,5], gt_boxes [N,5] -> assign_gt_ids [M] int64 (-2 ignored, -1 negative, >= 0 GT index).
`assign_labels_batched` does a whole batch (ragged GT counts) in the same two launches.  The IoU
matrix is never materialised.
"""
import torch

from . import _lib


def assign_labels_batched(anchors, gt_boxes, gt_counts=None, imgs_size=(1024, 1024), pos_iou_thr=0.5, neg_iou_thr=0.4,
                          min_pos_iou_thr=0, gt_max_assign_all=True, filter_invalid_anchors=True):
    """anchors [B,M,5], gt_boxes [B,Nmax,5] (rows >= gt_counts[b] are padding), gt_counts [B] int or None."""
    dev = _lib.require_cuda(anchors, gt_boxes, gt_counts)
    if anchors.dim() != 3 or anchors.size(2) != 5 or gt_boxes.dim() != 3 or gt_boxes.size(2) != 5:
        raise ValueError("assign_labels: expected anchors [B,M,5] and gt_boxes [B,N,5]")
    if gt_boxes.size(0) != anchors.size(0):
        raise ValueError("assign_labels: batch sizes differ")
    a = anchors.detach().to(torch.float32).contiguous()
    g = gt_boxes.detach().to(torch.float32).contiguous()
    c = None if gt_counts is None else gt_counts.to(torch.int32).contiguous()
    B, M, N = a.size(0), a.size(1), g.size(1)
    out = torch.empty((B, M), dtype=torch.int64, device=dev)
    lib = _lib.load()
    wbytes = lib.s2a_assign_labels_workspace_bytes(B, M, N)
    work = torch.empty((wbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.s2a_assign_labels(_lib.ptr(a), _lib.ptr(g), _lib.ptr(c), B, M, N, float(imgs_size[0]), float(imgs_size[1]),
                                   float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou_thr),
                                   1 if gt_max_assign_all else 0, 1 if filter_invalid_anchors else 0, _lib.ptr(out),
                                   _lib.ptr(work), wbytes, _lib.stream_ptr(dev))
    _lib.check(rc, "assign_labels")
    return out


def assign_labels(anchors, gt_boxes, imgs_size=(1024, 1024), pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou_thr=0,
                  gt_max_assign_all=True, filter_invalid_anchors=True, filter_invalid_ious=True):
    """Drop-in for models/utils.py:33 (one image).  `filter_invalid_ious` is always on: an IoU outside [0, 1]
    can be neither a positive nor a negative match (models/utils.py:89-96)."""
    if not filter_invalid_ious:
        raise NotImplementedError("assign_labels: filter_invalid_ious=False is not supported")
    return assign_labels_batched(anchors[None], gt_boxes[None], None, imgs_size, pos_iou_thr, neg_iou_thr, min_pos_iou_thr,
                                 gt_max_assign_all, filter_invalid_anchors)[0]
