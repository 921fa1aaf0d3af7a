"""DOTA result-merging NMS on the GPU (fp64 polygons), with the reference's Python surface.

reference: DOTA_devkit/ResultMerge_multi_process.py:62-123 (py_cpu_nms_poly_fast) and the SWIG module it calls,
DOTA_devkit/polyiou (csrc/polyiou.cpp:110-126, iou_poly).

  poly_nms(dets, thresh)                 dets [N, 9] float64 CUDA tensor (8 polygon coordinates + score) -> keep int64
  py_cpu_nms_poly_fast(dets, thresh)     the reference's name and contract: numpy [N, 9] in, Python list of kept row
                                         indices out (descending score) -- the rows go to the GPU and back
  iou_poly(p, q)                         polyiou.iou_poly for one pair of 8-number polygons (sequence or VectorDouble)
  VectorDouble                           stand-in for polyiou.VectorDouble (a list of floats)
  iou_poly_pairs(p, q)                   [n, 8] x [n, 8] float64 CUDA tensors -> [n] IoUs (parity probe)

Tie rule (stated, because the reference's is not): the reference orders with `scores.argsort()[::-1]`
(ResultMerge_multi_process.py:81) -- numpy's default introsort is not stable, so which of two EQUAL scores comes
first is unspecified there (for short arrays the insertion-sort path happens to put the higher index first).  Here
the order is a stable descending sort: among equal scores the LOWER row index is visited first, as in nms_rotated.
With distinct scores (the goldens, and every test of this repo) the keep lists are identical; with duplicated scores
-- possible after merging patches -- the kept SET can differ from one particular numpy build's by which duplicate
survives.  NaN IoUs (0/0 of two zero-area polygons) suppress, as `np.where(hbb_ovr <= thresh)` makes them do.
"""
import numpy as np
import torch

from . import _lib


def poly_nms(dets, thresh=0.5):
    dev = _lib.require_cuda(dets)
    if dets.dim() != 2 or dets.size(1) < 9:
        raise ValueError("dets must be [N, 9]: x0 y0 x1 y1 x2 y2 x3 y3 score")
    n = dets.size(0)
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    if dets.dtype != torch.float64 or dets.stride(1) != 1:
        dets = dets.to(torch.float64).contiguous()
    lib = _lib.load()
    ws_bytes = lib.s2a_poly_nms_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    keep = torch.empty((n,), dtype=torch.int64, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.s2a_poly_nms(_lib.ptr(dets), dets.stride(0), n, float(thresh), _lib.ptr(keep), _lib.ptr(num),
                              _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
    _lib.check(rc, "poly_nms")
    return keep[: int(num.item())]


def iou_poly_pairs(p, q):
    dev = _lib.require_cuda(p, q)
    if p.shape != q.shape or p.dim() != 2 or p.size(1) != 8:
        raise ValueError("iou_poly_pairs: p and q must both be [n, 8]")
    p = p.to(torch.float64).contiguous()
    q = q.to(torch.float64).contiguous()
    out = torch.empty((p.size(0),), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_poly_iou_pairs(_lib.ptr(p), _lib.ptr(q), p.size(0), _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(rc, "poly_iou_pairs")
    return out


class VectorDouble(list):
    """polyiou.VectorDouble: the reference builds one per detection before calling iou_poly."""

    def __init__(self, values=()):
        super(VectorDouble, self).__init__(float(v) for v in values)


def iou_poly(p, q, device="cuda"):
    """polyiou.iou_poly(p, q) -> float."""
    tp = torch.tensor([list(p)], dtype=torch.float64, device=device)
    tq = torch.tensor([list(q)], dtype=torch.float64, device=device)
    return float(iou_poly_pairs(tp, tq)[0])


def py_cpu_nms_poly_fast(dets, thresh=0.5, device="cuda"):
    """Drop-in for ResultMerge_multi_process.py:62-123: numpy [N, 9] -> list of kept indices."""
    dets = np.asarray(dets, dtype=np.float64)
    if dets.shape[0] == 0:
        return []
    keep = poly_nms(torch.from_numpy(np.ascontiguousarray(dets)).to(device), thresh)
    return [int(i) for i in keep.cpu().tolist()]
