"""Box decode stages of the S2ANet head as single multi-level launches.

  fam_decode     <- fam_bbox_decode + gen_grid_anchors (models/head.py:27-52, models/anchors.py:75-126)
  select_decode  <- get_bboxes_single_img up to the NMS (models/head.py:684-717), batched over images

plus `rboxes_decode_torch` / `select_and_decode_torch`, the reference's own PyTorch formulation
(models/boxes.py:82-162, utils/general.py:925-930) kept here as the parity witness for the tests
(it is NOT a fallback: the ops below refuse CPU tensors).
"""
import ctypes as C
import math

import torch

from . import _lib


def norm_angle(angle):
    """utils/general.py:925-930: wrap into [-pi/4, 3pi/4)."""
    lo = -math.pi / 4
    return (angle - lo) % math.pi + lo


def rboxes_decode_torch(anchors, deltas, wh_ratio_clip=16 / 1000):
    """models/boxes.py:82-162 (delta2bbox_rotated, is_encode_relative=True), line by line, with the
    reference's dtype behaviour: fp32 anchors, deltas in the network dtype (half tensors stay half
    through clamp / exp / pi*dangle and are promoted where they meet an anchor value)."""
    dx, dy, dw, dh, da = deltas.unbind(-1)
    max_ratio = abs(math.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    rx, ry, rw, rh, ra = anchors.unbind(-1)
    cosa, sina = torch.cos(ra), torch.sin(ra)
    gx = dx * rw * cosa - dy * rh * sina + rx
    gy = dx * rw * sina + dy * rh * cosa + ry
    gw = rw * dw.exp()
    gh = rh * dh.exp()
    ga = norm_angle(math.pi * da + ra)
    return torch.stack([gx, gy, gw, gh, ga], dim=-1)


def _ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _ints(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def _strides(tensors):
    flat = [int(s) for t in tensors for s in t.stride()]
    return (C.c_int64 * len(flat))(*flat)


def _same_dtype(tensors, what):
    dt = tensors[0].dtype
    for t in tensors:
        if t.dtype != dt:
            raise TypeError("%s: all levels must share one dtype (%s vs %s)" % (what, dt, t.dtype))
        if t.dim() != 4:
            raise ValueError("%s: expected [B, C, H, W] tensors" % what)
    return dt


def fam_decode(fam_bbox_preds, strides, anchor_scale=4.0, anchor_angle=0.0, wh_ratio_clip=1e-6):
    """fam_bbox_preds[l] [B,5,H_l,W_l] (any strides; fp32/bf16/fp16) -> refined anchors [B,H_l,W_l,5] fp32."""
    preds = [p.detach() for p in fam_bbox_preds]
    dev = _lib.require_cuda(*preds)
    _same_dtype(preds, "fam_decode")
    B = preds[0].size(0)
    for p in preds:
        if p.size(1) != 5 or p.size(0) != B:
            raise ValueError("fam_decode: expected [B, 5, H, W] deltas with one batch size")
    outs = [torch.empty((B, p.size(2), p.size(3), 5), dtype=torch.float32, device=dev) for p in preds]
    st = (C.c_float * len(preds))(*[float(s) for s in strides])
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_fam_decode(len(preds), _ptrs(preds), _strides(preds), _ptrs(outs),
                                        _ints([p.size(2) for p in preds]), _ints([p.size(3) for p in preds]), st, B,
                                        float(anchor_scale), float(anchor_angle), float(wh_ratio_clip),
                                        _lib.dtype_code(preds[0]), _lib.stream_ptr(dev))
    _lib.check(rc, "fam_decode")
    return outs


def select_decode(cls_preds, bbox_preds, anchors, topk=2000, wh_ratio_clip=16 / 1000, return_index=False):
    """cls_preds[l] [B,C,H,W] logits, bbox_preds[l] [B,5,H,W] deltas, anchors[l] [B,H,W,5] (or [B,H*W,5]) fp32
    -> bboxes [B,n,5] fp32, scores [B,n,C] fp32 with n = sum_l min(H_l*W_l, topk)."""
    cls = [c.detach() for c in cls_preds]
    reg = [r.detach() for r in bbox_preds]
    dev = _lib.require_cuda(*cls, *reg, *anchors)
    dt = _same_dtype(cls + reg, "select_decode")
    B, Cn = cls[0].shape[:2]
    an = [a.to(torch.float32).contiguous() for a in anchors]
    Hs, Ws = [c.size(2) for c in cls], [c.size(3) for c in cls]
    for c, r, a in zip(cls, reg, an):
        if r.shape != (B, 5, c.size(2), c.size(3)) or c.size(1) != Cn or a.numel() != B * c.size(2) * c.size(3) * 5:
            raise ValueError("select_decode: inconsistent level shapes")
    n_total = sum(min(h * w, topk) if topk > 0 else h * w for h, w in zip(Hs, Ws))
    bboxes = torch.empty((B, n_total, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((B, n_total, Cn), dtype=torch.float32, device=dev)
    index = torch.empty((B, n_total), dtype=torch.int32, device=dev) if return_index else None
    lib = _lib.load()
    hs, ws = _ints(Hs), _ints(Ws)
    wbytes = lib.s2a_select_decode_workspace_bytes(len(cls), hs, ws, B)
    work = torch.empty((wbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.s2a_select_decode(len(cls), _ptrs(cls), _strides(cls), _ptrs(reg), _strides(reg), _ptrs(an), hs, ws, B,
                                   Cn, int(topk), float(wh_ratio_clip), _lib._DTYPES[dt], _lib.ptr(bboxes),
                                   _lib.ptr(scores), _lib.ptr(index), n_total, _lib.ptr(work), wbytes,
                                   _lib.stream_ptr(dev))
    _lib.check(rc, "select_decode")
    return (bboxes, scores, index) if return_index else (bboxes, scores)


@torch.no_grad()
def select_and_decode_torch(cls_preds, bbox_preds, anchors, topk=2000, wh_ratio_clip=16 / 1000):
    """models/head.py:684-717 in plain PyTorch, batched over images (parity witness for select_decode).
    Returns (bboxes [B,n,5], scores [B,n,C]) in the dtypes the reference produces."""
    scores_l, deltas_l, anchors_l = [], [], []
    for cls, reg, refine in zip(cls_preds, bbox_preds, anchors):
        B, Cn, H, W = cls.shape
        sc = cls.permute(0, 2, 3, 1).reshape(B, H * W, Cn).sigmoid()
        dl = reg.permute(0, 2, 3, 1).reshape(B, H * W, 5)
        an = refine.reshape(B, H * W, 5)
        if topk > 0 and H * W > topk:
            _, idx = sc.max(dim=2)[0].topk(topk, dim=1)
            sc = sc.gather(1, idx[..., None].expand(-1, -1, Cn))
            dl = dl.gather(1, idx[..., None].expand(-1, -1, 5))
            an = an.gather(1, idx[..., None].expand(-1, -1, 5))
        scores_l.append(sc)
        deltas_l.append(dl)
        anchors_l.append(an)
    bboxes = rboxes_decode_torch(torch.cat(anchors_l, 1), torch.cat(deltas_l, 1), wh_ratio_clip)
    return bboxes, torch.cat(scores_l, 1)
