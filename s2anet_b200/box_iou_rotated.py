"""box_iou_rotated -- same call as the reference's `utils.box_iou_rotated.box_iou_rotated`
(reference: utils/box_iou_rotated/__init__.py:1, src/box_iou_rotated.h:22-37, _cuda.cu:65-101)."""
import torch

from . import _lib, _torch_ext


def box_iou_rotated(boxes1, boxes2, _flags=0):
    """IoU of rotated boxes: boxes1 [N,5] x boxes2 [M,5] (x, y, w, h, theta[rad]) -> [N,M] float32.

    Unlike the reference CUDA op (which reads raw pointers and silently mis-reads strided
    input, models/utils.py:51-56) non-contiguous inputs are accepted.
    """
    dev = _lib.require_cuda(boxes1, boxes2)
    if boxes1.dim() != 2 or boxes2.dim() != 2 or boxes1.size(-1) != 5 or boxes2.size(-1) != 5:
        if not (boxes1.numel() == 0 or boxes2.numel() == 0):
            raise ValueError("boxes must be [N,5] and [M,5]")
    n, m = boxes1.size(0), boxes2.size(0)
    ext = _torch_ext.module() if _flags == 0 else None
    if ext is not None:                 # the torch-extension binding: same C-ABI call, a third of the per-call host cost
        out = ext.box_iou_rotated(boxes1, boxes2)
        _lib.check(0, "box_iou_rotated" if n and m else "")        # (launch accounting of bench.py's gpu_launches)
        return out
    out = torch.empty((n, m), dtype=torch.float32, device=dev)
    if n == 0 or m == 0:
        return out
    b1 = boxes1.to(torch.float32).contiguous()
    b2 = boxes2.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_box_iou_rotated(_lib.ptr(b1), n, _lib.ptr(b2), m, 1, _lib.ptr(out), m, 0, n, _flags,
                                             _lib.stream_ptr(dev))
    _lib.check(rc, "box_iou_rotated")
    return out


def box_iou_rotated_batched(boxes1, boxes2, row_begin=0, row_end=None, out=None, _flags=0):
    """Batched anchor x GT IoU (BASELINE config 4): boxes1 [B,N,5], boxes2 [B,M,5] -> [B,N,M].

    row_begin/row_end compute only anchor rows [row_begin, row_end) of every image (the rows a
    rank owns when the matrix is sharded by anchor rows); other rows of `out` are left untouched.
    """
    dev = _lib.require_cuda(boxes1, boxes2)
    B, n, _ = boxes1.shape
    m = boxes2.size(1)
    if boxes2.size(0) != B:
        raise ValueError("batch mismatch")
    row_end = n if row_end is None else row_end
    if out is None:
        out = torch.empty((B, n, m), dtype=torch.float32, device=dev)
    b1 = boxes1.to(torch.float32).contiguous()
    b2 = boxes2.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_box_iou_rotated(_lib.ptr(b1), n, _lib.ptr(b2), m, B, _lib.ptr(out), m, row_begin,
                                             row_end, _flags, _lib.stream_ptr(dev))
    _lib.check(rc, "box_iou_rotated_batched")
    return out


IOU_TILE_ROWS = 32      # the unit the anchor rows are dealt out in (32, 64, 128 or 256; CTAs take up to eight tiles)


def tile_rows_of(n, tile_first, tile_step, tile_rows=IOU_TILE_ROWS):
    """Global row indices (a LongTensor on the CPU) of the row tiles tile_first, tile_first + tile_step, ... of an
    n-row matrix, in the packed order s2a_box_iou_rotated_tiles(compact=1) writes them."""
    ntiles = -(-n // tile_rows)
    rows = [torch.arange(t * tile_rows, min(n, (t + 1) * tile_rows)) for t in range(tile_first, ntiles, tile_step)]
    return torch.cat(rows) if rows else torch.zeros((0,), dtype=torch.long)


def box_iou_rotated_tiles(boxes1, boxes2, tile_first, tile_step, compact=True, out=None, tile_rows=IOU_TILE_ROWS,
                          _flags=0):
    """The anchor-row shard of one rank: boxes1 [B,N,5], boxes2 [B,M,5]; the N rows are cut into tiles
    of `tile_rows` and tiles tile_first, tile_first + tile_step, ... are computed (rank r of w: (r, w) -- a cyclic
    deal, so every rank gets the same mix of FPN levels).  compact=True returns [B, n_mine_padded, M] holding only
    those tiles, packed (n_mine_padded = tiles x tile_rows; `tile_rows_of` maps packed rows back to anchors);
    compact=False writes them at their global rows of a [B,N,M] tensor (`out`, or a new uninitialised one)."""
    dev = _lib.require_cuda(boxes1, boxes2)
    B, n, _ = boxes1.shape
    m = boxes2.size(1)
    if boxes2.size(0) != B:
        raise ValueError("batch mismatch")
    if tile_rows not in (32, 64, 128, 256):
        raise ValueError("tile_rows must be 32, 64, 128 or 256")
    ntiles = -(-n // tile_rows)
    mine = len(range(tile_first, ntiles, tile_step))
    if out is None:
        out = torch.empty((B, mine * tile_rows if compact else n, m), dtype=torch.float32, device=dev)
    b1 = boxes1.to(torch.float32).contiguous()
    b2 = boxes2.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_box_iou_rotated_tiles(_lib.ptr(b1), n, _lib.ptr(b2), m, B, _lib.ptr(out), out.stride(1),
                                                   out.stride(0), tile_rows, tile_first, tile_step, 1 if compact else 0,
                                                   _flags, _lib.stream_ptr(dev))
    _lib.check(rc, "box_iou_rotated_tiles")
    return out


def bbox_iou_rotated(rboxes1, rboxes2):
    """reference: utils/metrics.py:85-107 -- accepts 5- or 6-column boxes, casts to fp32."""
    assert rboxes1.size(-1) in [0, 5, 6]
    assert rboxes2.size(-1) in [0, 5, 6]
    if rboxes2.size(-1) == 6:
        rboxes2 = rboxes2[..., :5]
    if rboxes1.size(-1) == 6:
        rboxes1 = rboxes1[..., :5]
    return box_iou_rotated(rboxes1.float(), rboxes2.float())
