"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink on the box, gloo in the CPU tests).

* inference shards by image batch: every rank runs the head + NMS on its own images; the only
  exchange is a fixed-shape all-gather of the packed detections (<= max_per_img x 7 floats per image);
* the anchor x GT IoU matrix shards by anchor rows: rank r computes rows [begin_r, end_r) of every
  image with the same kernel (row_begin/row_end of s2a_box_iou_rotated).  Either the row blocks
  are all-gathered (literal box_iou_rotated result on every rank) or -- the form label assignment
  needs (models/utils.py:115-144) -- each rank keeps its rows, and only the per-GT column maxima are
  combined with one MAX all-reduce of a [B, M] tensor.

No collective is invented where the path has none: NMS itself is never sharded across GPUs.
"""
import torch
import torch.distributed as dist


def shard_rows(n, rank, world, align=64):
    """Contiguous row block of rank `rank`: blocks are multiples of `align` (the IoU kernel's row
    tile) except the last one; returns (begin, end), possibly empty."""
    per = -(-n // world)
    per = -(-per // align) * align
    begin = min(n, rank * per)
    end = min(n, begin + per)
    return begin, end


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def sharded_box_iou(anchors, gts, iou_fn=None, group=None, gather=True):
    """anchors [B,N,5], gts [B,M,5] (replicated on every rank) -> IoU [B,N,M].

    iou_fn(anchors, gts, row_begin, row_end, out) must fill out[:, row_begin:row_end]; the default is
    the CUDA kernel.  gather=True returns the full matrix on every rank (all_gather of equal-sized,
    padded row blocks); gather=False returns (local_rows [B, end-begin, M], (begin, end))."""
    if iou_fn is None:
        from .box_iou_rotated import box_iou_rotated_batched

        def iou_fn(a, g, rb, re, out):
            return box_iou_rotated_batched(a, g, rb, re, out=out)
    rank, world = _world(group)
    B, N, _ = anchors.shape
    M = gts.size(1)
    begin, end = shard_rows(N, rank, world)
    if world == 1:
        out = torch.empty((B, N, M), dtype=torch.float32, device=anchors.device)
        iou_fn(anchors, gts, 0, N, out)
        return out if gather else (out, (0, N))
    per = shard_rows(N, 0, world)[1]
    full = torch.empty((B, N, M), dtype=torch.float32, device=anchors.device)
    if end > begin:
        iou_fn(anchors, gts, begin, end, full)
    local = full[:, begin:end]
    if not gather:
        return local.contiguous(), (begin, end)
    send = torch.zeros((B, per, M), dtype=torch.float32, device=anchors.device)
    send[:, : end - begin] = local
    recv = torch.empty((world * B, per, M), dtype=torch.float32, device=anchors.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, B, per, M)
    for r in range(world):
        rb, re = shard_rows(N, r, world)
        if re > rb:
            full[:, rb:re] = recv[r, :, : re - rb]
    return full


def sharded_assign_stats(anchors, gts, iou_fn=None, group=None):
    """The quantities label assignment consumes (models/utils.py:115-130) without moving the matrix:
    per-anchor max/argmax over GTs for the local rows, and the per-GT maxima over ALL anchors
    (one [B, M] MAX all-reduce).  Returns (row_max [B,n_loc], row_argmax [B,n_loc], gt_max [B,M],
    (begin, end), local_iou)."""
    local, (begin, end) = sharded_box_iou(anchors, gts, iou_fn=iou_fn, group=group, gather=False)
    B, M = gts.size(0), gts.size(1)
    if end > begin:
        row_max, row_arg = local.max(dim=2)
        gt_max = local.max(dim=1)[0]
    else:
        row_max = local.new_zeros((B, 0))
        row_arg = torch.zeros((B, 0), dtype=torch.long, device=local.device)
        gt_max = local.new_full((B, M), -1.0)
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(gt_max, op=dist.ReduceOp.MAX, group=group)
    return row_max, row_arg, gt_max, (begin, end), local


def gather_detections(dets, labels, counts, group=None):
    """Fixed-shape detection exchange: dets [B,K,6], labels [B,K], counts [B] of every rank ->
    ([world*B,K,6], [world*B,K], [world*B]) on every rank.  One packed buffer, one collective."""
    rank, world = _world(group)
    if world == 1:
        return dets, labels, counts
    B, K, _ = dets.shape
    packed = torch.empty((B, K + 1, 7), dtype=torch.float32, device=dets.device)
    packed[:, :K, :6] = dets
    packed[:, :K, 6] = labels
    packed[:, K, :] = 0
    packed[:, K, 0] = counts.to(torch.float32)
    recv = torch.empty((world * B, K + 1, 7), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(recv, packed, group=group)
    return recv[:, :K, :6], recv[:, :K, 6], recv[:, K, 0].to(torch.int32)
