"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink on the box, gloo in the CPU tests).

* inference shards by image batch: every rank runs the head + NMS on its own images; the only exchange is the
  detections (<= max_per_img x 7 floats + a count per image).  `DetectionExchange` does it without a pack step and
  without a library collective: the NMS finaliser (mc_emit_kernel) stores every kept detection straight into the
  packed [world*B, K+1, 8] buffer of EVERY rank through NVLink peer pointers (torch symmetric memory), followed by
  one signal-pad barrier; `gather_detections` is the plain NCCL form (one all_gather_into_tensor of the same packed
  buffer, which the emit kernel also fills directly);
* the anchor x GT IoU matrix shards by anchor rows, dealt out in 32-row tiles CYCLICALLY (rank r owns tiles r,
  r + world, ...): contiguous blocks would give the last rank every P5-P7 anchor -- large boxes whose pairs mostly
  reach the clipper -- and the first ranks only 32 px P3 anchors the reject test kills (round 1: 0.59 efficiency at 8
  GPUs with NO collective, pure imbalance).  Either the tiles are all-gathered (literal box_iou_rotated result on
  every rank: NVLink-bound, 2.8 GB) or -- the form label assignment needs (models/utils.py:115-144) -- every rank
  keeps its rows and only the per-GT column maxima are combined with one MAX all-reduce of a [B, M] tensor.

No collective is invented where the path has none: NMS itself is never sharded across GPUs.
"""
import torch
import torch.distributed as dist

TILE_ROWS = 32           # box_iou_rotated.IOU_TILE_ROWS


def shard_rows(n, rank, world, align=64):
    """Contiguous row block of rank `rank` (kept for callers that need one span per rank, e.g. a row-range call of
    s2a_box_iou_rotated): blocks are multiples of `align` except the last one; returns (begin, end), possibly empty.
    The IoU sharding below does NOT use it (see the module docstring)."""
    per = -(-n // world)
    per = -(-per // align) * align
    begin = min(n, rank * per)
    end = min(n, begin + per)
    return begin, end


def shard_tiles(n, rank, world, tile_rows=TILE_ROWS):
    """Row tiles of rank `rank` under the cyclic deal: (tile indices, global row indices in packed order)."""
    ntiles = -(-n // tile_rows)
    tiles = list(range(rank, ntiles, world))
    rows = [torch.arange(t * tile_rows, min(n, (t + 1) * tile_rows)) for t in tiles]
    return tiles, (torch.cat(rows) if rows else torch.zeros((0,), dtype=torch.long))


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _default_iou_fn(anchors, gts, tile_first, tile_step, tile_rows):
    from .box_iou_rotated import box_iou_rotated_tiles
    return box_iou_rotated_tiles(anchors, gts, tile_first, tile_step, compact=True, tile_rows=tile_rows)


def sharded_box_iou(anchors, gts, iou_fn=None, group=None, gather=True, tile_rows=TILE_ROWS):
    """anchors [B,N,5], gts [B,M,5] (replicated on every rank) -> IoU.

    iou_fn(anchors, gts, tile_first, tile_step, tile_rows) returns the PACKED local tiles
    [B, ntiles_mine * tile_rows, M] (rows past N in the last tile unspecified); the default is the CUDA kernel
    (s2a_box_iou_rotated_tiles, compact).  Only the local rows are ever allocated on a rank.
    gather=False -> (local [B, n_mine, M], rows [n_mine] global row index of each local row).
    gather=True  -> the full [B,N,M] matrix on every rank: one all_gather_into_tensor of the equally sized packed
    blocks (ranks with one tile fewer pad) and one strided copy that undoes the cyclic deal."""
    iou_fn = iou_fn or _default_iou_fn
    rank, world = _world(group)
    B, N, _ = anchors.shape
    M = gts.size(1)
    tiles, rows = shard_tiles(N, rank, world, tile_rows)
    local = iou_fn(anchors, gts, rank, world, tile_rows)
    if not gather:
        return local[:, : rows.numel()], rows            # (only the last tile of the matrix can be partial: it is last)
    if world == 1:
        return local[:, :N]
    ntiles = -(-N // tile_rows)
    tmax = -(-ntiles // world)                          # tiles of rank 0 = the most any rank has
    send = local
    if len(tiles) < tmax:
        send = torch.zeros((B, tmax * tile_rows, M), dtype=local.dtype, device=local.device)
        send[:, : local.size(1)] = local
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv.view(world * B, tmax * tile_rows, M), send.contiguous(), group=group)
    # recv[r, b, k*tile_rows + i] is global row (k * world + r) * tile_rows + i: [r, b, k, i, m] -> [b, k, r, i, m]
    full = recv.view(world, B, tmax, tile_rows, M).permute(1, 2, 0, 3, 4).reshape(B, tmax * world * tile_rows, M)
    return full[:, :N]


def sharded_assign_stats(anchors, gts, iou_fn=None, group=None, tile_rows=TILE_ROWS):
    """RAW IoU statistics label assignment starts from (models/utils.py:115-130), without moving the matrix: per-anchor
    max / argmax over GTs for the local rows, and the per-GT maxima over ALL anchors (one [B, M] MAX all-reduce).
    The reference's invalid-anchor and out-of-range masks (models/utils.py:88-113) are NOT applied here -- the fused
    s2a_assign_labels kernel is the op that implements the whole rule set; this is the sharded building block the
    matrix form of BASELINE config 4 asks for.
    Returns (row_max [B,n_mine], row_argmax [B,n_mine], gt_max [B,M], rows [n_mine], local_iou [B,n_mine,M])."""
    local, rows = sharded_box_iou(anchors, gts, iou_fn=iou_fn, group=group, gather=False, tile_rows=tile_rows)
    B, M = gts.size(0), gts.size(1)
    if rows.numel() > 0:
        row_max, row_arg = local.max(dim=2)
        gt_max = local.max(dim=1)[0]
    else:
        row_max = local.new_zeros((B, 0))
        row_arg = torch.zeros((B, 0), dtype=torch.long, device=local.device)
        gt_max = local.new_full((B, M), -1.0)
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(gt_max, op=dist.ReduceOp.MAX, group=group)
    return row_max, row_arg, gt_max, rows, local


# ---- detections ---------------------------------------------------------------------------------------------

def packed_views(packed, K):
    """(dets [n,K,6], labels [n,K], counts [n] int32) views / values of a packed [n, K+1, 8] detection buffer: rows
    < K are (x, y, w, h, theta, score, label, 0), row K holds the count in column 0."""
    return packed[:, :K, :6], packed[:, :K, 6], packed[:, K, 0].to(torch.int32)


def gather_detections(dets, labels, counts, group=None, packed=None):
    """NCCL form of the detection exchange: dets [B,K,6], labels [B,K], counts [B] of every rank ->
    ([world*B,K,6], [world*B,K], [world*B]) on every rank.  One packed buffer, one collective.  `packed`
    ([B,K+1,8], filled directly by multiclass_nms_rotated_packed) skips the pack."""
    rank, world = _world(group)
    if world == 1 and packed is None:
        return dets, labels, counts
    if packed is None:
        B, K, _ = dets.shape
        packed = torch.zeros((B, K + 1, 8), dtype=torch.float32, device=dets.device)
        packed[:, :K, :6] = dets
        packed[:, :K, 6] = labels
        packed[:, K, :] = 0
        packed[:, K, 0] = counts.to(torch.float32)
    B, K = packed.size(0), packed.size(1) - 1
    if world == 1:
        return packed_views(packed, K)
    recv = torch.empty((world * B, K + 1, 8), dtype=torch.float32, device=packed.device)
    dist.all_gather_into_tensor(recv, packed, group=group)
    return packed_views(recv, K)


class DetectionExchange:
    """The detection all-gather fused into the NMS finaliser over NVLink peer memory.

    Every rank owns `nbuf` symmetric buffers [world*B, K+1, 8] (torch.distributed._symmetric_memory: cudaMalloc'd
    with peer access, pointers exchanged once at construction).  `targets(i)` are the peer pointers of buffer i, which
    multiclass_nms_rotated_batched(push_to=...) hands to mc_emit_kernel: each kept detection is stored into slot
    rank*B + b of EVERY rank's buffer as it is ranked -- no pack kernels, no staging copy, no collective call.
    `finish(i)` enqueues one signal-pad barrier on the current stream: after it, every rank's buffer i holds all
    world*B images.  Use the buffers round-robin (nbuf >= 2): a rank may run one step ahead of a peer that is still
    reading the previous buffer, never two (the barrier of step s+1 is passed only after every rank has enqueued the
    reads of step s before its step-s+1 kernels, in stream order)."""

    def __init__(self, B, K, device, group=None, nbuf=2):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.B, self.K = B, K
        self.bufs, self.handles = [], []
        for _ in range(nbuf):
            t = symm.empty((self.world * B, K + 1, 8), dtype=torch.float32, device=device)
            t.zero_()
            h = symm.rendezvous(t, self.group.group_name)
            self.bufs.append(t)
            self.handles.append(h)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def targets(self, i):
        """(peer base pointers of buffer i, first slot of this rank)."""
        return [int(p) for p in self.handles[i].buffer_ptrs], self.rank * self.B

    def finish(self, i):
        self.handles[i].barrier(channel=0)
        return packed_views(self.bufs[i], self.K)
