"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in s2anet_b200/dist.py.  The compute
function is the oracle here (CPU box, no GPU); on the GPU box bench.py --gpus N runs the same
code with the CUDA kernel over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_iou_fn(a, g, tile_first, tile_step, tile_rows):
    """Same contract as box_iou_rotated_tiles(compact=True), computed by the CPU oracle."""
    from oracle import oracle as O
    B, N, M = a.size(0), a.size(1), g.size(1)
    ntiles = -(-N // tile_rows)
    tiles = list(range(tile_first, ntiles, tile_step))
    out = torch.full((B, len(tiles) * tile_rows, M), float("nan"))
    for k, t in enumerate(tiles):
        r0, r1 = t * tile_rows, min(N, (t + 1) * tile_rows)
        for b in range(B):
            out[b, k * tile_rows:k * tile_rows + (r1 - r0)] = torch.from_numpy(O.box_iou_rotated(a[b, r0:r1].numpy(), g[b].numpy()))
    return out


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from s2anet_b200 import dist as sd
    from s2anet_b200 import synth
    try:
        B, N, M, T = 2, 333, 40, 32            # 11 tiles of 32 rows (the last one partial): ranks get 6 and 5
        an = torch.from_numpy(synth.all_level_anchors(B, 3)[:, :N].copy())
        gt = torch.from_numpy(np.stack([synth.dota_like_gt(M, 50 + i) for i in range(B)]))
        ref = _oracle_iou_fn(an, gt, 0, 1, T)[:, :N]
        full = sd.sharded_box_iou(an, gt, iou_fn=_oracle_iou_fn, gather=True, tile_rows=T)
        assert torch.equal(full, ref), "all-gathered IoU differs from the single-process matrix"
        row_max, row_arg, gt_max, rows, local = sd.sharded_assign_stats(an, gt, iou_fn=_oracle_iou_fn, tile_rows=T)
        tiles, rows_expect = sd.shard_tiles(N, rank, world, T)
        assert tiles == list(range(rank, 11, world)) and torch.equal(rows, rows_expect)
        assert local.size(1) == rows.numel() and torch.equal(local, ref[:, rows])
        assert torch.equal(row_max, ref[:, rows].max(dim=2)[0]) and torch.equal(row_arg, ref[:, rows].max(dim=2)[1])
        assert torch.equal(gt_max, ref.max(dim=1)[0]), "per-GT maxima after MAX all-reduce"
        # detections: each rank contributes its own images
        K = 5
        dets = torch.full((B, K, 6), float(rank)) + torch.arange(B).view(B, 1, 1)
        labels = torch.full((B, K), float(10 + rank))
        counts = torch.tensor([rank + 1, rank + 2], dtype=torch.int32)
        gd, gl, gc = sd.gather_detections(dets, labels, counts)
        assert tuple(gd.shape) == (world * B, K, 6) and tuple(gl.shape) == (world * B, K)
        for r in range(world):
            assert torch.equal(gd[r * B:(r + 1) * B], torch.full((B, K, 6), float(r)) + torch.arange(B).view(B, 1, 1))
            assert torch.equal(gl[r * B:(r + 1) * B], torch.full((B, K), float(10 + r)))
            assert gc[r * B:(r + 1) * B].tolist() == [r + 1, r + 2]
        # the packed form (what the NMS finaliser writes): one collective, no pack step
        packed = torch.zeros((B, K + 1, 8))
        packed[:, :K, :6] = dets
        packed[:, :K, 6] = labels
        packed[:, K, 0] = counts.float()
        pd, pl, pc = sd.gather_detections(None, None, None, packed=packed)
        assert torch.equal(pd, gd) and torch.equal(pl, gl) and torch.equal(pc, gc)
        q.put((rank, "ok"))
    except Exception as e:      # surface the failure to the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_rows_cover_exactly():
    from s2anet_b200.dist import shard_rows
    for n in (0, 1, 63, 64, 65, 21824, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(b % 64 == 0 for b, _ in spans if b < n)
    assert shard_rows(21824, 7, 8) == (19264, 21824)


def test_shard_tiles_deal_rows_cyclically_and_cover_exactly():
    from s2anet_b200.box_iou_rotated import tile_rows_of
    from s2anet_b200.dist import shard_tiles
    for n in (0, 1, 255, 256, 257, 21824, 1000):
        for world in (1, 2, 3, 8):
            seen = torch.cat([shard_tiles(n, r, world)[1] for r in range(world)])
            assert sorted(seen.tolist()) == list(range(n))
            for r in range(world):
                assert torch.equal(shard_tiles(n, r, world)[1], tile_rows_of(n, r, world))
                assert torch.equal(shard_tiles(n, r, world, 256)[1], tile_rows_of(n, r, world, 256))
    # BASELINE config 4 on 8 GPUs: 682 tiles of 32 rows; the P5-P7 anchors (rows 20,480 ...: 42 tiles) reach every rank
    for r in range(8):
        tiles, rows = shard_tiles(21824, r, 8)
        assert tiles == list(range(r, 682, 8)) and sum(t >= 640 for t in tiles) in (5, 6)
        assert rows.numel() == 32 * len(tiles)               # 21,824 = 682 x 32: no partial tile


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
