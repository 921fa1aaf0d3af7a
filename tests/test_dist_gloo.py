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


def _oracle_iou_fn(a, g, rb, re, out):
    from oracle import oracle as O
    for b in range(a.size(0)):
        out[b, rb:re] = torch.from_numpy(O.box_iou_rotated(a[b, rb:re].numpy(), g[b].numpy()))
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
        B, N, M = 2, 333, 40
        an = torch.from_numpy(synth.all_level_anchors(B, 3)[:, :N].copy())
        gt = torch.from_numpy(np.stack([synth.dota_like_gt(M, 50 + i) for i in range(B)]))
        ref = torch.empty(B, N, M)
        _oracle_iou_fn(an, gt, 0, N, ref)
        full = sd.sharded_box_iou(an, gt, iou_fn=_oracle_iou_fn, gather=True)
        assert torch.equal(full, ref), "all-gathered IoU differs from the single-process matrix"
        row_max, row_arg, gt_max, (b0, e0), local = sd.sharded_assign_stats(an, gt, iou_fn=_oracle_iou_fn)
        assert (b0, e0) == sd.shard_rows(N, rank, world)
        assert torch.equal(local, ref[:, b0:e0])
        assert torch.equal(row_max, ref[:, b0:e0].max(dim=2)[0]) and torch.equal(row_arg, ref[:, b0:e0].max(dim=2)[1])
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
