#!/usr/bin/env python
"""bench.py -- S2ANet head + rotated NMS images/s at 1024x1024 on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic input: the FAM/ODM head
(stock cuDNN conv towers + AlignConv + ORConv2d/RotationInvariantPooling from this library) over the
five FPN levels of `--batch` 1024x1024 images per GPU, box decode, per-level top-2000 and the fused
15-class rotated NMS.  `value` times it with the FPN features resident in HBM; `e2e` times the
same call with the features in pinned host memory (H2D inside the timed region, double-buffered on a
copy stream) and the detections read back to the host.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "S2ANet head+NMS images/s at 1024^2"
UNIT = "images/s"
IMG = 1024
STRIDES = (8, 16, 32, 64, 128)
NUM_CLASSES = 15
POSITIONS = sum((IMG // s) ** 2 for s in STRIDES)            # 21,824
ALIGN_FLOPS_PER_IMAGE = 2.0 * POSITIONS * 256 * 2304          # 25.744 GFLOP


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p.get("bf16_tflops"), "bf16_sustained": p.get("bf16_tflops_sustained"),
                "hbm": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic(batch):
    """DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json; captured at batch 8, the default): read + write, or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if batch != 8 or not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get("conv_tc_kernel<ALIGN,bf16>")
    return None if not t else {"dram_bytes": t["dram_bytes_read"] + t["dram_bytes_write"],
                               "algorithmic_bytes": t["algorithmic_bytes"], "source": t["source"]}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def make_feats(torch, batch, seed, device, dtype, pinned=False):
    g = torch.Generator().manual_seed(seed)
    feats = []
    for s in STRIDES:
        h = IMG // s
        t = torch.randn(batch, 256, h, h, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
        if pinned:
            t = t.pin_memory()
        else:
            t = t.to(device)
        feats.append(t)
    return feats


# ------------------------------------------------------------------------------------------------
# CPU reference path (the reference's own CPU kernels where it has them)
# ------------------------------------------------------------------------------------------------
class CpuReferenceHead:
    """Head + NMS on the host: stock torch CPU convs, AlignConv = the reference's get_offset formula +
    torchvision.ops.deform_conv2d (the reference has no CPU deform conv, deform_conv.py:58-59),
    ORConv2d = ARF index scatter + F.conv2d, RotationInvariantPooling, decode, and ml_nms_rotated
    through the reference's CPU extension (oracle/_ref/ext_cpu, kind "reference") or, where that
    was not prebuilt, the oracle port (kind "port")."""

    def __init__(self, head):
        import torch
        from oracle import build_oracle
        from oracle import oracle as O
        self.torch, self.O = torch, O
        self.head = head
        self.sd = {k: v.detach().float().cpu() for k, v in head.state_dict().items()}
        self.ml_ext = build_oracle.load_ref_extension("ml_nms_rotated_cuda", "cpu")
        self.kind = "reference" if self.ml_ext is not None else "port"
        w = self.sd["or_conv.weight"]
        self.w_rot = torch.from_numpy(O.arf_forward(w.numpy(), self.sd["or_conv.indices"].numpy()))

    def tower(self, x, name, n):
        F = self.torch.nn.functional
        for i in range(n):
            x = F.relu(F.conv2d(x, self.sd["%s.%d.0.weight" % (name, i)], self.sd["%s.%d.0.bias" % (name, i)], padding=1))
        return x

    def level(self, x, stride):
        import torchvision
        torch, F, h = self.torch, self.torch.nn.functional, self.head
        fam_reg = F.conv2d(self.tower(x, "fam_reg_ls", 2), self.sd["fam_reg_head.weight"], self.sd["fam_reg_head.bias"])
        _ = F.conv2d(self.tower(x, "fam_cls_ls", 2), self.sd["fam_cls_head.weight"], self.sd["fam_cls_head.bias"])
        B, _, H, W = fam_reg.shape
        from s2anet_b200.head import rboxes_decode
        anchors = h.grid_anchors(H, W, stride, "cpu")
        refine = rboxes_decode(anchors[None], fam_reg.permute(0, 2, 3, 1).reshape(B, H * W, 5), wh_ratio_clip=1e-6)
        offs = torch.stack([h.align_conv.get_offset(refine[i], (H, W), stride) for i in range(B)])
        al = F.relu(torchvision.ops.deform_conv2d(x, offs, self.sd["align_conv.deform_conv.weight"], padding=1))
        orf = F.conv2d(al, self.w_rot, self.sd["or_conv.bias"], padding=1)
        pooled = orf.view(B, -1, 8, H, W).max(dim=2)[0]
        cls = F.conv2d(self.tower(pooled, "odm_cls_ls", 2), self.sd["odm_cls_head.weight"], self.sd["odm_cls_head.bias"], padding=1)
        reg = F.conv2d(self.tower(orf, "odm_reg_ls", 2), self.sd["odm_reg_head.weight"], self.sd["odm_reg_head.bias"], padding=1)
        return (None, fam_reg, cls, reg, anchors, refine.reshape(B, H, W, 5))

    def detect(self, feats):
        torch = self.torch
        outs = [self.level(x, s) for x, s in zip(feats, STRIDES)]
        bboxes, scores = self.head.select_and_decode(outs)
        results = []
        for b in range(bboxes.size(0)):
            if self.ml_ext is not None:
                mask = scores[b] > self.head.score_thres_before_nms
                bx = bboxes[b][:, None].expand(-1, scores.size(2), 5)[mask]
                sc = scores[b][mask]
                lb = mask.nonzero(as_tuple=False)[:, 1].to(bx)
                if bx.shape[0] == 0:
                    results.append((bx.new_zeros((0, 6)), lb))
                    continue
                keep = self.ml_ext.ml_nms_rotated(bx.contiguous(), sc.contiguous(), lb.contiguous(), self.head.iou_thres_nms)
                keep = keep[: self.head.max_per_img]
                results.append((torch.cat([bx[keep], sc[keep, None]], 1), lb[keep]))
            else:
                d, l = self.O.multiclass_nms_rotated(bboxes[b].numpy(), scores[b].numpy(), self.head.score_thres_before_nms,
                                                     self.head.iou_thres_nms, self.head.max_per_img)
                results.append((torch.from_numpy(d), torch.from_numpy(l)))
        return results


def build_head(torch, device, dtype, seed=0):
    from s2anet_b200.head import S2ANetHead
    head = S2ANetHead(NUM_CLASSES)
    head.init_synthetic(seed)
    head = head.eval()
    if device is not None:
        head = head.to(device)
    if dtype is not None and dtype != torch.float32:
        head = head.to(dtype)
    for m in head.modules():                      # channels_last for the stock 4-D conv weights only
        if type(m) is torch.nn.Conv2d:
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    return head


def run_reference(args, rank, world):
    """Reference arm: the CPU path, rank 0 only, a bounded sample (1 image per step)."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    head = build_head(torch, None, torch.float32, seed=0)
    calib = os.path.join(ROOT, "gpurun_out", "calibrated_bias.pt")
    if os.path.exists(calib):
        head.odm_cls_head.bias.data.copy_(torch.load(calib))
    ref = CpuReferenceHead(head)
    feats = make_feats(torch, 1, 4, "cpu", torch.float32)
    feats = [f.contiguous() for f in feats]
    for _ in range(args.warmup):
        res = ref.detect(feats)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ref.detect(feats)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "S2ANet R-50-FPN head + 15-class rotated NMS, 1024x1024 (CPU reference path)",
                   "batch_per_step": 1, "detections_img0": int(res[0][0].shape[0])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
                         "sample": "1 image per step, all 5 FPN levels, torch CPU convs + torchvision deform_conv2d "
                                   "(all threads) + reference ml_nms_rotated CPU kernel (1 thread)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--candidates", type=int, default=3000, help="target (box, class) candidates per image")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repository has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cudnn.benchmark = True
    from s2anet_b200 import _lib
    from s2anet_b200 import dist as sdist
    from s2anet_b200.alignconv import alignconv_forward
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    B = args.batch
    head = build_head(torch, dev, dtype, seed=0)
    # two rotating input sets (> L2 together with the ~GBs of activations each step writes)
    feat_sets = [make_feats(torch, B, 4 + rank + 100 * i, dev, dtype) for i in range(2)]
    ncand = head.calibrate_scores(feat_sets[0], args.candidates)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        torch.save(head.odm_cls_head.bias.detach().float().cpu(), os.path.join(ROOT, "gpurun_out", "calibrated_bias.pt"))

    graphs = {}

    def step(feats):
        """One pass of the hot path.  The sync-free head+NMS is captured once per input buffer set into
        a CUDA graph and replayed (the detection all-gather stays outside the graph)."""
        key = id(feats)
        if args.no_graph:
            out_ = head.detect(feats)
        else:
            if key not in graphs:
                for _ in range(2):                   # warm every lazy path (weight packing, cuDNN autotune)
                    head.detect(feats)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                before = _lib.launches
                with torch.cuda.graph(g):
                    out_ = head.detect(feats)
                graphs[key] = (g, out_, _lib.launches - before)     # kernels of this library inside the graph
            g, out_, n_mine = graphs[key]
            g.replay()
            _lib.launches += n_mine
        dets, labels, counts = out_
        if world > 1:
            dets, labels, counts = sdist.gather_detections(dets, labels, counts)
        return dets, labels, counts

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        out = step(feat_sets[i % 2])
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        out = step(feat_sets[i % 2])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launches - l0
    clocks = sampler.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)
    ndet = int(out[2][:B].float().mean().item())

    # ---- e2e: host (pinned) features in, detections back on the host, copies inside the timed region
    host_sets = [make_feats(torch, B, 4 + rank + 100 * i, dev, dtype, pinned=True) for i in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_bytes = sum(t_.numel() * t_.element_size() for t_ in host_sets[0])
    K = head.max_per_img
    host_out = (torch.empty((world * B if world > 1 else B, K, 6), dtype=torch.float32).pin_memory(),
                torch.empty((world * B if world > 1 else B, K), dtype=torch.float32).pin_memory(),
                torch.empty((world * B if world > 1 else B,), dtype=torch.int32).pin_memory())
    d2h_bytes = sum(t_.numel() * t_.element_size() for t_ in host_out)
    dev_bufs = [[torch.empty_like(f, device=dev) for f in host_sets[0]] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            for d, h in zip(dev_bufs[i % 2], host_sets[i % 2]):
                d.copy_(h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream(dev)
        for j in range(2):
            freed[j].record(cur)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            cur.wait_event(ready[i % 2])
            dets, labels, counts = step(dev_bufs[i % 2])
            freed[i % 2].record(cur)
            host_out[0].copy_(dets, non_blocking=True)
            host_out[1].copy_(labels, non_blocking=True)
            host_out[2].copy_(counts, non_blocking=True)
        cur.synchronize()

    e2e_loop(min(3, args.warmup))
    sync_all()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())

    # ---- second headline metric: rotated IoU pairs/s on BASELINE configs[3]
    # (21,824 anchors x 500 GTs x 64 images; anchor rows sharded over the ranks, no collective needed
    # for the assignment-aware form: every rank keeps its row block)
    from s2anet_b200 import synth
    from s2anet_b200.box_iou_rotated import box_iou_rotated_batched
    import numpy as np
    IB, IM = 64, 500
    an = torch.from_numpy(synth.all_level_anchors(IB, 3)).to(dev)
    gt = torch.from_numpy(np.stack([synth.dota_like_gt(IM, 100 + i) for i in range(IB)])).to(dev)
    rb, re = sdist.shard_rows(an.size(1), rank, world)
    iou_out = torch.empty((IB, an.size(1), IM), dtype=torch.float32, device=dev)      # 2.8 GB, rows [rb, re) written
    for _ in range(3):
        box_iou_rotated_batched(an, gt, rb, re, out=iou_out)
    sync_all()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    i0.record()
    for _ in range(reps):
        box_iou_rotated_batched(an, gt, rb, re, out=iou_out)
    i1.record()
    sync_all()
    ti = torch.tensor([i0.elapsed_time(i1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ti, op=dist.ReduceOp.MAX)
    iou_ms = float(ti.item())
    pairs = IB * an.size(1) * IM
    iou_bytes = 4.0 * pairs + 20.0 * IB * (an.size(1) + IM)
    peaks0 = load_peaks()
    iou_metric = {"metric": "rotated IoU pairs/s", "value": pairs / (iou_ms / 1e3), "unit": "pairs/s", "ms": iou_ms,
                  "config": "21,824 anchors x 500 GTs x 64 images (BASELINE configs[3]), anchor rows sharded over %d GPU(s)" % world,
                  "roofline": {"bound": "hbm", "achieved": iou_bytes / (iou_ms / 1e3) / 1e9 / world, "peak": peaks0["hbm"],
                               "unit": "GB/s", "frac": iou_bytes / (iou_ms / 1e3) / 1e9 / world / peaks0["hbm"],
                               "note": "algorithmic bytes 4*N*M + 20*(N+M) per image; per-GPU figure"}}
    del iou_out, an, gt

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel of this library: the AlignConv tcgen05 implicit GEMM,
    # timed alone with CUDA events on its launch stream over all five levels of the batch
    peaks = load_peaks()
    roof = None
    if dtype != torch.float32:
        w = head.align_conv.deform_conv.weight
        outs = head.forward_levels(feat_sets[0])
        refines = [o[5] for o in outs]
        del outs
        from s2anet_b200.conv_tc import alignconv_forward_tc_multi

        def align_all():                 # what the step runs: ONE persistent launch over the five levels
            alignconv_forward_tc_multi(feat_sets[0], refines, w, STRIDES)
        for _ in range(3):
            align_all()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a0.record()
        for _ in range(reps):
            align_all()
        a1.record()
        torch.cuda.synchronize()
        t_align = a0.elapsed_time(a1) / reps / 1e3
        p3 = feat_sets[0][0]
        for _ in range(3):
            alignconv_forward(p3, refines[0], w, 8)
        torch.cuda.synchronize()
        a0.record()
        for _ in range(reps):
            alignconv_forward(p3, refines[0], w, 8)
        a1.record()
        torch.cuda.synchronize()
        t_p3 = a0.elapsed_time(a1) / reps / 1e3
        flops_p3 = 2.0 * B * 128 * 128 * 256 * 2304
        ach = flops_p3 / t_p3 / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel<ALIGN,bf16> (AlignConv, P3 level of the batch, one launch)",
                "achieved": ach, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_burst"],
                "peak_source": peaks["source"] + " bf16_tflops (burst: kernel timed alone)", "traffic": load_traffic(B),
                "all_levels_tflops": ALIGN_FLOPS_PER_IMAGE * B / t_align / 1e12,
                "alignconv_share_of_step": t_align / (ms / 1e3 / args.steps)}

    cpu = None
    line_iou = iou_metric
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_head = build_head(torch, None, torch.float32, seed=0)
        cpu_head.odm_cls_head.bias.data.copy_(head.odm_cls_head.bias.detach().float().cpu())
        ref = CpuReferenceHead(cpu_head)
        cf = [f[:1].float().cpu().contiguous() for f in feat_sets[0]]
        t0 = time.perf_counter()
        r = ref.detect(cf)
        dt = time.perf_counter() - t0
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": ref.kind,
               "sample": "1 of the %d images of one step (all 5 FPN levels, %d candidates -> %d detections), fp32, "
                         "timed once: %.1f s" % (B, int(ncand), int(r[0][0].shape[0]), dt)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "S2ANet R-50-FPN head + 15-class rotated NMS at 1024x1024, batch %d per GPU "
                               "(BASELINE configs[4] per-GPU shard; configs[2] is the same at batch 1)" % B,
                   "global_batch": world * B, "positions_per_image": POSITIONS, "nms_candidates_per_image": int(ncand),
                   "detections_per_image": ndet, "parallelism": "dp%d (image-sharded, detection all-gather)" % world,
                   "l2": "two rotating feature sets (2 x %.0f MB) plus >1 GB of activations per step: working set > 126 MB L2"
                         % (h2d_bytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "iou": line_iou,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
