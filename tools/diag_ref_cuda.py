"""Diagnostic (GPU box): where does the reference's own CUDA IoU kernel (FMA-contracted build)
differ from the oracle semantics, and who is closer to a float64 polygon clipper?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import build_oracle, oracle as O
from s2anet_b200 import synth
from s2anet_b200.box_iou_rotated import box_iou_rotated, box_iou_rotated_batched


def poly(b):
    x, y, w, h, a = [float(v) for v in b]
    c, s = np.cos(a), np.sin(a)
    pts = []
    for dx, dy in ((-w / 2, -h / 2), (w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2)):
        pts.append((x + dx * c - dy * s, y + dx * s + dy * c))
    return pts


def clip(subject, clipper):
    def inside(p, a, b):
        return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0
    def inter(p1, p2, a, b):
        x1, y1, x2, y2 = *p1, *p2
        x3, y3, x4, y4 = *a, *b
        d = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / d
        return (x1 + t * (x2 - x1), y1 + t * (y2 - y1))
    out = subject
    for i in range(len(clipper)):
        a, b = clipper[i], clipper[(i + 1) % len(clipper)]
        inp, out = out, []
        if not inp:
            break
        s = inp[-1]
        for e in inp:
            if inside(e, a, b):
                if not inside(s, a, b):
                    out.append(inter(s, e, a, b))
                out.append(e)
            elif inside(s, a, b):
                out.append(inter(s, e, a, b))
            s = e
    return out


def area(p):
    return 0.5 * abs(sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p))))


def iou64(b1, b2):
    p1, p2 = poly(b1), poly(b2)
    it = clip(p1, p2)
    ia = area(it) if len(it) >= 3 else 0.0
    return ia / (area(p1) + area(p2) - ia)


dev = "cuda:0"
ref_ext = build_oracle.load_ref_extension("box_iou_rotated_cuda", "gpu")
b, s, l = synth.clustered_boxes(seed=0)
tb = torch.from_numpy(b).to(dev)
mine = box_iou_rotated(tb, tb).cpu().numpy()
ref = ref_ext.box_iou_rotated(tb, tb).cpu().numpy()
d = np.abs(mine - ref)
print("pairs:", d.size, " >1e-5:", (d > 1e-5).sum(), " >1e-6:", (d > 1e-6).sum(), " >1e-7:", (d > 1e-7).sum(),
      " bit-different:", (mine.view(np.uint32) != ref.view(np.uint32)).sum())
idx = np.argwhere(d > 1e-5)
for i, j in idx[:20]:
    print(i, j, "mine", mine[i, j], "ref_cuda", ref[i, j], "oracle", O.single_iou(b[i], b[j]), "f64", iou64(b[i], b[j]))
    print("   box1", b[i].tolist(), "box2", b[j].tolist())
an = torch.from_numpy(synth.all_level_anchors(1, 3)[0]).to(dev)
gt = torch.from_numpy(synth.dota_like_gt(500, 3)).to(dev)
m2, r2 = box_iou_rotated(an, gt).cpu().numpy(), ref_ext.box_iou_rotated(an, gt).cpu().numpy()
d2 = np.abs(m2 - r2)
print("anchor x gt pairs:", d2.size, " >1e-5:", (d2 > 1e-5).sum(), " >1e-6:", (d2 > 1e-6).sum(), "max", d2.max())
for i, j in np.argwhere(d2 > 1e-5)[:10]:
    a_, g_ = an[i].cpu().numpy(), gt[j].cpu().numpy()
    print(i, j, "mine", m2[i, j], "ref_cuda", r2[i, j], "f64", iou64(a_, g_), a_.tolist(), g_.tolist())

# ---- first timings -------------------------------------------------------------------------------
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = 16
anB = torch.from_numpy(synth.all_level_anchors(B, 3)).to(dev)
gtB = torch.from_numpy(np.stack([synth.dota_like_gt(500, 100 + i) for i in range(B)])).to(dev)
out = torch.empty((B, anB.shape[1], 500), device=dev)
ms = timeit(lambda: box_iou_rotated_batched(anB, gtB, out=out))
pairs = B * anB.shape[1] * 500
print("IoU batched B=16: %.3f ms  %.1f Gpairs/s  out %.1f GB/s" % (ms, pairs / ms / 1e6, pairs * 4 / ms / 1e6))
ms = timeit(lambda: box_iou_rotated_batched(anB, gtB, out=out, _flags=1), n=3, warm=1)
print("IoU batched no-reject: %.3f ms  %.2f Gpairs/s" % (ms, pairs / ms / 1e6))
ms = timeit(lambda: ref_ext.box_iou_rotated(anB[0], gtB[0]))
print("reference CUDA kernel, 1 image: %.3f ms  %.2f Gpairs/s" % (ms, anB.shape[1] * 500 / ms / 1e6))
from s2anet_b200.nms_rotated import nms_rotated_op, multiclass_nms_rotated_batched
ts = torch.from_numpy(s).to(dev)
ms = timeit(lambda: nms_rotated_op(tb, ts, 0.5))
print("nms_rotated N=2000: %.3f ms" % ms)
nms_ref = build_oracle.load_ref_extension("nms_rotated_cuda", "gpu")
ms = timeit(lambda: nms_ref.nms_rotated(tb, ts, 0.5))
print("reference nms_rotated CUDA N=2000: %.3f ms" % ms)
b20, s20, _ = synth.clustered_boxes(n_seed=4000, rep=5, seed=21)
tb20, ts20 = torch.from_numpy(b20).to(dev), torch.from_numpy(s20).to(dev)
ms = timeit(lambda: nms_rotated_op(tb20, ts20, 0.5), n=3, warm=1)
print("nms_rotated N=20000: %.3f ms" % ms)
ms = timeit(lambda: nms_ref.nms_rotated(tb20, ts20, 0.5), n=3, warm=1)
print("reference nms_rotated CUDA N=20000: %.3f ms" % ms)
rng = np.random.default_rng(6)
bb, _, _ = synth.clustered_boxes(n_seed=1336, rep=4, seed=44)
sc = (rng.uniform(0, 1, (bb.shape[0], 15)) ** 8).astype(np.float32)
tbb = torch.from_numpy(bb).to(dev)[None].repeat(8, 1, 1); tsc = torch.from_numpy(sc).to(dev)[None].repeat(8, 1, 1)
print("candidates/img:", int((sc > 0.05).sum()))
ms = timeit(lambda: multiclass_nms_rotated_batched(tbb, tsc))
print("multiclass batched B=8 n=5344: %.3f ms" % ms)
ms = timeit(lambda: multiclass_nms_rotated_batched(tbb[:1], tsc[:1]))
print("multiclass batched B=1 n=5344: %.3f ms" % ms)
from s2anet_b200.alignconv import alignconv_forward
from s2anet_b200.orn import orconv_forward
x = torch.randn(1, 256, 128, 128, device=dev); anc = torch.from_numpy(synth.refined_anchors(1, 128, 128, 8, 1)).to(dev)
w = torch.randn(256, 256, 3, 3, device=dev) * 0.01
ms = timeit(lambda: alignconv_forward(x, anc, w, 8))
print("alignconv fp32 P3: %.3f ms  %.1f TFLOP/s" % (ms, 19.327 / ms))
wo = torch.randn(32, 256, 1, 3, 3, device=dev) * 0.01; idx = torch.from_numpy(O.arf_indices(1, 8, 3)).to(dev)
ms = timeit(lambda: orconv_forward(x, wo, idx, None, with_pool=True))
print("orconv fp32 P3: %.3f ms  %.1f TFLOP/s" % (ms, 19.327 / ms))
dc = build_oracle.load_ref_extension("deform_conv_cuda", "gpu")
import torchvision
off = torch.zeros(1, 18, 128, 128, device=dev); outb = torch.empty(1, 256, 128, 128, device=dev)
ms = timeit(lambda: dc.deform_conv_forward_cuda(x, w, off, outb, x.new_empty(0), x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1))
print("reference deform_conv_cuda fp32 P3: %.3f ms" % ms)

xh, wh, offh = x.half(), w.half(), off.half(); outh = torch.empty(1, 256, 128, 128, device=dev, dtype=torch.half)
ms = timeit(lambda: dc.deform_conv_forward_cuda(xh, wh, offh, outh, xh.new_empty(0), xh.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1))
print("reference deform_conv_cuda fp16 P3 B=1: %.3f ms" % ms)
x8 = torch.randn(8, 256, 128, 128, device=dev).half(); off8 = torch.zeros(8, 18, 128, 128, device=dev).half(); out8 = torch.empty(8, 256, 128, 128, device=dev, dtype=torch.half)
ms = timeit(lambda: dc.deform_conv_forward_cuda(x8, wh, off8, out8, xh.new_empty(0), xh.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 8))
print("reference deform_conv_cuda fp16 P3 B=8: %.3f ms" % ms)
xcl = x8.contiguous(memory_format=torch.channels_last); a8 = torch.from_numpy(synth.refined_anchors(8, 128, 128, 8, 1)).to(dev)
ms = timeit(lambda: alignconv_forward(xcl, a8, wh, 8))
print("s2a alignconv fp16 P3 B=8 (incl. offset generation): %.3f ms" % ms)
import torch.nn.functional as F
ms = timeit(lambda: F.conv2d(xcl, w.half().contiguous(memory_format=torch.channels_last), padding=1))
print("cuDNN conv fp16 P3 B=8: %.3f ms" % ms)
