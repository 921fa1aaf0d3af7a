"""End-to-end check of the host-side caller (s2anet_b200/head.py) on the GPU ops against the CPU
reference path of bench.py (torch CPU convs + torchvision deform_conv2d + ARF scatter + the
reference's CPU ml_nms extension / the oracle): same weights, same features, fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _feats(B, img, seed, device, dtype):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 256, img // s, img // s, generator=g).to(dtype).to(device) for s in (8, 16, 32, 64, 128)]


def test_head_fp32_matches_cpu_reference_path():
    from bench import CpuReferenceHead
    from s2anet_b200.head import S2ANetHead
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    head = S2ANetHead(15).eval()
    head.init_synthetic(3)
    feats_cpu = _feats(2, 256, 5, "cpu", torch.float32)
    gpu_head = S2ANetHead(15).eval()
    gpu_head.load_state_dict(head.state_dict())
    gpu_head = gpu_head.to(DEV)
    feats_gpu = [f.to(DEV) for f in feats_cpu]
    n = gpu_head.calibrate_scores(feats_gpu, 1500)
    assert 800 < n < 2500
    head.odm_cls_head.bias.data.copy_(gpu_head.odm_cls_head.bias.detach().cpu())
    ref = CpuReferenceHead(head)
    # level outputs first (AlignConv / ORConv / towers), tight tolerance
    outs_gpu = gpu_head.forward_levels(feats_gpu)
    outs_cpu = [ref.level(x, s) for x, s in zip(feats_cpu, (8, 16, 32, 64, 128))]
    for og, oc in zip(outs_gpu, outs_cpu):
        for k in (1, 2, 3):      # fam_bbox_pred, odm_cls_pred, odm_bbox_pred
            a, b = og[k].float().cpu(), oc[k]
            assert float((a - b).abs().max()) <= 2e-3 * (1.0 + float(b.abs().max())), k
        assert float((og[5].cpu() - oc[5]).abs().max()) <= 1e-2      # refined anchors (pixels / radians)
    # detections: same count up to threshold flips of near-tied scores, same boxes for the confident ones
    res_gpu = gpu_head.get_bboxes(feats_gpu)
    res_cpu = ref.detect(feats_cpu)
    for (dg, lg), (dc, lc) in zip(res_gpu, res_cpu):
        dg, lg, dc, lc = dg.cpu(), lg.cpu(), dc, lc
        assert abs(dg.shape[0] - dc.shape[0]) <= max(3, int(0.02 * dc.shape[0]))
        k = min(100, dg.shape[0], dc.shape[0])
        assert k > 10
        # every one of the top-k has a counterpart: same label, score within 1e-3, box within 0.05 px (two detections of
        # one class can have scores closer than the two paths agree, so the nearest score alone is not the match)
        for i in range(k):
            cand = ((dc[:, 5] - dg[i, 5]).abs() <= 1e-3) & (lc.view(-1) == lg[i])
            assert bool(cand.any()), i
            assert float((dc[cand, :4] - dg[i, :4]).abs().max(dim=1)[0].min()) <= 0.05, i


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_head_16bit_runs_multilevel_and_agrees_with_fp32(dtype):
    """bf16/fp16 head (channels_last, multi-level tcgen05 launches) vs the fp32 head: the ODM outputs
    agree within 16-bit tolerance and the detection counts are of the same order."""
    from s2anet_b200.head import S2ANetHead
    head32 = S2ANetHead(15).eval()
    head32.init_synthetic(1)
    head32 = head32.to(DEV)
    feats32 = _feats(2, 512, 9, DEV, torch.float32)
    head32.calibrate_scores(feats32, 2000)
    head16 = S2ANetHead(15).eval()
    head16.load_state_dict(head32.state_dict())
    head16 = head16.to(DEV).to(dtype)
    feats16 = [f.to(dtype).contiguous(memory_format=torch.channels_last) for f in feats32]
    o32 = head32.forward_levels(feats32)
    o16 = head16.forward_levels(feats16)
    tol = 0.12 if dtype == torch.bfloat16 else 0.03
    for a, b in zip(o16, o32):
        for k in (2, 3):
            rel = float((a[k].float() - b[k]).norm() / (b[k].norm() + 1e-9))
            assert rel < tol, (k, rel)
    d16, l16, c16 = head16.detect(feats16)
    d32, l32, c32 = head32.detect(feats32)
    assert tuple(d16.shape) == (2, 2000, 6)
    for i in range(2):
        assert int(c16[i]) > 0 and abs(int(c16[i]) - int(c32[i])) <= 0.3 * int(c32[i]) + 20
        k = int(c16[i])
        assert bool((d16[i, 1:k, 5] <= d16[i, : k - 1, 5]).all())            # descending scores
