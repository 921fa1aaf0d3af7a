"""Deformable-convolution backward: DeformConvFunction.backward and the extension-level
deform_conv_backward_input_cuda / deform_conv_backward_parameters_cuda.

Witnesses: torchvision.ops.deform_conv2d's autograd (same operator, fp32/fp64 on the GPU and on the CPU) and, when
prebuilt, the reference's own deform_conv_cuda extension (unmodified sources compiled for sm_100a) running the
same three calls.  Tolerance: 1e-4 + 1e-4 * max|ref| in fp32 (summation order differs: atomics, GEMM tiling)."""
import numpy as np
import pytest
import torch

from s2anet_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, what):
    err = float((a - b).abs().max())
    tol = 1e-4 + 1e-4 * float(b.abs().max())
    assert err <= tol, "%s: max-abs err %g > %g" % (what, err, tol)


CASES = [
    # B, C, H, W, Co, k, stride, pad, dil, groups, dgroups
    (2, 16, 12, 10, 24, 3, 1, 1, 1, 1, 1),
    (3, 16, 9, 11, 12, 3, 2, 2, 2, 1, 2),
    (1, 16, 7, 8, 128, 3, 1, 1, 1, 2, 1),
    (1, 4, 6, 6, 4, 1, 1, 0, 1, 1, 1),
]


@pytest.mark.parametrize("B,C,H,W,Co,k,stride,pad,dil,groups,dgroups", CASES)
def test_autograd_matches_torchvision(B, C, H, W, Co, k, stride, pad, dil, groups, dgroups):
    import torchvision
    from s2anet_b200.dcn import deform_conv
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(B * 100 + C)
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    x = torch.randn(B, C, H, W, generator=g).to(DEV).requires_grad_()
    off = (torch.randn(B, dgroups * 2 * k * k, Ho, Wo, generator=g) * 1.5).to(DEV).requires_grad_()
    w = (torch.randn(Co, C // groups, k, k, generator=g) * 0.2).to(DEV).requires_grad_()
    gy = torch.randn(B, Co, Ho, Wo, generator=g).to(DEV)
    y = deform_conv(x, off, w, stride, pad, dil, groups, dgroups)
    y.backward(gy)
    x2, off2, w2 = (t_.detach().double().requires_grad_() for t_ in (x, off, w))
    y2 = torchvision.ops.deform_conv2d(x2, off2, w2, stride=stride, padding=pad, dilation=dil)
    y2.backward(gy.double())
    close(y.detach().double(), y2.detach(), "forward")
    close(x.grad.double(), x2.grad, "grad_input")
    close(off.grad.double(), off2.grad, "grad_offset")
    close(w.grad.double(), w2.grad, "grad_weight")


def test_alignconv_training_step_p4_vs_torchvision_and_reference_cuda():
    """AlignConv at P4 size (64x64, 256 -> 256, batch 2) with autograd: gradients of the feature map and of the
    weight against torchvision, and the three extension-level calls against the reference's own CUDA binary."""
    import torchvision
    from oracle import build_oracle
    from s2anet_b200 import dcn
    from s2anet_b200.alignconv import AlignConv
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(3)
    B, C, H = 2, 256, 64
    x = torch.randn(B, C, H, H, generator=g).to(DEV).requires_grad_()
    anc = torch.from_numpy(synth.refined_anchors(B, H, H, 16, seed=4)).to(DEV)
    ac = AlignConv(C, C).to(DEV)
    ac.init_weights()
    y = ac(x, anc, 16)
    assert y.requires_grad
    gy = torch.randn(y.shape, generator=g).to(DEV)
    y.backward(gy)
    off = torch.stack([ac.get_offset(anc[i].reshape(-1, 5), (H, H), 16) for i in range(B)])
    x2 = x.detach().clone().requires_grad_()
    w2 = ac.deform_conv.weight.detach().clone().requires_grad_()
    pre2 = torchvision.ops.deform_conv2d(x2, off, w2, padding=1)
    # the ReLU mask is taken from OUR forward: an output within rounding of zero (the two forwards differ by ~1e-6
    # relative -- summation order, 3 x TF32 split) may land on either side of it, and one flipped element moves
    # grad_input by a whole |gy * w| term.  Flips are allowed only where the pre-activation is that small.
    mask = y.detach() > 0
    flips = mask != (pre2.detach() > 0)
    assert int(flips.sum()) <= 8 and float(pre2.detach()[flips].abs().max() if flips.any() else 0.0) <= 1e-5
    y2 = pre2 * mask
    y2.backward(gy)
    close(y.detach(), y2.detach(), "forward")
    close(x.grad, x2.grad, "grad_input")
    close(ac.deform_conv.weight.grad, w2.grad, "grad_weight")
    ref = build_oracle.load_ref_extension("deform_conv_cuda", "gpu")
    if ref is None:
        pytest.skip("oracle/_ref/ext_gpu/deform_conv_cuda not prebuilt")
    w = ac.deform_conv.weight.detach()
    xd = x.detach()
    gpre = gy * (y.detach() > 0)                       # gradient in front of the ReLU
    outs = []
    for ext in (ref, dcn):
        gi, gof, gw = torch.zeros_like(xd), torch.zeros_like(off), torch.zeros_like(w)
        ext.deform_conv_backward_input_cuda(xd, off, gpre, gi, gof, w, xd.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
        ext.deform_conv_backward_parameters_cuda(xd, off, gpre, gw, xd.new_empty(0), xd.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1,
                                                 1, 1, 1.0, B)
        torch.cuda.synchronize()
        outs.append((gi, gof, gw))
    for a, b, name in zip(outs[1], outs[0], ("grad_input", "grad_offset", "grad_weight")):
        close(a, b, name + " vs the reference CUDA extension")


def test_extension_entries_accumulate_and_check_arguments():
    from s2anet_b200 import dcn
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 5, 5, generator=g).to(DEV)
    off = torch.randn(2, 18, 5, 5, generator=g).to(DEV)
    w = torch.randn(3, 4, 3, 3, generator=g).to(DEV)
    gy = torch.randn(2, 3, 5, 5, generator=g).to(DEV)
    gi, gof, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
    e = x.new_empty(0)
    assert dcn.deform_conv_backward_input_cuda(x, off, gy, gi, gof, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2) == 1
    assert dcn.deform_conv_backward_parameters_cuda(x, off, gy, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0.5, 1) == 1
    gi2, gof2, gw2 = gi.clone(), gof.clone(), gw.clone()
    dcn.deform_conv_backward_input_cuda(x, off, gy, gi2, gof2, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1)
    dcn.deform_conv_backward_parameters_cuda(x, off, gy, gw2, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0.5, 2)
    close(gi2, 2 * gi, "accumulated grad_input")          # the entries ADD to what the caller passes in
    close(gof2, 2 * gof, "accumulated grad_offset")
    close(gw2, 2 * gw, "accumulated grad_weight (scale 0.5 twice)")
    with pytest.raises(RuntimeError):
        dcn.deform_conv_backward_input_cuda(x, off, gy, gi, gof, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 3)    # 3 does not divide 2
    with pytest.raises(NotImplementedError):
        dcn.deform_conv_backward_input_cuda(x.cpu(), off.cpu(), gy.cpu(), gi.cpu(), gof.cpu(), w.cpu(), e.cpu(), 3, 3, 1, 1, 1,
                                            1, 1, 1, 1, 1, 1)


# ---- 16-bit backward on tcgen05 (round 2): dgrad = nine 1x1 implicit GEMMs with a scatter epilogue ----------------------

def errs16(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12)), float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("B,C,H,W,Co", [(1, 64, 9, 13, 64), (2, 256, 24, 40, 256), (3, 32, 5, 3, 128)])
def test_dgrad_tc_vs_fp32_kernels(dtype, B, C, H, W, Co):
    """grad_input and grad_offset of the tcgen05 dgrad against this repo's fp32 backward (itself checked against
    torchvision autograd and the reference binary above) fed with the SAME 16-bit-rounded operands.  Differences:
    the column gradient's fp32 accumulation order, and x / grad_out / W enter the GEMM as 16-bit values (they ARE
    16-bit values here).  Stated tolerance: bf16 / fp16 max-abs <= 2e-3 of max, rel-L2 <= 2e-3 (fp32 accumulation on
    both sides; only the atomics' order differs)."""
    from s2anet_b200 import dcn
    from s2anet_b200.conv_tc import deform_conv_dgrad_tc
    g = torch.Generator().manual_seed(C + H + B)
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(dtype)
    w = (torch.randn(Co, C, 3, 3, generator=g) * 0.05).to(DEV).to(dtype)
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).to(DEV)
    off[0, :, 0, 0] = 50.0                                   # a pixel whose samples all fall outside the map
    gy = torch.randn(B, Co, H, W, generator=g).to(DEV).to(dtype)
    gi, goff = deform_conv_dgrad_tc(gy, off, w, x=x, need_offset_grad=True)
    assert gi.dtype == torch.float32 and tuple(gi.shape) == (B, C, H, W) and tuple(goff.shape) == (B, 18, H, W)
    gi_r, goff_r = torch.zeros(B, C, H, W, device=DEV), torch.zeros_like(off)
    e = gi_r.new_empty(0)
    dcn.deform_conv_backward_input_cuda(x.float(), off, gy.float(), gi_r, goff_r, w.float(), e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
    for name, a, b in (("grad_input", gi, gi_r), ("grad_offset", goff, goff_r)):
        mx, l2 = errs16(a, b)
        assert mx <= 2e-3 and l2 <= 2e-3, (name, mx, l2)
    gi2, none = deform_conv_dgrad_tc(gy, off.to(dtype), w)       # 16-bit offsets, no offset gradient
    assert none is None
    gi_r2, goff_r2 = torch.zeros_like(gi_r), torch.zeros_like(off)
    dcn.deform_conv_backward_input_cuda(x.float(), off.to(dtype).float(), gy.float(), gi_r2, goff_r2, w.float(), e, 3, 3, 1, 1, 1,
                                        1, 1, 1, 1, 1, B)
    mx, l2 = errs16(gi2, gi_r2)
    assert mx <= 2e-3 and l2 <= 2e-3, (mx, l2)


def test_half_precision_training_step_uses_tc_dgrad_and_matches_reference_binary():
    """The reference's fp16 route end to end: DeformConvFunction (this repo's) forward + backward on half tensors at
    P4 size -- forward and dgrad on tcgen05 -- against the reference's deform_conv_cuda binary run in half, and the
    extension-level `deform_conv_backward_input_cuda` shim with half tensors against the same binary."""
    from oracle import build_oracle
    from s2anet_b200 import _lib, dcn
    ref = build_oracle.load_ref_extension("deform_conv_cuda", "gpu")
    if ref is None:
        pytest.skip("oracle/_ref/ext_gpu/deform_conv_cuda not prebuilt")
    g = torch.Generator().manual_seed(6)
    B, C, H = 2, 256, 64
    # smooth features: the half kernels round sampling positions to 1/32 - 1/16 px
    x = torch.nn.functional.interpolate(torch.randn(B, C, H // 4, H // 4, generator=g), size=(H, H), mode="bilinear").to(DEV).half()
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.02).to(DEV).half()
    off = (torch.randn(B, 18, H, H, generator=g) * 1.2).to(DEV).half()
    gy = torch.randn(B, C, H, H, generator=g).to(DEV).half()
    xr = x.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    before = _lib.launches
    y = dcn.deform_conv(xr, off, wr, 1, 1, 1, 1, 1)
    y.backward(gy)
    assert _lib.launches - before >= 10                       # 1 forward + 9 dgrad launches of this library (+ wgrad path)
    e = x.new_empty(0)
    gi_ref, gof_ref, gw_ref = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
    ref.deform_conv_backward_input_cuda(x, off, gy, gi_ref, gof_ref, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
    ref.deform_conv_backward_parameters_cuda(x, off, gy, gw_ref, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
    # ... and the same binary in fp32 on the same half-valued operands: the yardstick.  (Its half run accumulates
    # grad_input with half-precision atomics and rounds the sampling positions to half; it is itself ~1e-2 away.)
    x32, off32, gy32, w32 = x.float(), off.float(), gy.float(), w.float()
    e32 = x32.new_empty(0)
    gi_r32, gof_r32, gw_r32 = torch.zeros_like(x32), torch.zeros_like(off32), torch.zeros_like(w32)
    ref.deform_conv_backward_input_cuda(x32, off32, gy32, gi_r32, gof_r32, w32, e32, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
    ref.deform_conv_backward_parameters_cuda(x32, off32, gy32, gw_r32, e32, e32, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
    torch.cuda.synchronize()
    ref_half_err = errs16(gi_ref, gi_r32)[1]
    mx, l2 = errs16(xr.grad, gi_r32)
    assert xr.grad.dtype == torch.float16 and l2 <= 2e-3, ("grad_input vs the reference binary in fp32", mx, l2)
    assert l2 <= ref_half_err + 1e-4                          # at least as close to fp32 as the reference's own half run
    assert errs16(xr.grad, gi_ref)[1] <= 5e-2
    mx, l2 = errs16(wr.grad, gw_r32)
    assert l2 <= 2e-3, ("grad_weight vs the reference binary in fp32", mx, l2)
    gi_s, gof_s = torch.zeros_like(x), torch.zeros_like(off)
    assert dcn.deform_conv_backward_input_cuda(x, off, gy, gi_s, gof_s, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B) == 1
    mx, l2 = errs16(gi_s, gi_r32)
    assert l2 <= 2e-3, ("shim grad_input", mx, l2)
    mx, l2 = errs16(gof_s, gof_r32)
    assert l2 <= 5e-3, ("shim grad_offset", mx, l2)
    print("reference half backward vs its own fp32 run: rel-L2 %.2e; tcgen05 dgrad vs that fp32 run: %.2e" % (ref_half_err, errs16(xr.grad, gi_r32)[1]))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("B,C,H,W,Co", [(1, 128, 8, 16, 128), (2, 256, 24, 40, 256), (3, 128, 13, 7, 256), (1, 256, 64, 64, 256)])
def test_wgrad_tc_vs_fp32_kernels(dtype, B, C, H, W, Co):
    """grad_weight of the tcgen05 wgrad (MN-major operands, accumulator in tensor memory across all pixel tiles of a
    CTA, one tap per CTA) against this repo's fp32 path (im2col + GEMM, checked against torchvision autograd and the
    reference binary above) on the SAME 16-bit-rounded operands.  Difference: the bilinear samples are rounded to 16
    bits before the MMA (as in the forward) and the summation order.  Stated tolerance: bf16 max-abs <= 2e-2 of max,
    rel-L2 <= 5e-3; fp16 4e-3 / 1e-3."""
    from s2anet_b200 import dcn
    from s2anet_b200.conv_tc import deform_conv_wgrad_tc
    g = torch.Generator().manual_seed(C + H + B + Co)
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(dtype)
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).to(DEV)
    off[0, :, 0, 0] = -40.0
    gy = torch.randn(B, Co, H, W, generator=g).to(DEV).to(dtype)
    dw = deform_conv_wgrad_tc(x, off, gy)
    assert dw.dtype == torch.float32 and tuple(dw.shape) == (Co, C, 3, 3)
    ref = torch.zeros(Co, C, 3, 3, device=DEV)
    e = ref.new_empty(0)
    dcn.deform_conv_backward_parameters_cuda(x.float(), off, gy.float(), ref, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
    mx, l2 = errs16(dw, ref)
    amax, arel = (2e-2, 5e-3) if dtype == torch.bfloat16 else (4e-3, 1e-3)
    assert mx <= amax and l2 <= arel, (mx, l2)
    # accumulation + scale through the reference-shaped entry with 16-bit tensors
    gw = torch.ones(Co, C, 3, 3, device=DEV, dtype=dtype)
    e16 = x.new_empty(0)
    assert dcn.deform_conv_backward_parameters_cuda(x, off.to(dtype), gy, gw, e16, e16, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0.5, B) == 1
    ref2 = torch.zeros(Co, C, 3, 3, device=DEV)
    dcn.deform_conv_backward_parameters_cuda(x.float(), off.to(dtype).float(), gy.float(), ref2, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0.5, B)
    mx, l2 = errs16(gw.float() - 1.0, ref2)
    assert l2 <= 2e-2, (mx, l2)                       # (the 16-bit gradWeight tensor itself rounds the sum)
