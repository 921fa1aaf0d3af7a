"""GPU parity tests (-m gpu) for the tcgen05 (bf16 / fp16) AlignConv and ORConv2d kernels.

Reference for these floating-point kernels = this repo's fp32 kernels (themselves checked against
the oracle / torchvision / the reference CUDA op in test_gpu_conv.py) fed with the SAME 16-bit
rounded inputs and weights, so the only differences are (a) the blended A operand is rounded to
16 bits before the MMA and (b) summation order.  Stated tolerance (north_star "bf16/fp32 tolerance,
max-abs and relative error stated"):  bf16: max-abs <= 2e-2 * max|ref|, relative L2 <= 5e-3;
fp16: max-abs <= 4e-3 * max|ref|, relative L2 <= 1e-3.
"""
import numpy as np
import pytest
import torch

from s2anet_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {torch.bfloat16: (2e-2, 5e-3), torch.float16: (4e-3, 1e-3)}


def check(y, ref, dtype):
    y, ref = y.float(), ref.float()
    mx = float(ref.abs().max()) + 1e-12
    err = float((y - ref).abs().max())
    rel = float((y - ref).norm() / (ref.norm() + 1e-12))
    amax, arel = TOL[dtype]
    assert err <= amax * mx and rel <= arel, "max-abs %g (ref max %g), rel-L2 %g" % (err, mx, rel)
    return err / mx, rel


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,C,H,W,Co,stride", [(1, 64, 8, 16, 32, 8), (2, 128, 13, 21, 256, 16), (1, 256, 8, 8, 256, 128),
                                               (3, 64, 5, 3, 64, 64)])
def test_alignconv_tc_vs_fp32(dtype, B, C, H, W, Co, stride):
    from s2anet_b200.alignconv import alignconv_forward
    g = torch.Generator().manual_seed(C + H + W)
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(dtype)
    w = (torch.randn(Co, C, 3, 3, generator=g) * 0.05).to(DEV).to(dtype)
    anc = torch.from_numpy(synth.refined_anchors(B, H, W, stride, seed=H)).to(DEV)
    y = alignconv_forward(x.contiguous(memory_format=torch.channels_last), anc, w, stride)
    assert y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
    ref = alignconv_forward(x.float(), anc, w.float(), stride)
    check(y, ref, dtype)
    # NCHW-contiguous input is accepted too (converted once)
    y2 = alignconv_forward(x.contiguous(), anc, w, stride)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,I,H,W,O", [(1, 64, 8, 16, 4), (2, 256, 11, 19, 32), (1, 128, 3, 5, 8)])
def test_orconv_tc_vs_fp32(dtype, B, I, H, W, O):
    from s2anet_b200.orn import ORConv2d, orconv_forward
    m = ORConv2d(I, O, 3, padding=1, arf_config=(1, 8)).to(DEV)
    g = torch.Generator().manual_seed(I + H)
    with torch.no_grad():
        m.weight.copy_((torch.randn(m.weight.shape, generator=g) * 0.05).to(dtype).float())
        m.bias.copy_(torch.randn(O * 8, generator=g) * 0.1)
    x = torch.randn(B, I, H, W, generator=g).to(DEV).to(dtype)
    y, yp = orconv_forward(x, m.weight.to(dtype), m.indices, m.bias, with_pool=True)
    ref, refp = orconv_forward(x.float(), m.weight, m.indices, m.bias, with_pool=True)
    check(y, ref, dtype)
    check(yp, refp, dtype)
    # fused pooling == pooling of the rounded output
    assert torch.equal(yp, y.view(B, O, 8, H, W).max(dim=2)[0])


def test_p3_full_size_tc_and_head_chain():
    """BASELINE config 2 shape in bf16: AlignConv -> ORConv2d(+pool) chained in channels_last."""
    from s2anet_b200.alignconv import AlignConv
    from s2anet_b200.orn import ORConv2d, RotationInvariantPooling
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 256, 128, 128, generator=g).to(DEV)
    anc = torch.from_numpy(synth.refined_anchors(1, 128, 128, 8, seed=1)).to(DEV)
    ac = AlignConv(256, 256).to(DEV)
    ac.init_weights()
    oc = ORConv2d(256, 32, 3, padding=1, arf_config=(1, 8)).to(DEV)
    torch.nn.init.normal_(oc.weight, 0, 0.01)
    oc.fuse_pool = True
    pool = RotationInvariantPooling(256, 8)
    with torch.no_grad():
        ac16, oc16 = ac.to(dt), oc.to(dt)
        xa = x.to(dt).contiguous(memory_format=torch.channels_last)
        ya = ac16(xa, anc, 8)
        yo = oc16(ya)
        yp = pool(yo)
    assert tuple(ya.shape) == (1, 256, 128, 128) and tuple(yo.shape) == (1, 256, 128, 128) and tuple(yp.shape) == (1, 32, 128, 128)
    # references in fp32 from the same 16-bit weights
    ac32 = AlignConv(256, 256).to(DEV)
    oc32 = ORConv2d(256, 32, 3, padding=1, arf_config=(1, 8)).to(DEV)
    with torch.no_grad():
        ac32.deform_conv.weight.copy_(ac16.deform_conv.weight.float())
        oc32.weight.copy_(oc16.weight.float())
        oc32.bias.copy_(oc16.bias.float())
        ra = ac32(x.to(dt).float(), anc, 8)
        check(ya, ra, dt)
        ro = oc32(ya.float())
        check(yo, ro, dt)
    assert torch.equal(yp, yo.view(1, 32, 8, 128, 128).max(dim=2)[0])


def test_multi_level_launch_equals_per_level():
    """Five FPN levels of a batch in one persistent launch == five single-level launches, bit for bit."""
    from s2anet_b200.conv_tc import alignconv_forward_tc, alignconv_forward_tc_multi, orconv_forward_tc, orconv_forward_tc_multi
    from s2anet_b200.orn import ORConv2d
    dt = torch.bfloat16
    B, C = 3, 128
    g = torch.Generator().manual_seed(7)
    strides = (8, 16, 32, 64, 128)
    sizes = ((40, 24), (20, 12), (10, 6), (5, 3), (3, 2))
    xs = [torch.randn(B, C, h, w, generator=g).to(DEV).to(dt).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    ancs = [torch.from_numpy(synth.refined_anchors(B, h, w, s, seed=h)).to(DEV) for (h, w), s in zip(sizes, strides)]
    w = (torch.randn(256, C, 3, 3, generator=g) * 0.05).to(DEV).to(dt)
    multi = alignconv_forward_tc_multi(xs, ancs, w, strides)
    for x, a, s, y in zip(xs, ancs, strides, multi):
        assert torch.equal(y, alignconv_forward_tc(x, a, w, s))
    m = ORConv2d(C, 32, 3, padding=1, arf_config=(1, 8)).to(DEV).to(dt)
    outs, pooled = orconv_forward_tc_multi(xs, m.weight, m.indices, m.bias, with_pool=True)
    for x, y, yp in zip(xs, outs, pooled):
        r, rp = orconv_forward_tc(x, m.weight, m.indices, m.bias, with_pool=True)
        assert torch.equal(y, r) and torch.equal(yp, rp)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,Co,ks,relu", [(256, 256, 3, True), (32, 256, 3, True), (256, 15, 3, False), (256, 5, 1, False),
                                          (64, 32, 1, True), (72, 40, 3, False), (512, 128, 3, True), (320, 32, 3, False),
                                          (512, 96, 1, False)])
def test_conv2d_tc_multi_vs_torch(dtype, C, Co, ks, relu):
    """The stock nn.Conv2d layers of the head on the tcgen05 kernel: kernel 1 / 3, channel counts that need
    zero-padded k-blocks (32, 72) and padded outputs (5, 15, 40), several levels incl. partial tiles; 5 to 8 channel
    blocks per tile so that the operand-halo ring (2 buffers; 4 in the narrow and the 1 x 1 mode) wraps many times."""
    from s2anet_b200.conv_tc import conv2d_forward_tc_multi
    g = torch.Generator().manual_seed(C + Co + ks)
    sizes = ((24, 40), (13, 21), (3, 5))
    xs = [torch.randn(2, C, h, w, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    wt = (torch.randn(Co, C, ks, ks, generator=g) * 0.05).to(DEV).to(dtype)
    bias = (torch.randn(Co, generator=g) * 0.1).to(DEV)
    ys = conv2d_forward_tc_multi(xs, wt, bias, relu=relu)
    for x, y in zip(xs, ys):
        assert tuple(y.shape) == (2, Co, x.size(2), x.size(3)) and y.dtype == dtype
        ref = torch.nn.functional.conv2d(x.float(), wt.float(), bias, padding=ks // 2)
        if relu:
            ref = ref.relu()
        check(y, ref, dtype)
    # no bias
    y0 = conv2d_forward_tc_multi(xs[:1], wt, None, relu=relu)[0]
    ref0 = torch.nn.functional.conv2d(xs[0].float(), wt.float(), None, padding=ks // 2)
    check(y0, ref0.relu() if relu else ref0, dtype)


@pytest.mark.parametrize("scale", [0.05, 1.0, 12.0])
def test_alignconv_tc_halo_and_global_fallback(scale):
    """The tcgen05 AlignConv reads the feature map through a 14 x 22-pixel shared-memory halo and falls back to
    global loads, per sample, for corners outside it.  Tiny anchors keep every sample in the halo, huge anchors
    (12x: tap offsets of ~16 pixels, many samples outside the map) send almost everything through the fallback."""
    from s2anet_b200.alignconv import alignconv_forward
    dtype = torch.bfloat16
    B, C, H, W, Co, stride = 2, 128, 40, 56, 64, 8
    g = torch.Generator().manual_seed(int(scale * 100))
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Co, C, 3, 3, generator=g) * 0.05).to(DEV).to(dtype)
    anc = synth.refined_anchors(B, H, W, stride, seed=3)
    anc[..., 2:4] *= scale
    anc[0, 0, 0] = [-5000.0, 9000.0, 300.0, 2.0, 0.3]          # far outside the image: contributes zeros
    anc = torch.from_numpy(anc).to(DEV)
    y = alignconv_forward(x, anc, w, stride)
    ref = alignconv_forward(x.float(), anc, w.float(), stride)
    check(y, ref, dtype)


@pytest.mark.parametrize("Co,ks", [(64, 3), (15, 3), (40, 1)])
def test_conv2d_tc_several_tiles_per_cta(Co, ks):
    """More tiles than CTAs (3 x 136 x 128 pixels = 408 tiles on 148 CTAs): the persistent loop carries the weight
    ring, the operand-halo ring and the two accumulators across tile boundaries, the last CTA pairs run a ghost tile."""
    from s2anet_b200.conv_tc import conv2d_forward_tc_multi
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(Co * 10 + ks)
    x = torch.randn(3, 128, 136, 128, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(Co, 128, ks, ks, generator=g) * 0.05).to(DEV).to(dtype)
    bias = torch.randn(Co, generator=g).to(DEV)
    y = conv2d_forward_tc_multi([x], wt, bias, relu=True)[0]
    ref = torch.relu(torch.nn.functional.conv2d(x.float(), wt.float(), bias, padding=ks // 2))
    check(y, ref, dtype)


def test_alignconv_tc_several_tiles_per_cta():
    from s2anet_b200.alignconv import alignconv_forward
    dtype = torch.bfloat16
    B, C, H, W, Co, stride = 3, 64, 128, 128, 64, 8
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, C, H, W, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Co, C, 3, 3, generator=g) * 0.05).to(DEV).to(dtype)
    anc = torch.from_numpy(synth.refined_anchors(B, H, W, stride, seed=4)).to(DEV)
    check(alignconv_forward(x, anc, w, stride), alignconv_forward(x.float(), anc, w.float(), stride), dtype)


@pytest.mark.parametrize("Co0,Co1,ks", [(256, 256, 3), (15, 5, 1), (15, 5, 3)])
def test_conv2d_tc_pair_equals_two_launches(Co0, Co1, ks):
    """Two convs of one shape class in ONE launch (levels >= split use the second weights / bias; the second problem
    starts at an even tile so a CTA pair never mixes weights -- odd tile counts get a padding tile): bit-identical to
    two separate launches, also when both read the same inputs and when the shapes force the two-launch fallback."""
    from s2anet_b200.conv_tc import conv2d_forward_tc_multi, conv2d_forward_tc_pair
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(Co0 + Co1 + ks)
    sizes = ((24, 40), (13, 21), (3, 5))                     # 2 * (10 + 4 + 1) = 15 tiles per problem: odd
    xs0 = [torch.randn(1, 128, h, w, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    xs1 = [torch.randn(1, 128, h, w, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last) for h, w in sizes]
    w0 = (torch.randn(Co0, 128, ks, ks, generator=g) * 0.05).to(DEV).to(dtype)
    w1 = (torch.randn(Co1, 128, ks, ks, generator=g) * 0.05).to(DEV).to(dtype)
    b0, b1 = torch.randn(Co0, generator=g).to(DEV), torch.randn(Co1, generator=g).to(DEV)
    for a, b in ((xs0, xs1), (xs0, xs0)):
        o0, o1 = conv2d_forward_tc_pair(a, w0, b0, b, w1, b1, relu=True)
        r0 = conv2d_forward_tc_multi(a, w0, b0, relu=True)
        r1 = conv2d_forward_tc_multi(b, w1, b1, relu=True)
        for x, y in zip(o0 + o1, r0 + r1):
            assert x.shape == y.shape and torch.equal(x, y)
    # different input channel counts: falls back to two launches, same results
    w2 = (torch.randn(Co1, 64, ks, ks, generator=g) * 0.05).to(DEV).to(dtype)
    xs2 = [x[:, :64].contiguous(memory_format=torch.channels_last) for x in xs1]
    o0, o2 = conv2d_forward_tc_pair(xs0, w0, b0, xs2, w2, b1, relu=False)
    for x, y in zip(o2, conv2d_forward_tc_multi(xs2, w2, b1, relu=False)):
        assert torch.equal(x, y)
