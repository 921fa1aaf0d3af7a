"""GPU parity tests (-m gpu) for AlignConv / DeformConv / ORConv2d / ARF / RotationInvariantPooling.

Floating-point bar (written here, from BASELINE.json north_star): fp32 path within
max-abs 1e-4 + 1e-4 * |ref| of the oracle / reference (accumulation order differs, nothing else).
"""
import numpy as np
import pytest
import torch

from s2anet_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL, ATOL = 1e-4, 1e-4


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_arf_forward_backward_exact(oracle, golden):
    from s2anet_b200.orn import active_rotating_filter, arf_backward, arf_forward
    g = golden("orconv_small.npz")
    out = arf_forward(t(g["weight"]), t(g["indices_1_8"]))
    np.testing.assert_array_equal(out.cpu().numpy(), g["rotated"])
    rng = np.random.default_rng(0)
    # S2ANet's shape (the reference's CPU kernel overflows uint16 here) and an 8-orientation bank
    for (O, I, nOri) in ((32, 256, 1), (3, 5, 8)):
        idx = oracle.arf_indices(nOri, 8, 3)
        w = rng.normal(size=(O, I, nOri, 3, 3)).astype(np.float32)
        np.testing.assert_array_equal(arf_forward(t(w), t(idx)).cpu().numpy(), oracle.arf_forward(w, idx))
        go = rng.normal(size=(O * 8, I * nOri, 3, 3)).astype(np.float32)
        np.testing.assert_allclose(arf_backward(t(idx), t(go)).cpu().numpy(), oracle.arf_backward(idx, go, O, I),
                                   rtol=1e-6, atol=1e-6)
    # half / bf16 are pure data movement
    for dt in (torch.float16, torch.bfloat16):
        w16 = t(w).to(dt)
        assert torch.equal(arf_forward(w16, t(idx)).float(), arf_forward(w16.float(), t(idx)))
    # autograd Function (reference: active_rotating_filter.py:12-33)
    wd = t(w).requires_grad_(True)
    active_rotating_filter(wd, t(idx)).backward(t(go))
    np.testing.assert_allclose(wd.grad.cpu().numpy(), oracle.arf_backward(idx, go, O, I), rtol=1e-6, atol=1e-6)


def test_ri_pool_exact(oracle):
    from s2anet_b200.orn import RotationInvariantPooling
    x = torch.randn(2, 256, 16, 24, device=DEV)
    out = RotationInvariantPooling(256, 8)(x)
    assert torch.equal(out, x.view(2, 32, 8, 16, 24).max(dim=2)[0])
    np.testing.assert_array_equal(out.cpu().numpy(), oracle.ri_pool(x.cpu().numpy(), 8))
    xb = x.to(torch.bfloat16)
    assert torch.equal(RotationInvariantPooling(256, 8)(xb), xb.view(2, 32, 8, 16, 24).max(dim=2)[0])


def test_alignconv_small_golden_and_oracle(oracle, golden):
    from s2anet_b200.alignconv import AlignConv
    g = golden("alignconv_small.npz")
    x, anc, w, stride = g["x"], g["anchors"], g["weight"], float(g["stride"])
    m = AlignConv(x.shape[1], w.shape[0]).to(DEV)
    with torch.no_grad():
        m.deform_conv.weight.copy_(t(w))
        y = m(t(x), t(anc), stride).cpu().numpy()
        off = torch.stack([m.get_offset(t(anc)[i].reshape(-1, 5), x.shape[2:], stride) for i in range(x.shape[0])])
    np.testing.assert_allclose(off.cpu().numpy(), g["offset_ref_py"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y, g["out_torchvision"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(y, oracle.alignconv_forward(x, anc, w, stride), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B,C,H,W,Co,stride", [(1, 8, 5, 7, 8, 8), (2, 24, 9, 9, 40, 16), (1, 64, 16, 16, 64, 32),
                                               (3, 16, 3, 4, 72, 128)])
def test_alignconv_shapes_vs_oracle(oracle, B, C, H, W, Co, stride):
    from s2anet_b200.alignconv import alignconv_forward
    rng = np.random.default_rng(C + H)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    anc = synth.refined_anchors(B, H, W, stride, seed=H)
    w = (rng.normal(size=(Co, C, 3, 3)) * 0.1).astype(np.float32)
    y = alignconv_forward(t(x), t(anc), t(w), stride).cpu().numpy()
    np.testing.assert_allclose(y, oracle.alignconv_forward(x, anc, w, stride), rtol=RTOL, atol=ATOL)
    assert y.min() >= 0.0                                                 # ReLU


def test_alignconv_degenerate_anchors(oracle):
    """anchors far outside the map, zero-size, huge: every sample falls outside -> exact zeros /
    plain-centre taps, same as the oracle."""
    from s2anet_b200.alignconv import alignconv_forward
    rng = np.random.default_rng(1)
    B, C, H, W, Co = 1, 8, 6, 6, 8
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    w = rng.normal(size=(Co, C, 3, 3)).astype(np.float32)
    anc = synth.refined_anchors(B, H, W, 8, seed=3)
    anc[0, 0, :, 0] = -1e4            # off the map
    anc[0, 1, :, 2:4] = 0.0           # zero size: all nine taps sample the centre
    anc[0, 2, :, 2:4] = 1e5           # huge
    y = alignconv_forward(t(x), t(anc), t(w), 8).cpu().numpy()
    ref = oracle.alignconv_forward(x, anc, w, 8)
    np.testing.assert_allclose(y, ref, rtol=RTOL, atol=ATOL)
    assert np.all(y[0, :, 0, :] == 0.0)


def test_deform_conv_generic_vs_golden_and_oracle(oracle, golden):
    from s2anet_b200.dcn import DeformConv, deform_conv
    g = golden("alignconv_small.npz")
    x, w, off2 = g["x"], g["weight"], g["offset2"]
    y2 = deform_conv(t(x), t(off2), t(w), 2, 2, 2, 1, 2)
    np.testing.assert_allclose(y2.cpu().numpy(), g["out2_torchvision"], rtol=RTOL, atol=ATOL)
    # module form, zero offsets == plain convolution
    m = DeformConv(16, 24, 3, padding=1).to(DEV)
    xx = torch.randn(2, 16, 10, 11, device=DEV)
    with torch.no_grad():
        y = m(xx, torch.zeros(2, 18, 10, 11, device=DEV))
        ref = torch.nn.functional.conv2d(xx, m.weight, padding=1)
    torch.testing.assert_close(y, ref, rtol=RTOL, atol=ATOL)
    # 1x1 kernel with stride, and the reference's "input smaller than kernel" padding branch
    rng = np.random.default_rng(3)
    x1 = rng.normal(size=(1, 8, 7, 5)).astype(np.float32)
    w1 = rng.normal(size=(8, 8, 1, 1)).astype(np.float32)
    o1 = rng.normal(size=(1, 2, 4, 3)).astype(np.float32)
    np.testing.assert_allclose(deform_conv(t(x1), t(o1), t(w1), 2).cpu().numpy(),
                               oracle.deform_conv_forward(x1, o1, w1, stride=(2, 2)), rtol=RTOL, atol=ATOL)
    tiny = DeformConv(8, 8, 3, padding=1).to(DEV)
    with torch.no_grad():
        yt = tiny(torch.randn(1, 8, 2, 2, device=DEV), torch.zeros(1, 18, 2, 2, device=DEV))
    assert tuple(yt.shape) == (1, 8, 2, 2)
    with pytest.raises(RuntimeError):
        deform_conv(t(x), t(off2[:, :10]), t(w), 2, 2, 2, 1, 2)            # wrong offset channels


def test_orconv_small_golden(oracle, golden):
    from s2anet_b200.orn import ORConv2d, RotationInvariantPooling
    g = golden("orconv_small.npz")
    O, I = g["weight"].shape[:2]
    m = ORConv2d(I, O, 3, padding=1, arf_config=(1, 8)).to(DEV)
    assert torch.equal(m.indices.cpu(), torch.from_numpy(g["indices_1_8"]))
    with torch.no_grad():
        m.weight.copy_(t(g["weight"]))
        m.bias.copy_(t(g["bias"]))
        y = m(t(g["x"]))
        np.testing.assert_allclose(y.cpu().numpy(), g["out"], rtol=RTOL, atol=ATOL)
        m.fuse_pool = True
        y2 = m(t(g["x"]))
        p = RotationInvariantPooling(O * 8, 8)(y2)
        np.testing.assert_allclose(p.cpu().numpy(), g["pooled"], rtol=RTOL, atol=ATOL)
        assert torch.equal(p, y2.view(2, -1, 8, 9, 7).max(dim=2)[0])      # fused pool == unfused pool
        np.testing.assert_array_equal(m.rotate_arf().cpu().numpy(), g["rotated"])


def test_orconv_eight_orientations_vs_oracle(oracle):
    from s2anet_b200.orn import ORConv2d
    rng = np.random.default_rng(9)
    m = ORConv2d(4, 3, 3, padding=1, arf_config=(8, 8)).to(DEV)          # weight [3,4,8,3,3], input 32 ch
    x = rng.normal(size=(2, 32, 6, 5)).astype(np.float32)
    with torch.no_grad():
        m.bias.copy_(torch.randn(24))
        y = m(t(x)).cpu().numpy()
    ref = oracle.orconv_forward(x, m.weight.detach().cpu().numpy(), m.indices.cpu().numpy(),
                                m.bias.detach().cpu().numpy(), pad=1)
    np.testing.assert_allclose(y, ref, rtol=RTOL, atol=ATOL)


def test_p3_full_size_vs_torch_and_reference_cuda():
    """BASELINE config 2: 1x256x128x128, 256->256.  Too big for the scalar oracle, so compare with
    (a) torchvision.ops.deform_conv2d / F.conv2d in fp32 on the GPU fed by the reference-order
    offsets, and (b) the reference's own deform_conv_cuda extension when it was prebuilt."""
    import torchvision
    from s2anet_b200.alignconv import AlignConv
    from s2anet_b200.orn import ORConv2d
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 256, 128, 128, generator=g).to(DEV)
    anc = t(synth.refined_anchors(1, 128, 128, 8, seed=1))
    ac = AlignConv(256, 256).to(DEV)
    ac.init_weights()
    oc = ORConv2d(256, 32, 3, padding=1, arf_config=(1, 8)).to(DEV)
    torch.nn.init.normal_(oc.weight, 0, 0.01)
    with torch.no_grad():
        y = ac(x, anc, 8)
        off = ac.get_offset(anc[0].reshape(-1, 5), (128, 128), 8)[None]
        ref = torch.relu(torchvision.ops.deform_conv2d(x, off, ac.deform_conv.weight, padding=1))
        err = float((y - ref).abs().max())
        assert err <= ATOL + RTOL * float(ref.abs().max()), err
        z = oc(y)
        zref = torch.nn.functional.conv2d(ref, oc.rotate_arf(), oc.bias, padding=1)
        assert float((z - zref).abs().max()) <= ATOL + RTOL * float(zref.abs().max())
    from oracle import build_oracle
    dc = build_oracle.load_ref_extension("deform_conv_cuda", "gpu")
    if dc is not None:
        out = torch.empty_like(y)
        with torch.no_grad():
            dc.deform_conv_forward_cuda(x, ac.deform_conv.weight, off.contiguous(), out, x.new_empty(0), x.new_empty(0),
                                        3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1)
        e2 = float((y - torch.relu(out)).abs().max())
        assert e2 <= ATOL + RTOL * float(out.abs().max()), e2


def test_orconv2d_and_pooling_keep_the_autograd_graph(oracle):
    """with grad enabled ORConv2d must take the differentiable route (ARF autograd Function +
    F.conv2d, reference ORConv.py:77-82) -- the fused kernel builds no graph.  Gradients are compared with a pure
    torch formulation of the same layer (index gather of the rotated bank + F.conv2d + view/max)."""
    from s2anet_b200.orn import ORConv2d, RotationInvariantPooling
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(4)
    m = ORConv2d(16, 4, 3, padding=1, arf_config=(1, 8)).to(DEV)
    m.fuse_pool = True
    pool = RotationInvariantPooling(32, 8)
    x = torch.randn(2, 16, 9, 11, generator=g).to(DEV).requires_grad_(True)
    y = m(x)
    assert y.requires_grad and getattr(y, "_s2a_pooled", None) is None
    z = pool(y)
    assert z.requires_grad
    (y.square().sum() + z.sum()).backward()
    gw, gb, gx = m.weight.grad.clone(), m.bias.grad.clone(), x.grad.clone()
    # pure-torch twin
    w2 = m.weight.detach().clone().requires_grad_(True)
    b2 = m.bias.detach().clone().requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    idx = m.indices.long().reshape(9, 8) - 1                           # [entry l, rotation k] -> source tap (nOri = 1)
    # ARF forward (ActiveRotatingFilter_cuda.cu:31-45): out[o*8 + k, i, idx[l, k]] = w[o, i, l]
    wf = w2.reshape(4, 16, 9)
    bank = torch.zeros(4, 8, 16, 9, device=DEV)
    bank = bank.scatter(3, idx.t()[None, :, None, :].expand(4, 8, 16, 9), wf[:, None].expand(4, 8, 16, 9))
    bank = bank.reshape(32, 16, 3, 3)
    assert torch.equal(bank.detach(), m.rotate_arf().detach())
    y2 = torch.nn.functional.conv2d(x2, bank, b2, padding=1)
    z2 = y2.view(2, 4, 8, 9, 11).max(dim=2)[0]
    (y2.square().sum() + z2.sum()).backward()
    for a, b in ((gw, w2.grad), (gb, b2.grad), (gx, x2.grad)):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-6
    # inference keeps the fused kernel (and its pooled by-product)
    with torch.no_grad():
        yi = m(x.detach())
    assert getattr(yi, "_s2a_pooled", None) is not None
    np.testing.assert_allclose(yi.cpu().numpy(), y.detach().cpu().numpy(), rtol=RTOL, atol=ATOL)


# ---- fp32 tensors on the tensor cores: 3 x TF32 split (round 2) ---------------------------------------------------------

@pytest.mark.parametrize("B,C,H,W,Co,stride", [(1, 32, 8, 16, 32, 8), (2, 64, 13, 21, 96, 16), (1, 256, 20, 20, 256, 32),
                                               (3, 96, 5, 3, 64, 64)])
def test_tf32x3_alignconv_matches_the_simt_fp32_kernel_and_the_oracle(oracle, B, C, H, W, Co, stride):
    """conv_tf32x3_kernel (tcgen05.mma.kind::tf32, three MMAs per K step on hi / lo splits) against the exact SIMT
    kernel of conv_f32.cu and, at the smallest shape, the scalar oracle: the fp32 bar 1e-4 + 1e-4 |ref|; measured
    differences are ~1e-6 relative (21 mantissa bits per product, fp32 accumulation)."""
    from s2anet_b200 import alignconv
    rng = np.random.default_rng(C + H + Co)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    anc = synth.refined_anchors(B, H, W, stride, seed=H)
    w = (rng.normal(size=(Co, C, 3, 3)) * 0.1).astype(np.float32)
    y = alignconv.alignconv_forward(t(x), t(anc), t(w), stride)
    assert y.is_contiguous() and y.dtype == torch.float32
    alignconv._FORCE_SIMT_F32 = True
    try:
        ref = alignconv.alignconv_forward(t(x), t(anc), t(w), stride)
    finally:
        alignconv._FORCE_SIMT_F32 = False
    err = float((y - ref).abs().max())
    assert err <= 1e-5 + 2e-5 * float(ref.abs().max()), err
    if C <= 32:
        np.testing.assert_allclose(y.cpu().numpy(), oracle.alignconv_forward(x, anc, w, stride), rtol=RTOL, atol=ATOL)


def test_tf32x3_orconv_pool_and_generic_offsets(oracle):
    """The regular-grid mode (ORConv2d: ARF folded into the packed hi / lo planes, bias, fused orientation max) and
    the explicit-offset mode (`deform_conv_forward_cuda` with fp32 tensors) of the same kernel."""
    import torchvision
    from s2anet_b200 import alignconv, dcn
    from s2anet_b200.orn import ORConv2d, orconv_forward
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(2)
    m = ORConv2d(64, 8, 3, padding=1, arf_config=(1, 8)).to(DEV)
    with torch.no_grad():
        m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.05)
        m.bias.copy_(torch.randn(64, generator=g) * 0.1)
    x = torch.randn(2, 64, 19, 27, generator=g).to(DEV)
    y, yp = orconv_forward(x, m.weight, m.indices, m.bias, with_pool=True)
    ref = torch.nn.functional.conv2d(x, m.rotate_arf(), m.bias, padding=1)
    assert float((y - ref).abs().max()) <= 1e-5 + 2e-5 * float(ref.abs().max())
    assert torch.equal(yp, y.view(2, 8, 8, 19, 27).max(dim=2)[0])
    w = (torch.randn(96, 64, 3, 3, generator=g) * 0.05).to(DEV)
    off = (torch.randn(2, 18, 19, 27, generator=g) * 2.0).to(DEV)
    out = torch.empty(2, 96, 19, 27, device=DEV)
    e = x.new_empty(0)
    assert dcn.deform_conv_forward_cuda(x, w, off, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2) == 1
    ref = torchvision.ops.deform_conv2d(x, off, w, padding=1)
    assert float((out - ref).abs().max()) <= ATOL + RTOL * float(ref.abs().max())
    alignconv._FORCE_SIMT_F32 = True
    try:
        out2 = torch.empty_like(out)
        dcn.deform_conv_forward_cuda(x, w, off, out2, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2)
    finally:
        alignconv._FORCE_SIMT_F32 = False
    assert float((out - out2).abs().max()) <= 1e-5 + 2e-5 * float(out2.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 256, 16, 16), (1, 24, 5, 3), (3, 33, 7, 9), (1, 64, 128, 128)])
def test_native_layout_conversion_is_a_bit_exact_copy(dtype, shape):
    """s2a_transpose_planes (NCHW <-> NHWC around the tensor-core kernels) against torch's own conversion."""
    from s2anet_b200 import conv_tc
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to(dtype).to(DEV)
    y = conv_tc._nhwc(x)
    assert y.is_contiguous(memory_format=torch.channels_last) and y.shape == x.shape
    assert torch.equal(y, x)
    assert y.data_ptr() != x.data_ptr() or shape[1] == 1
    back = torch.empty_like(x)
    conv_tc.nchw_from_nhwc(y, back)
    assert back.is_contiguous() and torch.equal(back, x)
    assert conv_tc._nhwc(y) is y                         # already NHWC: no copy


def test_tf32x3_edge_cases_tiny_maps_and_samples_outside(oracle):
    """1 x 1 and 2 x 3 maps (a single partial tile), offsets that throw every sample out of the map (all-zero output,
    deform_conv_cuda_kernel.cu:228) and an empty batch, through the fp32 tensor-core route of deform_conv_forward_cuda."""
    import torchvision
    from s2anet_b200 import dcn
    g = torch.Generator().manual_seed(11)
    e = torch.empty(0, device=DEV)
    for (B, H, W) in ((1, 1, 1), (2, 2, 3)):
        x = torch.randn(B, 32, H, W, generator=g).to(DEV)
        w = (torch.randn(32, 32, 3, 3, generator=g) * 0.1).to(DEV)
        off = (torch.randn(B, 18, H, W, generator=g) * 0.7).to(DEV)
        out = torch.full((B, 32, H, W), 7.0, device=DEV)
        assert dcn.deform_conv_forward_cuda(x, w, off, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B) == 1
        ref = torchvision.ops.deform_conv2d(x, off, w, padding=1)
        assert float((out - ref).abs().max()) <= ATOL + RTOL * float(ref.abs().max())
        far = torch.full_like(off, 1000.0)
        dcn.deform_conv_forward_cuda(x, w, far, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
        assert float(out.abs().max()) == 0.0
    x0 = torch.empty(0, 32, 4, 4, device=DEV)
    out0 = torch.empty(0, 32, 4, 4, device=DEV)
    assert dcn.deform_conv_forward_cuda(x0, w, torch.empty(0, 18, 4, 4, device=DEV), out0, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1) == 1
