"""ORN surface: `ORConv2d`, `RotationInvariantPooling`, `active_rotating_filter`,
`ActiveRotatingFilter`, plus the extension-level `arf_forward` / `arf_backward`
(reference: models/orn/modules/ORConv.py:12-101, models/orn/functions/active_rotating_filter.py:12-43,
models/orn/functions/rotation_invariant_pooling.py:7-27, models/orn/src/vision.cpp:7-12)."""
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules import Conv2d
from torch.nn.modules.utils import _pair
from torch.nn.parameter import Parameter

from . import _lib, _torch_ext


def arf_forward(weight, indices):
    """orn_cuda.arf_forward: weight [O,I,nOri,kH,kW], indices uint8 [nOri,kH,kW,nRot] ->
    [O*nRot, I*nOri, kH, kW] (new tensor)."""
    dev = _lib.require_cuda(weight, indices)
    if weight.dim() != 5:
        raise RuntimeError("only supports a batch of ARFs.")
    ext = _torch_ext.module()
    if ext is not None and indices.dtype == torch.uint8:
        out = ext.arf_forward(weight, indices)
        _lib.check(0, "arf_forward")
        return out
    O, I, nOri, kH, kW = weight.shape
    nRot = indices.size(3)
    w = weight.contiguous()
    idx = indices.to(torch.uint8).contiguous()
    out = torch.empty((O * nRot, I * nOri, kH, kW), dtype=weight.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_arf_forward(_lib.ptr(w), _lib.ptr(idx), _lib.ptr(out), O, I, nOri, kH, kW, nRot,
                                         _lib.dtype_code(w), _lib.stream_ptr(dev))
    _lib.check(rc, "arf_forward")
    return out


def arf_backward(indices, grad_output):
    """orn_cuda.arf_backward: indices uint8 [nOri,kH,kW,nRot], grad [O*nRot, I*nOri, kH, kW] ->
    [O, I, nOri, kH, kW]."""
    dev = _lib.require_cuda(indices, grad_output)
    nOri, kH, kW, nRot = indices.shape
    O, I = grad_output.size(0) // nRot, grad_output.size(1) // nOri
    g = grad_output.contiguous()
    idx = indices.to(torch.uint8).contiguous()
    out = torch.empty((O, I, nOri, kH, kW), dtype=g.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_arf_backward(_lib.ptr(g), _lib.ptr(idx), _lib.ptr(out), O, I, nOri, kH, kW, nRot,
                                          _lib.dtype_code(g), _lib.stream_ptr(dev))
    _lib.check(rc, "arf_backward")
    return out


class _ActiveRotatingFilter(Function):
    @staticmethod
    def forward(ctx, input, indices):
        indices = indices.byte()
        ctx.input = input
        output = arf_forward(input, indices)
        ctx.save_for_backward(indices)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        indices, = ctx.saved_tensors
        return arf_backward(indices, grad_output), None


active_rotating_filter = _ActiveRotatingFilter.apply


class ActiveRotatingFilter(nn.Module):
    def __init__(self, indices):
        super(ActiveRotatingFilter, self).__init__()
        self.indices = indices

    def forward(self, input):
        return active_rotating_filter(input, self.indices)


def ri_pool(x, n_orientation=8):
    dev = _lib.require_cuda(x)
    N, c, h, w = x.shape
    xc = x.contiguous()
    out = torch.empty((N, c // n_orientation, h, w), dtype=x.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_ri_pool_forward(_lib.ptr(xc), _lib.ptr(out), N, c, h * w, n_orientation,
                                             _lib.dtype_code(xc), _lib.stream_ptr(dev))
    _lib.check(rc, "ri_pool")
    return out


class RotationInvariantPooling(nn.Module):
    """reference: models/orn/functions/rotation_invariant_pooling.py:7-27."""

    def __init__(self, nInputPlane, nOrientation=8):
        super(RotationInvariantPooling, self).__init__()
        self.nInputPlane = nInputPlane
        self.nOrientation = nOrientation

    def forward(self, x):
        pooled = getattr(x, "_s2a_pooled", None)     # produced by the fused ORConv2d epilogue
        if pooled is not None and self.nOrientation == 8:
            return pooled
        if torch.is_grad_enabled() and x.requires_grad:
            # training: the reference's own differentiable formulation (rotation_invariant_pooling.py:19-27)
            N, c, h, w = x.size()
            return x.view(N, -1, self.nOrientation, h, w).max(dim=2, keepdim=False)[0]
        return ri_pool(x, self.nOrientation)


def orconv_forward(x, weight, indices, bias, with_pool=False):
    """conv2d(x, ARF(weight), bias, padding=1) for a 3x3 ORConv in ONE kernel (ARF folded into the
    weight-tile load); optionally also the 8-way orientation max-pool from the same epilogue."""
    dev = _lib.require_cuda(x, weight, indices, bias)
    O, I, nOri, kH, kW = weight.shape
    nRot = indices.size(3)
    B, C, H, W = x.shape
    if x.dtype != torch.float32:
        from . import conv_tc
        return conv_tc.orconv_forward_tc(x, weight, indices, bias, with_pool)
    from . import alignconv, conv_tc
    if conv_tc.tf32x3_supported(C, O * nRot) and C == I * nOri and not alignconv._FORCE_SIMT_F32 and \
            (not with_pool or (O * nRot) % 8 == 0):
        # fp32 on the tensor cores (3 x TF32 split); the ARF rotation is folded into the weight packing
        return conv_tc.conv_forward_tf32x3(x, None, 2, conv_tc.pack_weight_tf32(weight, indices), bias=bias, relu=False,
                                           with_pool=with_pool)
    xc = x.contiguous()
    w = weight.to(torch.float32).contiguous()
    idx = indices.to(torch.uint8).contiguous()
    b = None if bias is None else bias.to(torch.float32).contiguous()
    out = torch.empty((B, O * nRot, H, W), dtype=torch.float32, device=dev)
    pooled = torch.empty((B, O * nRot // 8, H, W), dtype=torch.float32, device=dev) if with_pool else None
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_orconv_forward_f32(_lib.ptr(xc), _lib.ptr(w), _lib.ptr(idx), _lib.ptr(b), _lib.ptr(out),
                                                _lib.ptr(pooled), B, H, W, O, I, nOri, nRot, _lib.stream_ptr(dev))
    _lib.check(rc, "orconv_forward")
    return (out, pooled) if with_pool else out


class ORConv2d(Conv2d):
    """reference: models/orn/modules/ORConv.py:12-101.  Same constructor, parameters
    (`weight [O,I,nOri,kH,kW]`, `bias [O*nRot]`), buffer (`indices` uint8 [nOri,kH,kW,nRot]) and
    attributes; `forward` is fused for the configuration S2ANet uses (3x3, stride 1, padding 1,
    dilation 1, groups 1) and otherwise follows the reference literally (ARF kernel + F.conv2d)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, arf_config=None, stride=1, padding=0, dilation=1,
                 groups=1, bias=True):
        self.nOrientation, self.nRotation = _pair(arf_config)
        assert (math.log(self.nOrientation) + 1e-5) % math.log(2) < 1e-3, 'invalid nOrientation {}'.format(
            self.nOrientation)
        assert (math.log(self.nRotation) + 1e-5) % math.log(2) < 1e-3, 'invalid nRotation {}'.format(self.nRotation)
        super(ORConv2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.register_buffer("indices", self.get_indices())
        self.weight = Parameter(torch.Tensor(out_channels, in_channels, self.nOrientation, *self.kernel_size))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels * self.nRotation))
        self.fuse_pool = False          # set by the head when RotationInvariantPooling follows directly
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels * self.nOrientation
        for k in self.kernel_size:
            n *= k
        self.weight.data.normal_(0, math.sqrt(2.0 / n))
        if self.bias is not None:
            self.bias.data.zero_()

    def get_indices(self, mode='fast'):
        kernel_indices = {
            1: {0: (1,), 45: (1,), 90: (1,), 135: (1,), 180: (1,), 225: (1,), 270: (1,), 315: (1,)},
            3: {
                0: (1, 2, 3, 4, 5, 6, 7, 8, 9),
                45: (2, 3, 6, 1, 5, 9, 4, 7, 8),
                90: (3, 6, 9, 2, 5, 8, 1, 4, 7),
                135: (6, 9, 8, 3, 5, 7, 2, 1, 4),
                180: (9, 8, 7, 6, 5, 4, 3, 2, 1),
                225: (8, 7, 4, 9, 5, 1, 6, 3, 2),
                270: (7, 4, 1, 8, 5, 2, 9, 6, 3),
                315: (4, 1, 2, 7, 5, 3, 8, 9, 6),
            },
        }
        delta_orientation = 360 / self.nOrientation
        delta_rotation = 360 / self.nRotation
        kH, kW = self.kernel_size
        indices = torch.zeros(self.nOrientation * kH * kW, self.nRotation, dtype=torch.uint8)
        for i in range(0, self.nOrientation):
            for j in range(0, kH * kW):
                for k in range(0, self.nRotation):
                    angle = delta_rotation * k
                    layer = (i + math.floor(angle / delta_orientation)) % self.nOrientation
                    kernel = kernel_indices[kW][angle][j]
                    indices[i * kH * kW + j, k] = int(layer * kH * kW + kernel)
        return indices.view(self.nOrientation, kH, kW, self.nRotation)

    def rotate_arf(self):
        return active_rotating_filter(self.weight, self.indices)

    def _fusable(self):
        return (self.kernel_size == (3, 3) and self.stride == (1, 1) and self.padding == (1, 1)
                and self.dilation == (1, 1) and self.groups == 1)

    def _needs_grad(self, input):
        return torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad or
                                            (self.bias is not None and self.bias.requires_grad))

    def forward(self, input):
        # The fused kernel is an inference path (no autograd graph): whenever a gradient can flow, take the
        # reference's route -- the ARF autograd Function (arf_forward / arf_backward kernels) + F.conv2d.
        if self._fusable() and input.is_cuda and not self._needs_grad(input):
            pool = self.fuse_pool and self.nRotation == 8
            res = orconv_forward(input, self.weight, self.indices, self.bias, with_pool=pool)
            if pool:
                out, pooled = res
                out._s2a_pooled = pooled
                return out
            return res
        return F.conv2d(input, self.rotate_arf(), self.bias, self.stride, self.padding, self.dilation, self.groups)

    def __repr__(self):
        arf_config = '[{}]'.format(self.nOrientation) if self.nOrientation == self.nRotation \
            else '[{}-{}]'.format(self.nOrientation, self.nRotation)
        s = ('{name}({arf_config} {in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}')
        if self.padding != (0,) * len(self.padding):
            s += ', padding={padding}'
        if self.dilation != (1,) * len(self.dilation):
            s += ', dilation={dilation}'
        if self.groups != 1:
            s += ', groups={groups}'
        if self.bias is None:
            s += ', bias=False'
        s += ')'
        return s.format(name=self.__class__.__name__, arf_config=arf_config, **self.__dict__)
