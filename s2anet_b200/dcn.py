"""Deformable convolution: `DeformConvFunction`, `deform_conv`, `DeformConv` with the reference's
names, argument order and error behaviour (reference: models/dcn/deform_conv.py:13-109, 205-272;
extension entry deform_conv_forward_cuda, models/dcn/src/deform_conv_cuda.cpp:152-260)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _lib


def deform_conv_forward_cuda(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW,
                             dilationH, group, deformable_group, im2col_step):
    """Extension-level entry with the reference's positional signature (width-first kernel /
    stride / pad / dilation order, deform_conv_cuda.cpp:152-157).  Writes into the caller's
    `output`; `columns` / `ones` are the reference's scratch handles and are ignored (no im2col
    buffer exists); `im2col_step` only has to divide the batch.  Returns 1."""
    dev = _lib.require_cuda(input, weight, offset, output)
    if weight.dim() != 4:
        raise RuntimeError("4D weight tensor (nOutputPlane,nInputPlane,kH,kW) expected, but got: %d" % weight.dim())
    if weight.size(2) != kH or weight.size(3) != kW:
        raise RuntimeError("kernel size should be consistent with weight")
    squeeze = input.dim() == 3
    if squeeze:
        input, offset = input.unsqueeze(0), offset.unsqueeze(0)
    if input.dim() != 4:
        raise RuntimeError("3D or 4D input tensor expected but got: %d" % input.dim())
    B, C, H, W = input.shape
    if offset.size(0) != B:
        raise RuntimeError("invalid batch size of offset")
    if im2col_step <= 0 or B % im2col_step != 0:
        raise RuntimeError("im2col step must divide batchsize")
    Co = weight.size(0)
    if weight.size(1) * group != C:
        raise RuntimeError("invalid number of input planes, expected: %d, but got: %d" % (weight.size(1) * group, C))
    Ho = (H + 2 * padH - (dilationH * (kH - 1) + 1)) // dH + 1
    Wo = (W + 2 * padW - (dilationW * (kW - 1) + 1)) // dW + 1
    if offset.size(1) != deformable_group * 2 * kH * kW:
        raise RuntimeError("invalid number of channels of offset")
    if offset.size(2) != Ho or offset.size(3) != Wo:
        raise RuntimeError("invalid spatial size of offset, expected height: %d width: %d, but got height: %d "
                           "width: %d" % (Ho, Wo, offset.size(2), offset.size(3)))
    if input.dtype in (torch.float16, torch.bfloat16):
        # The reference dispatches scalar_t = half here (deform_conv_cuda_kernel.cu:258) and val.py runs the model in
        # fp16 by default (val.py:126,196,246).  S2ANet's configuration goes to the tcgen05 kernel (offsets ->
        # gather recipes; positions and bilinear weights rounded like the reference's half arithmetic); every other
        # geometry runs the exact fp32 kernel on up-cast operands and rounds the result once.
        from . import conv_tc
        if conv_tc.deform_conv_tc_supported(C, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group):
            # (the packer rounds an fp32 weight to the 16-bit type once, which is what `weight.type_as(input)` does)
            y = conv_tc.deform_conv_forward_tc(input, offset, weight, out=output)
            if y is not output:
                conv_tc.nchw_from_nhwc(y, output)
            return 1
        out32 = torch.empty((B, Co, Ho, Wo), dtype=torch.float32, device=dev)
        deform_conv_forward_cuda(input.float(), weight.float(), offset.float(), out32, columns, ones, kW, kH, dW, dH, padW,
                                 padH, dilationW, dilationH, group, deformable_group, im2col_step)
        output.copy_(out32.view_as(output))
        return 1
    if input.dtype != torch.float32:
        raise TypeError("deform_conv_forward_cuda: unsupported dtype %s (float32, float16, bfloat16)" % input.dtype)
    from . import alignconv, conv_tc
    s2a_geometry = (kH == 3 and kW == 3 and dH == 1 and dW == 1 and padH == 1 and padW == 1 and dilationH == 1 and
                    dilationW == 1 and group == 1 and deformable_group == 1)
    if s2a_geometry and conv_tc.tf32x3_supported(C, Co) and not alignconv._FORCE_SIMT_F32:
        # fp32 on the tensor cores (3 x TF32 split): explicit offsets -> sampling positions inside the kernel
        y = conv_tc.conv_forward_tf32x3(input, offset, 1, conv_tc.pack_weight_tf32(weight), out=output)
        if y is not output:
            output.copy_(y.view_as(output))
        return 1
    x = input.contiguous()
    off = offset.to(torch.float32).contiguous()
    w = weight.to(torch.float32).contiguous()
    out = output if (output.is_contiguous() and output.dtype == torch.float32) else torch.empty(
        (B, Co, Ho, Wo), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_deform_conv_forward_f32(_lib.ptr(x), _lib.ptr(off), _lib.ptr(w), _lib.ptr(out), B, C, H, W,
                                                     Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
                                                     deformable_group, 0, _lib.stream_ptr(dev))
    _lib.check(rc, "deform_conv_forward_cuda")
    if out is not output:
        output.copy_(out.view_as(output))
    return 1


def _bwd_common(input, offset, weight, kW, kH, group, deformable_group, im2col_step, what):
    dev = _lib.require_cuda(input, offset, weight)
    if input.dim() != 4:
        raise RuntimeError("%s: 4D input tensor expected but got: %d" % (what, input.dim()))
    if weight.size(2) != kH or weight.size(3) != kW:
        raise RuntimeError("kernel size should be consistent with weight")
    if im2col_step <= 0 or input.size(0) % im2col_step != 0:
        raise RuntimeError("im2col step must divide batchsize")
    if weight.size(1) * group != input.size(1):
        raise RuntimeError("invalid number of input planes, expected: %d, but got: %d" % (weight.size(1) * group, input.size(1)))
    if offset.size(1) != deformable_group * 2 * kH * kW or offset.size(0) != input.size(0):
        raise RuntimeError("invalid shape of offset")
    return dev


def deform_conv_backward_input_cuda(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH,
                                    padW, padH, dilationW, dilationH, group, deformable_group, im2col_step):
    """reference: deform_conv_cuda.cpp:262-373.  ACCUMULATES the input and offset gradients into the caller's
    pre-zeroed `gradInput` / `gradOffset` (deform_conv.py:88-89); `columns` is the reference's scratch handle and is
    ignored.  The column gradient W^T x gradOutput is a library GEMM (as `addmm_` is in the reference, :331-337);
    the bilinear scatter and the coordinate gradient run in one pass of s2a_deform_col2im_f32.  Returns 1."""
    dev = _bwd_common(input, offset, weight, kW, kH, group, deformable_group, im2col_step, "deform_conv_backward_input_cuda")
    _lib.require_cuda(gradOutput, gradInput, gradOffset)
    B, C, H, W = input.shape
    Co = weight.size(0)
    if gradOutput.dtype in (torch.float16, torch.bfloat16) and gradOutput.dtype == input.dtype:
        from . import conv_tc
        if conv_tc.deform_conv_dgrad_tc_supported(C, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group):
            # 16-bit training (the reference's half route): tcgen05 implicit GEMMs with a scatter epilogue, no columns
            gi, goff = conv_tc.deform_conv_dgrad_tc(gradOutput.detach(), offset.detach(), weight, x=input.detach(),
                                                    need_offset_grad=True)
            gradInput.add_(gi.to(gradInput.dtype))
            gradOffset.add_(goff.to(gradOffset.dtype))
            return 1
    x = input.detach().float().contiguous()
    off = offset.detach().float().contiguous()
    go = gradOutput.detach().float().contiguous()
    Ho, Wo = go.shape[2:]
    kk = kH * kW
    wg = weight.detach().float().reshape(group, Co // group, (C // group) * kk)
    gi = gradInput if (gradInput.dtype == torch.float32 and gradInput.is_contiguous()) else gradInput.float().contiguous()
    gf = gradOffset if (gradOffset.dtype == torch.float32 and gradOffset.is_contiguous()) else gradOffset.float().contiguous()
    lib = _lib.load()
    for b0 in range(0, B, im2col_step):                   # the reference's im2col_step chunking bounds the scratch
        b1 = b0 + im2col_step
        gcol = torch.matmul(wg.transpose(1, 2)[None], go[b0:b1].reshape(b1 - b0, group, Co // group, Ho * Wo))
        gcol = gcol.reshape(b1 - b0, C * kk, Ho * Wo).contiguous()
        with torch.cuda.device(dev):
            rc = lib.s2a_deform_col2im_f32(_lib.ptr(gcol), _lib.ptr(x[b0:b1]), _lib.ptr(off[b0:b1]), _lib.ptr(gi[b0:b1]),
                                           _lib.ptr(gf[b0:b1]), b1 - b0, C, H, W, kH, kW, dH, dW, padH, padW, dilationH,
                                           dilationW, deformable_group, _lib.stream_ptr(dev))
        _lib.check(rc, "deform_col2im")
    if gi is not gradInput:
        gradInput.copy_(gi)
    if gf is not gradOffset:
        gradOffset.copy_(gf)
    return 1


def deform_conv_backward_parameters_cuda(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH,
                                         dilationW, dilationH, group, deformable_group, scale, im2col_step):
    """reference: deform_conv_cuda.cpp:376-489.  gradWeight += scale * gradOutput x columns^T with the columns of
    s2a_deform_im2col_f32 and a library GEMM (the reference's `addmm_`, :459-467).  Returns 1."""
    dev = _bwd_common(input, offset, weight=gradWeight, kW=kW, kH=kH, group=group, deformable_group=deformable_group,
                      im2col_step=im2col_step, what="deform_conv_backward_parameters_cuda")
    _lib.require_cuda(gradOutput)
    B, C, H, W = input.shape
    Co = gradWeight.size(0)
    if input.dtype in (torch.float16, torch.bfloat16) and gradOutput.dtype == input.dtype:
        from . import conv_tc
        if conv_tc.deform_conv_wgrad_tc_supported(C, Co, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group, deformable_group):
            # 16-bit training: tcgen05 wgrad (MN-major operands, accumulator in tensor memory), no columns
            dw = conv_tc.deform_conv_wgrad_tc(input.detach(), offset.detach(), gradOutput.detach())
            gradWeight.add_(dw.to(gradWeight.dtype), alpha=float(scale))
            return 1
    x = input.detach().float().contiguous()
    off = offset.detach().float().contiguous()
    go = gradOutput.detach().float().contiguous()
    Ho, Wo = go.shape[2:]
    kk = kH * kW
    acc = torch.zeros((group, Co // group, (C // group) * kk), dtype=torch.float32, device=dev)
    lib = _lib.load()
    for b0 in range(0, B, im2col_step):
        b1 = b0 + im2col_step
        col = torch.empty((b1 - b0, C * kk, Ho * Wo), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.s2a_deform_im2col_f32(_lib.ptr(x[b0:b1]), _lib.ptr(off[b0:b1]), _lib.ptr(col), b1 - b0, C, H, W, kH, kW,
                                           dH, dW, padH, padW, dilationH, dilationW, deformable_group, _lib.stream_ptr(dev))
        _lib.check(rc, "deform_im2col")
        g4 = go[b0:b1].reshape(b1 - b0, group, Co // group, Ho * Wo)
        c4 = col.reshape(b1 - b0, group, (C // group) * kk, Ho * Wo)
        acc += torch.matmul(g4, c4.transpose(2, 3)).sum(0)
    gradWeight.add_(acc.reshape(gradWeight.shape).to(gradWeight.dtype), alpha=float(scale))
    return 1


class DeformConvFunction(Function):
    """reference: models/dcn/deform_conv.py:13-109."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        ctx.stride = _pair(stride)
        ctx.padding = _pair(padding)
        ctx.dilation = _pair(dilation)
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.im2col_step = im2col_step
        offset = offset.type_as(input)          # :45-46, the dtype of `input` decides
        weight = weight.type_as(input)
        ctx.save_for_backward(input, offset, weight)
        output = input.new_empty(DeformConvFunction._output_size(input, weight, ctx.padding, ctx.dilation, ctx.stride))
        ctx.bufs_ = [input.new_empty(0), input.new_empty(0)]
        if not input.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(ctx.im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        deform_conv_forward_cuda(input, weight, offset, output, ctx.bufs_[0], ctx.bufs_[1], weight.size(3),
                                 weight.size(2), ctx.stride[1], ctx.stride[0], ctx.padding[1], ctx.padding[0],
                                 ctx.dilation[1], ctx.dilation[0], ctx.groups, ctx.deformable_groups, cur_im2col_step)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        """reference: models/dcn/deform_conv.py:72-109."""
        input, offset, weight = ctx.saved_tensors
        grad_input = grad_offset = grad_weight = None
        if not grad_output.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(ctx.im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        from . import conv_tc
        tc_dgrad = (grad_output.dtype in (torch.float16, torch.bfloat16) and grad_output.dtype == input.dtype and
                    conv_tc.deform_conv_dgrad_tc_supported(input.size(1), weight.size(0), weight.size(2), weight.size(3),
                                                           ctx.stride[0], ctx.stride[1], ctx.padding[0], ctx.padding[1],
                                                           ctx.dilation[0], ctx.dilation[1], ctx.groups, ctx.deformable_groups))
        if tc_dgrad and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            # 16-bit training: tcgen05 dgrad with the scatter epilogue; the offset gradient only when somebody wants it
            # (S2ANet detaches the anchors, models/head.py:333-335, so normally nobody does)
            gi, goff = conv_tc.deform_conv_dgrad_tc(grad_output, offset, weight, x=input,
                                                    need_offset_grad=bool(ctx.needs_input_grad[1]))
            grad_input = gi.to(input.dtype)
            grad_offset = goff.to(offset.dtype) if goff is not None else None
        elif ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_input = torch.zeros_like(input)
            grad_offset = torch.zeros_like(offset)
            deform_conv_backward_input_cuda(input, offset, grad_output, grad_input, grad_offset, weight, ctx.bufs_[0],
                                            weight.size(3), weight.size(2), ctx.stride[1], ctx.stride[0], ctx.padding[1],
                                            ctx.padding[0], ctx.dilation[1], ctx.dilation[0], ctx.groups,
                                            ctx.deformable_groups, cur_im2col_step)
        if ctx.needs_input_grad[2]:
            grad_weight = torch.zeros_like(weight)
            deform_conv_backward_parameters_cuda(input, offset, grad_output, grad_weight, ctx.bufs_[0], ctx.bufs_[1],
                                                 weight.size(3), weight.size(2), ctx.stride[1], ctx.stride[0],
                                                 ctx.padding[1], ctx.padding[0], ctx.dilation[1], ctx.dilation[0],
                                                 ctx.groups, ctx.deformable_groups, 1, cur_im2col_step)
        return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)

    @staticmethod
    def _output_size(input, weight, padding, dilation, stride):
        channels = weight.size(0)
        output_size = (input.size(0), channels)
        for d in range(input.dim() - 2):
            in_size = input.size(d + 2)
            pad = padding[d]
            kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
            stride_ = stride[d]
            output_size += ((in_size + (2 * pad) - kernel) // stride_ + 1,)
        if not all(map(lambda s: s > 0, output_size)):
            raise ValueError("convolution input is too small (output would be {})".format(
                'x'.join(map(str, output_size))))
        return output_size


deform_conv = DeformConvFunction.apply


class DeformConv(nn.Module):
    """reference: models/dcn/deform_conv.py:205-272 (weight [C_out, C_in/groups, kH, kW], no bias)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super(DeformConv, self).__init__()
        assert not bias
        assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(
            out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        input_pad = (x.size(2) < self.kernel_size[0] or x.size(3) < self.kernel_size[1])
        if input_pad:
            pad_h = max(self.kernel_size[0] - x.size(2), 0)
            pad_w = max(self.kernel_size[1] - x.size(3), 0)
            x = F.pad(x, (0, pad_w, 0, pad_h), 'constant', 0).contiguous()
            offset = F.pad(offset, (0, pad_w, 0, pad_h), 'constant', 0).contiguous()
        out = deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                          self.deformable_groups)
        if input_pad:
            out = out[:, :, :out.size(2) - pad_h, :out.size(3) - pad_w].contiguous()
        return out
