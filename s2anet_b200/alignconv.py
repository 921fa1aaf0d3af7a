"""AlignConv with the reference's module surface (reference: models/alignconv.py:8-98).

`forward(x, anchors, stride)` is ONE fused kernel: the anchor-derived sampling positions, the
bilinear gather, the 256x2304 contraction and the ReLU (the reference runs ~25 small ATen kernels
per image to build an 18-channel offset tensor, then im2col + cuBLAS + a transposing copy).
`get_offset` is kept for API parity and for tests; the fused forward never calls it.
"""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import _lib
from .dcn import DeformConv


_FORCE_SIMT_F32 = False      # set through set_fp32_path(); read by alignconv / orn / dcn when they route fp32 tensors


def set_fp32_path(path):
    """Which kernel fp32 tensors take in AlignConv / DeformConv / ORConv2d forward (C % 32 == 0, C_out % 32 == 0, <= 256):
    "tf32x3" (default) -- conv_tf32x3_kernel on tcgen05, three TF32 MMAs per K step on hi / lo splits: rel-L2 ~5e-6 of an
    fp64 evaluation (profiles/r2_parity_errors.json), 0.11 ms at P3 batch 1 where the reference takes 0.59 ms;
    "simt" -- the FMA kernel of conv_f32.cu: individually rounded fp32 operations in the reference's order, rel-L2 ~1e-6
    like the reference's own SGEMM path, 0.98 ms.  Other shapes always take the SIMT kernel."""
    global _FORCE_SIMT_F32
    if path not in ("tf32x3", "simt"):
        raise ValueError("set_fp32_path: expected 'tf32x3' or 'simt', got %r" % (path,))
    _FORCE_SIMT_F32 = path == "simt"


def alignconv_forward(x, anchors, weight, stride):
    """relu(deform_conv(x, offsets(anchors), weight)); x [B,C,H,W], anchors [B,H,W,5], weight [Co,C,3,3]."""
    dev = _lib.require_cuda(x, anchors, weight)
    B, C, H, W = x.shape
    if anchors.numel() != B * H * W * 5:
        raise ValueError("anchors must be [B,H,W,5] matching x")
    if tuple(weight.shape[1:]) != (C, 3, 3):
        raise ValueError("AlignConv weight must be [C_out, C_in, 3, 3]")
    Co = weight.size(0)
    if x.dtype != torch.float32:
        from . import conv_tc
        return conv_tc.alignconv_forward_tc(x, anchors, weight, stride)
    from . import conv_tc
    if conv_tc.tf32x3_supported(C, Co) and not _FORCE_SIMT_F32:
        # fp32 on the tensor cores (3 x TF32 split): conv_tf32x3_kernel
        return conv_tc.conv_forward_tf32x3(x, anchors.reshape(B, H, W, 5), 0, conv_tc.pack_weight_tf32(weight), relu=True,
                                           stride=stride)
    xc = x.contiguous()
    a = anchors.to(torch.float32).contiguous()
    w = weight.to(torch.float32).contiguous()
    out = torch.empty((B, Co, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_alignconv_forward_f32(_lib.ptr(xc), _lib.ptr(a), _lib.ptr(w), _lib.ptr(out), B, C, H, W, Co,
                                                   float(stride), _lib.stream_ptr(dev))
    _lib.check(rc, "alignconv_forward")
    return out


class AlignConv(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size=3, deformable_groups=1):
        super(AlignConv, self).__init__()
        self.kernel_size = _pair(kernel_size)
        self.padding = tuple((size - 1) // 2 for size in self.kernel_size)
        # parameter path `align_conv.deform_conv.weight` is part of the checkpoint contract
        self.deform_conv = DeformConv(in_channels, out_channels, kernel_size=self.kernel_size, padding=self.padding,
                                      deformable_groups=deformable_groups)
        self.relu = nn.ReLU(inplace=True)

    def init_weights(self):
        nn.init.normal_(self.deform_conv.weight, 0, 0.01)       # normal_init(std=0.01), alignconv.py:25-26

    @torch.no_grad()
    def get_offset(self, anchors, featmap_size, stride):
        """reference: models/alignconv.py:29-86 -- [H*W,5] anchors -> [18,H,W] offsets (dy, dx per tap)."""
        dtype, device = anchors.dtype, anchors.device
        feat_h, feat_w = featmap_size
        pady = (self.kernel_size[0] - 1) // 2
        padx = (self.kernel_size[1] - 1) // 2
        idy = torch.arange(-pady, pady + 1, dtype=dtype, device=device)
        idx = torch.arange(-padx, padx + 1, dtype=dtype, device=device)
        yy, xx = torch.meshgrid(idy, idx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        xc = torch.arange(0, feat_w, device=device, dtype=dtype)
        yc = torch.arange(0, feat_h, device=device, dtype=dtype)
        yc, xc = torch.meshgrid(yc, xc, indexing="ij")
        xc, yc = xc.reshape(-1), yc.reshape(-1)
        x_conv, y_conv = xc[:, None] + xx, yc[:, None] + yy
        x_ctr, y_ctr, w, h, a = torch.unbind(anchors, dim=1)
        x_ctr, y_ctr, w, h = x_ctr / stride, y_ctr / stride, w / stride, h / stride
        cos, sin = torch.cos(a), torch.sin(a)
        dw, dh = w / self.kernel_size[1], h / self.kernel_size[0]
        x, y = dw[:, None] * xx, dh[:, None] * yy
        xr = cos[:, None] * x - sin[:, None] * y
        yr = sin[:, None] * x + cos[:, None] * y
        x_anchor, y_anchor = xr + x_ctr[:, None], yr + y_ctr[:, None]
        offset = torch.stack([y_anchor - y_conv, x_anchor - x_conv], dim=-1)
        return offset.reshape(anchors.size(0), -1).permute(1, 0).reshape(-1, feat_h, feat_w)

    def forward(self, x, anchors, stride):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or self.deform_conv.weight.requires_grad)
        if self.kernel_size != (3, 3) or self.deform_conv.deformable_groups != 1 or needs_grad:
            # generic route of the reference: explicit offsets + DeformConv (the autograd Function: training)
            num_imgs, H, W = anchors.shape[:3]
            offset = torch.stack([self.get_offset(anchors[i].reshape(-1, 5), (H, W), stride) for i in range(num_imgs)])
            return self.relu(self.deform_conv(x, offset))
        return alignconv_forward(x, anchors, self.deform_conv.weight, stride)
