"""Host side of the tcgen05 implicit-GEMM path (csrc/conv_tc.cu): channels-last bookkeeping and the
packed-weight cache.  Activations are bf16 / fp16 tensors in torch.channels_last memory format
(physically NHWC), which is what cuDNN's tensor-core convolutions around these ops produce, so
no layout copies happen inside a channels-last model."""
import torch

from . import _lib

_PACK_CACHE = {}      # (data_ptr, version, dtype, device, kind) -> packed tensor


def _nhwc(x):
    """A [B,C,H,W] tensor whose memory is NHWC-contiguous (no copy if it already is)."""
    if x.dim() != 4:
        raise ValueError("expected a 4-D [B,C,H,W] tensor")
    return x.contiguous(memory_format=torch.channels_last)


def pack_weight(weight, dtype, indices=None):
    """Pack (and cache) conv weights for the tensor-core kernel; ORConv banks go through the ARF map."""
    key = (weight.data_ptr(), weight._version, dtype, weight.device, None if indices is None else indices.data_ptr())
    hit = _PACK_CACHE.get(key)
    if hit is not None:
        return hit
    dev = weight.device
    w = weight.detach().contiguous()
    if indices is None:
        Co, C = w.size(0), w.size(1)
        nOri = nRot = 1
        idx = None
    else:
        O, I, nOri = w.size(0), w.size(1), w.size(2)
        nRot = indices.size(3)
        Co, C = O * nRot, I * nOri
        idx = indices.to(torch.uint8).contiguous()
    packed = torch.empty((Co, 9 * C), dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_conv_pack_weight(_lib.ptr(w), _lib.dtype_code(w), _lib.ptr(idx), _lib.ptr(packed),
                                              _lib.dtype_code(packed), Co, C, nOri, nRot, _lib.stream_ptr(dev))
    _lib.check(rc, "conv_pack_weight")
    if len(_PACK_CACHE) > 64:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = packed
    return packed


def alignconv_forward_tc(x, anchors, weight, stride):
    """x [B,C,H,W] bf16/fp16 (channels_last preferred), anchors [B,H,W,5] -> relu(alignconv) as a
    channels_last [B,Co,H,W] tensor of x's dtype."""
    dev = _lib.require_cuda(x, anchors, weight)
    B, C, H, W = x.shape
    Co = weight.size(0)
    xc = _nhwc(x)
    a = anchors.to(torch.float32).contiguous()
    wp = pack_weight(weight, x.dtype)
    out = torch.empty((B, Co, H, W), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_alignconv_forward_tc(_lib.ptr(xc), _lib.ptr(a), _lib.ptr(wp), _lib.ptr(out), B, C, H, W, Co,
                                                  float(stride), _lib.dtype_code(xc), _lib.stream_ptr(dev))
    _lib.check(rc, "alignconv_forward_tc")
    return out


def orconv_forward_tc(x, weight, indices, bias, with_pool=False):
    dev = _lib.require_cuda(x, weight, indices, bias)
    B, C, H, W = x.shape
    O, I, nOri = weight.shape[:3]
    nRot = indices.size(3)
    Co = O * nRot
    if C != I * nOri:
        raise ValueError("input channels (%d) != I*nOrientation (%d)" % (C, I * nOri))
    xc = _nhwc(x)
    wp = pack_weight(weight, x.dtype, indices)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    out = torch.empty((B, Co, H, W), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
    pooled = torch.empty((B, Co // 8, H, W), dtype=x.dtype, device=dev,
                         memory_format=torch.channels_last) if with_pool else None
    with torch.cuda.device(dev):
        rc = _lib.load().s2a_orconv_forward_tc(_lib.ptr(xc), _lib.ptr(wp), _lib.ptr(b), _lib.ptr(out), _lib.ptr(pooled),
                                               B, C, H, W, Co, _lib.dtype_code(xc), _lib.stream_ptr(dev))
    _lib.check(rc, "orconv_forward_tc")
    return (out, pooled) if with_pool else out
