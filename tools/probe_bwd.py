"""Deformable-conv backward at P3 (128x128, 256 -> 256): the tcgen05 dgrad / wgrad (16-bit) against the fp32
column-buffer path (this repo's round-1 backward = the reference's structure) and the reference's own binary."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s2anet_b200 import dcn
from s2anet_b200.conv_tc import deform_conv_dgrad_tc, deform_conv_wgrad_tc
dev = "cuda:0"
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for B in (1, 8):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 256, 128, 128, generator=g).to(dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).to(dev).bfloat16()
    off = (torch.randn(B, 18, 128, 128, generator=g) * 1.2).to(dev)
    gy = torch.randn(B, 256, 128, 128, generator=g).to(dev).bfloat16().contiguous(memory_format=torch.channels_last)
    flops = 2.0 * B * 128 * 128 * 256 * 2304
    ms = t(lambda: deform_conv_dgrad_tc(gy, off, w))
    print("B=%d dgrad tc (grad_input only): %.3f ms (%.0f TFLOP/s)" % (B, ms, flops / ms / 1e9))
    ms = t(lambda: deform_conv_dgrad_tc(gy, off, w, x=x, need_offset_grad=True))
    print("B=%d dgrad tc (+ offset gradient): %.3f ms" % (B, ms))
    ms = t(lambda: deform_conv_wgrad_tc(x, off, gy))
    print("B=%d wgrad tc: %.3f ms (%.0f TFLOP/s)" % (B, ms, flops / ms / 1e9))
    x32, w32, gy32 = x.float().contiguous(), w.float(), gy.float().contiguous()
    e = x32.new_empty(0)
    def old_in():
        gi, gof = torch.zeros_like(x32), torch.zeros_like(off)
        dcn.deform_conv_backward_input_cuda(x32, off, gy32, gi, gof, w32, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
    def old_w():
        gw = torch.zeros_like(w32)
        dcn.deform_conv_backward_parameters_cuda(x32, off, gy32, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
    print("B=%d fp32 column path: backward_input %.3f ms, backward_parameters %.3f ms" % (B, t(old_in, 2), t(old_w, 2)))
    try:
        from oracle import build_oracle
        ref = build_oracle.load_ref_extension("deform_conv_cuda", "gpu")
        xh, wh, gyh, offh = x.half().contiguous(), w.half(), gy.half().contiguous(), off.half()
        eh = xh.new_empty(0)
        def ref_in():
            gi, gof = torch.zeros_like(xh), torch.zeros_like(offh)
            ref.deform_conv_backward_input_cuda(xh, offh, gyh, gi, gof, wh, eh, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
        def ref_w():
            gw = torch.zeros_like(wh)
            ref.deform_conv_backward_parameters_cuda(xh, offh, gyh, gw, eh, eh, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
        print("B=%d reference binary (fp16): backward_input %.3f ms, backward_parameters %.3f ms" % (B, t(ref_in, 2), t(ref_w, 2)))
        def ref_in32():
            gi, gof = torch.zeros_like(x32), torch.zeros_like(off)
            ref.deform_conv_backward_input_cuda(x32, off, gy32, gi, gof, w32, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, B)
        def ref_w32():
            gw = torch.zeros_like(w32)
            ref.deform_conv_backward_parameters_cuda(x32, off, gy32, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, B)
        print("B=%d reference binary (fp32): backward_input %.3f ms, backward_parameters %.3f ms" % (B, t(ref_in32, 2), t(ref_w32, 2)))
    except Exception as ex:
        print("reference binary not available:", ex)
