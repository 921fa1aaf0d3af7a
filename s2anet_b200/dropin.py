"""Drop-in installer: make the reference tree's own Python (models/detector.py, models/head.py,
val.py ...) run on the B200 ops WITHOUT editing it.

The reference imports six pybind11 extension modules by name:

    models.dcn.deform_conv_cuda          deform_conv_forward_cuda / backward_* / modulated_*
    models.dcn.deform_pool_cuda          (imported by models/dcn/__init__.py:4, never called by S2ANet)
    models.orn.orn_cuda                  arf_forward / arf_backward / rie_forward / rie_backward
    DOTA_devkit.polyiou.polyiou          VectorDouble / iou_poly (the SWIG module of the DOTA result merging)
    utils.box_iou_rotated.box_iou_rotated_cuda      box_iou_rotated
    utils.nms_rotated.nms_rotated_cuda              nms_rotated
    utils.ml_nms_rotated.ml_nms_rotated_cuda        ml_nms_rotated

`install()` registers Python modules with exactly those names and function signatures in
sys.modules.  `box_iou_rotated`, `nms_rotated`, `ml_nms_rotated` and `arf_forward` go through the thin torch
extension (`_s2a_torch.so`, csrc/torch_binding.cpp: at::Tensor in, one C-ABI call, at::Tensor out) when it has been
built, everything else -- and everything when it has not -- through the ctypes binding of libs2a_b200.so.  `deform_conv_forward_cuda`
takes fp32, fp16 (the reference's `AT_DISPATCH_FLOATING_TYPES_AND_HALF`, what val.py's `model.half()`
reaches) and bf16 tensors; the two deform-conv backward entries and `arf_backward` are implemented
(fp32 accumulate).  Functions outside the hot path (modulated DCN, PS-RoI pooling, RIE) exist so imports
succeed and raise NotImplementedError when called.  `accelerate(model)` additionally swaps the *fused* forward
paths in (AlignConv without an offset tensor, ORConv2d with the ARF folded into the weight load and
the orientation pooling into the epilogue) while keeping every parameter / buffer name.
"""
import sys
import types


def _not_built(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("s2anet_b200: %s is outside the accelerated hot path (SURVEY.md section 8f)" % name)
    fn.__name__ = name
    return fn


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__s2a_b200__ = True
    sys.modules[name] = m
    return m


def install(matplotlib_stub=True):
    """Register the six extension-module names.  Call BEFORE importing the reference's `models` /
    `utils` packages (with the reference root on sys.path)."""
    from . import dcn, nms_rotated, orn
    from .box_iou_rotated import box_iou_rotated

    _module("models.dcn.deform_conv_cuda",
            deform_conv_forward_cuda=dcn.deform_conv_forward_cuda,
            deform_conv_backward_input_cuda=dcn.deform_conv_backward_input_cuda,
            deform_conv_backward_parameters_cuda=dcn.deform_conv_backward_parameters_cuda,
            modulated_deform_conv_cuda_forward=_not_built("modulated_deform_conv_cuda_forward"),
            modulated_deform_conv_cuda_backward=_not_built("modulated_deform_conv_cuda_backward"))
    _module("models.dcn.deform_pool_cuda",
            deform_psroi_pooling_cuda_forward=_not_built("deform_psroi_pooling_cuda_forward"),
            deform_psroi_pooling_cuda_backward=_not_built("deform_psroi_pooling_cuda_backward"))
    _module("models.orn.orn_cuda", arf_forward=orn.arf_forward, arf_backward=orn.arf_backward,
            rie_forward=_not_built("rie_forward"), rie_backward=_not_built("rie_backward"))
    _module("utils.box_iou_rotated.box_iou_rotated_cuda", box_iou_rotated=box_iou_rotated)
    _module("utils.nms_rotated.nms_rotated_cuda", nms_rotated=nms_rotated.nms_rotated_op)
    _module("utils.ml_nms_rotated.ml_nms_rotated_cuda", ml_nms_rotated=nms_rotated.ml_nms_rotated)
    # DOTA result merging (DOTA_devkit/ResultMerge_multi_process.py:15 imports the SWIG module, whose generated
    # wrapper needs the `imp` module Python 3.12 no longer has): same two names, GPU polygon IoU underneath.  The fast
    # path is poly_nms.py_cpu_nms_poly_fast, which replaces the whole per-pair Python loop.
    from . import poly_nms
    pm = _module("DOTA_devkit.polyiou.polyiou", VectorDouble=poly_nms.VectorDouble, iou_poly=poly_nms.iou_poly)
    pkg = _module("DOTA_devkit.polyiou", polyiou=pm)
    pkg.__path__ = []
    if matplotlib_stub:
        try:
            import matplotlib  # noqa: F401
        except ImportError:       # utils/metrics.py:10 imports pyplot at module scope; plotting is off-path
            mpl = types.ModuleType("matplotlib")
            mpl.__path__ = []
            mpl.use = lambda *a, **k: None
            mpl.rc = lambda *a, **k: None
            sys.modules["matplotlib"] = mpl
            for sub in ("pyplot", "patches", "colors", "cm"):
                sm = types.ModuleType("matplotlib." + sub)

                def _lazy(name, _sub=sub):
                    if name.startswith("__"):
                        raise AttributeError(name)
                    return _not_built("matplotlib.%s.%s" % (_sub, name))
                sm.__getattr__ = _lazy
                sys.modules["matplotlib." + sub] = sm
                setattr(mpl, sub, sm)
    return sorted(n for n, m in sys.modules.items() if getattr(m, "__s2a_b200__", False))


def accelerate(model):
    """Swap the fused kernels into an already-built reference model (or any module tree that uses
    the reference's AlignConv / ORConv2d / RotationInvariantPooling classes by name): parameters and
    buffers are shared, not copied, so checkpoints keep loading."""
    from . import alignconv as s2a_align
    from . import orn as s2a_orn
    swapped = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            cls = type(child).__name__
            if cls == "AlignConv" and not isinstance(child, s2a_align.AlignConv):
                dc = child.deform_conv
                new = s2a_align.AlignConv(dc.in_channels, dc.out_channels, kernel_size=child.kernel_size,
                                          deformable_groups=dc.deformable_groups)
                new.deform_conv.weight = dc.weight
                setattr(parent, name, new)
                swapped += 1
            elif cls == "ORConv2d" and not isinstance(child, s2a_orn.ORConv2d):
                new = s2a_orn.ORConv2d(child.in_channels, child.out_channels, child.kernel_size,
                                       arf_config=(child.nOrientation, child.nRotation), stride=child.stride,
                                       padding=child.padding, dilation=child.dilation, groups=child.groups,
                                       bias=child.bias is not None)
                new.weight = child.weight
                if child.bias is not None:
                    new.bias = child.bias
                new.indices = child.indices
                new.fuse_pool = hasattr(parent, "or_pool")
                setattr(parent, name, new)
                swapped += 1
            elif cls == "RotationInvariantPooling" and not isinstance(child, s2a_orn.RotationInvariantPooling):
                setattr(parent, name, s2a_orn.RotationInvariantPooling(child.nInputPlane, child.nOrientation))
                swapped += 1
    return swapped
