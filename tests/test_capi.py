"""CPU tests of the C-ABI boundary: the library builds, loads, exports every symbol that
include/s2a_b200.h declares, and validates arguments before touching the GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from s2anet_b200 import build
    build.build()
    from s2anet_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "s2a_b200.h")).read()
    return sorted(set(re.findall(r"\b(s2a_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "libs2a_b200.so does not export %s" % s


def test_python_prototypes_cover_header(lib):
    from s2anet_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_version_and_error_string(lib):
    assert lib.s2a_version() >= 100
    assert isinstance(lib.s2a_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # every call below is rejected (or is a no-op) before any CUDA work is enqueued
    assert lib.s2a_box_iou_rotated(None, -1, None, 0, 1, None, 0, 0, 0, 0, None) == -1
    assert b"negative" in lib.s2a_last_error()
    assert lib.s2a_box_iou_rotated(None, 0, None, 5, 1, None, 5, 0, 0, 0, None) == 0        # n == 0: no-op
    assert lib.s2a_box_iou_rotated(None, 4, None, 5, 1, None, 5, 3, 2, 0, None) == -1       # bad row range
    assert lib.s2a_nms_rotated_workspace_bytes(0) > 0
    assert lib.s2a_nms_rotated_workspace_bytes(2000) > 2000 * 32 * 8
    assert lib.s2a_deform_conv_forward_f32(None, None, None, None, 1, 8, 2, 2, 8, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0,
                                           None) == -1
    assert b"smaller than kernel" in lib.s2a_last_error()
    assert lib.s2a_arf_forward(None, None, None, 4, 4, 16, 3, 3, 8, 0, None) == -1          # nOri*9 > 72
    assert lib.s2a_ri_pool_forward(None, None, 1, 12, 4, 8, 0, None) == -1                  # 12 % 8 != 0
    assert lib.s2a_multiclass_nms_rotated_workspace_bytes(5344, 15, 1) > 0
    # polygon NMS (DOTA result merging): n == 0 needs no pointers but num_keep_out; a row is 8 coordinates + score
    assert lib.s2a_poly_nms_workspace_bytes(0) > 0
    assert lib.s2a_poly_nms_workspace_bytes(3000) > 3000 * 47 * 8 + 3000 * 104
    assert lib.s2a_poly_nms(None, 9, 5, 0.3, None, None, None, 0, None) == -1
    assert b"num_keep_out" in lib.s2a_last_error()
    assert lib.s2a_poly_iou_pairs(None, None, 0, None, None) == 0
    assert lib.s2a_poly_iou_pairs(None, None, -1, None, None) == -1
    # conv entry points: shape constraints are reported, never rerouted
    assert lib.s2a_select_decode_workspace_bytes(0, None, None, 8) > 0


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the reference raises NotImplementedError for CPU deform conv
    (models/dcn/deform_conv.py:58-59); every op here does the same."""
    import torch
    from s2anet_b200.alignconv import AlignConv
    from s2anet_b200.box_iou_rotated import box_iou_rotated
    from s2anet_b200.nms_rotated import ml_nms_rotated, multiclass_nms_rotated, nms_rotated
    from s2anet_b200.orn import ORConv2d, RotationInvariantPooling
    b = torch.rand(4, 5)
    with pytest.raises(NotImplementedError):
        box_iou_rotated(b, b)
    with pytest.raises(NotImplementedError):
        nms_rotated(torch.rand(4, 6), 0.5)
    with pytest.raises(NotImplementedError):
        ml_nms_rotated(b, torch.rand(4), torch.zeros(4), 0.5)
    with pytest.raises(NotImplementedError):
        multiclass_nms_rotated(b, torch.rand(4, 3))
    with pytest.raises(NotImplementedError):
        AlignConv(8, 8)(torch.rand(1, 8, 4, 4), torch.rand(1, 4, 4, 5), 8)
    with pytest.raises(NotImplementedError):
        RotationInvariantPooling(16, 8)(torch.rand(1, 16, 2, 2))
    from s2anet_b200.assign import assign_labels
    from s2anet_b200.decode import fam_decode, select_decode
    from s2anet_b200.poly_nms import iou_poly_pairs, poly_nms
    with pytest.raises(NotImplementedError):
        poly_nms(torch.rand(4, 9, dtype=torch.float64), 0.3)
    with pytest.raises(NotImplementedError):
        iou_poly_pairs(torch.rand(4, 8, dtype=torch.float64), torch.rand(4, 8, dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        fam_decode([torch.rand(1, 5, 4, 4)], [8])
    with pytest.raises(NotImplementedError):
        select_decode([torch.rand(1, 15, 4, 4)], [torch.rand(1, 5, 4, 4)], [torch.rand(1, 4, 4, 5)])
    with pytest.raises(NotImplementedError):
        assign_labels(torch.rand(8, 5), torch.rand(2, 5))
    m = ORConv2d(8, 2, 3, padding=1, arf_config=(1, 8))
    assert tuple(m.weight.shape) == (2, 8, 1, 3, 3) and tuple(m.bias.shape) == (16,)
    assert tuple(m.indices.shape) == (1, 3, 3, 8) and m.indices.dtype == torch.uint8
    assert nms_rotated(torch.zeros(0, 6), 0.5).shape == (0, 6)        # reference returns the bare tensor


def test_module_surface_matches_reference_checkpoint_contract():
    """parameter / buffer names and shapes that checkpoints depend on."""
    import torch
    from s2anet_b200.alignconv import AlignConv
    from s2anet_b200.orn import ORConv2d
    ac = AlignConv(256, 256, kernel_size=3)
    assert [k for k, _ in ac.named_parameters()] == ["deform_conv.weight"]
    assert tuple(ac.deform_conv.weight.shape) == (256, 256, 3, 3)
    oc = ORConv2d(256, 32, kernel_size=3, padding=1, arf_config=(1, 8))
    sd = oc.state_dict()
    assert tuple(sd["weight"].shape) == (32, 256, 1, 3, 3)
    assert tuple(sd["bias"].shape) == (256,)
    assert tuple(sd["indices"].shape) == (1, 3, 3, 8) and sd["indices"].dtype == torch.uint8


def test_torch_extension_binding_loads_and_exports_the_reference_names():
    """The thin torch extension over the C ABI (s2anet_b200/_s2a_torch.so, built by __graft_entry__.build()): loads
    without a GPU, reports the library's ABI version and exports the reference's extension-level function names
    (utils/*/src/*.h, models/orn/src/vision.cpp:7-12)."""
    from s2anet_b200 import _lib, _torch_ext, build
    build.build()
    build.build_torch_ext()
    m = _torch_ext.module()
    assert m is not None
    assert m.abi_version() == _lib.load().s2a_version()
    for name in ("box_iou_rotated", "nms_rotated", "ml_nms_rotated", "arf_forward", "arf_backward"):
        assert callable(getattr(m, name))
    import pytest
    import torch
    with pytest.raises(RuntimeError):                   # CPU tensors: refused loudly, like every other entry of the library
        m.box_iou_rotated(torch.zeros(2, 5), torch.zeros(3, 5))
