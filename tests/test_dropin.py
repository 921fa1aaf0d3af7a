"""The drop-in boundary against the REAL reference tree (authoring container only: the test is
skipped where /root/reference is not mounted, e.g. on the GPU box).  The reference's own Python
files are imported unmodified on top of the shim modules; on CPU we can check import + surface +
error behaviour, on the GPU box the same path is exercised through tests/test_gpu_*."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


@pytest.fixture()
def ref_on_path():
    saved = dict(sys.modules)
    sys.path.insert(0, REF)
    yield
    sys.path.remove(REF)
    for k in list(sys.modules):
        if k not in saved and (k.split(".")[0] in ("models", "utils", "matplotlib")):
            del sys.modules[k]


def test_reference_python_imports_on_the_shims(ref_on_path):
    import torch
    from s2anet_b200 import dropin
    names = dropin.install()
    assert "models.dcn.deform_conv_cuda" in names and "utils.ml_nms_rotated.ml_nms_rotated_cuda" in names
    # the reference's own wrappers, unmodified
    from models.alignconv import AlignConv
    from models.dcn import DeformConv, deform_conv
    from models.orn import ORConv2d, RotationInvariantPooling
    from utils.bbox_nms_rotated import multiclass_nms_rotated
    from utils.box_iou_rotated import box_iou_rotated
    from utils.ml_nms_rotated import ml_nms_rotated
    from utils.nms_rotated import nms_rotated
    import models.dcn.deform_conv_cuda as ext
    assert ext.__s2a_b200__ and callable(ext.deform_conv_forward_cuda)
    ac = AlignConv(256, 256, kernel_size=3)
    assert tuple(ac.deform_conv.weight.shape) == (256, 256, 3, 3)
    oc = ORConv2d(256, 32, kernel_size=3, padding=1, arf_config=(1, 8))
    assert tuple(oc.weight.shape) == (32, 256, 1, 3, 3)
    # CPU tensors: the reference raises NotImplementedError for deform conv (deform_conv.py:58-59);
    # every B200 op refuses them the same way instead of falling back
    with pytest.raises(NotImplementedError):
        ac(torch.rand(1, 256, 4, 4), torch.rand(1, 4, 4, 5), 8)
    with pytest.raises(NotImplementedError):
        box_iou_rotated(torch.rand(3, 5), torch.rand(3, 5))
    with pytest.raises(NotImplementedError):
        multiclass_nms_rotated(torch.rand(4, 5), torch.rand(4, 15))
    with pytest.raises(NotImplementedError):
        ext.modulated_deform_conv_cuda_forward()
    # the reference's empty-input short cut never reaches the extension
    assert nms_rotated(torch.zeros(0, 6), 0.5).shape == (0, 6)


def test_reference_head_builds_and_accelerates(ref_on_path):
    from s2anet_b200 import dropin
    dropin.install()
    from models.head import S2ANetHead as RefHead          # the reference's head, unmodified
    head = RefHead(num_classes=15)
    names = [n for n, _ in head.named_parameters()]
    assert "align_conv.deform_conv.weight" in names and "or_conv.weight" in names and "or_conv.bias" in names
    assert tuple(head.state_dict()["or_conv.indices"].shape) == (1, 3, 3, 8)
    w_before = head.align_conv.deform_conv.weight
    n = dropin.accelerate(head)
    assert n == 3
    from s2anet_b200.alignconv import AlignConv
    from s2anet_b200.orn import ORConv2d
    assert isinstance(head.align_conv, AlignConv) and isinstance(head.or_conv, ORConv2d)
    assert head.align_conv.deform_conv.weight is w_before           # shared, not copied
    assert sorted(n for n, _ in head.named_parameters()) == sorted(names)
    # our own head mirrors the same state-dict keys (reference checkpoints load into it)
    from s2anet_b200.head import S2ANetHead
    mine = S2ANetHead(15)
    assert sorted(mine.state_dict().keys()) == sorted(head.state_dict().keys())
    mine.load_state_dict(head.state_dict())


def test_reference_result_merge_runs_on_the_polyiou_shim(ref_on_path, monkeypatch):
    """DOTA_devkit/ResultMerge_multi_process.py, unmodified, on the `DOTA_devkit.polyiou.polyiou` shim.  Its own
    py_cpu_nms_poly_fast (:62-123), fed with the CPU oracle's iou_poly, must reproduce the golden keep lists -- which
    pins both the oracle's NMS restatement and the golden generator's against the reference function itself."""
    import types
    import numpy as np
    from oracle import oracle as O
    from s2anet_b200 import dropin
    dropin.install()
    for name in ("shapely", "shapely.geometry"):            # dota_utils.py imports shapely at module scope (absent here)
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]
    import DOTA_devkit.ResultMerge_multi_process as R
    assert R.polyiou.__s2a_b200__ and callable(R.polyiou.iou_poly) and callable(R.polyiou.VectorDouble)
    monkeypatch.setattr(R.polyiou, "iou_poly",
                        lambda p, q: float(O.poly_iou_pairs(np.array([list(p)]), np.array([list(q)]))[0]))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poly_small.npz"))
    for thr, key in ((0.1, "keep_01"), (0.5, "keep_05")):
        keep = R.py_cpu_nms_poly_fast(g["dets"], thr)
        assert [int(i) for i in keep] == [int(i) for i in g[key]]
    for k in list(sys.modules):
        if k.startswith("DOTA_devkit"):
            del sys.modules[k]
