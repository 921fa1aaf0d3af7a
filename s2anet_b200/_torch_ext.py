"""Loader of the thin torch-extension binding (s2anet_b200/_s2a_torch.so, built by `python -m s2anet_b200.build` from
csrc/torch_binding.cpp): the reference's extension-level functions (`box_iou_rotated`, `nms_rotated`, `ml_nms_rotated`,
`arf_forward`, `arf_backward`) as pybind11 functions that take and return at::Tensors and make one C-ABI call each --
the shape of the reference's own boundary, at a third of the per-call cost of the ctypes wrappers.

`module()` returns the extension or None when it has not been built (the ctypes wrappers of the package are used then);
a built extension whose ABI version differs from the loaded library is an error, not a silent fallback."""
import importlib.util
import os

from . import _lib

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_s2a_torch.so")
_MOD = None
_TRIED = False


def module():
    global _MOD, _TRIED
    if _TRIED:
        return _MOD
    _TRIED = True
    if os.environ.get("S2A_NO_TORCH_EXT") or not os.path.exists(_PATH):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    lib = _lib.load()                   # libs2a_b200.so first: the extension's NEEDED entry resolves to the loaded copy
    spec = importlib.util.spec_from_file_location("_s2a_torch", _PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.abi_version() != lib.s2a_version():
        raise RuntimeError("s2anet_b200: _s2a_torch.so was built against ABI %d, libs2a_b200.so is %d -- rebuild with "
                           "python -m s2anet_b200.build --force" % (mod.abi_version(), lib.s2a_version()))
    _MOD = mod
    return _MOD
