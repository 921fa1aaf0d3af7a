"""The packed-weight cache of s2anet_b200.conv_tc is keyed on tensor identity + in-place version, never on the
address: a freed weight's address is handed to the next tensor of that size by the allocator."""
import gc

import torch

from s2anet_b200 import conv_tc


def test_cache_hit_needs_same_live_tensor_and_version():
    w = torch.randn(4, 4)
    key = ((id(w),), torch.float32, "t")
    conv_tc._cache_put(key, (w,), "packed-A")
    assert conv_tc._cache_get(key, (w,)) == "packed-A"
    w.add_(1.0)                                   # optimizer step / load_state_dict: version bump invalidates
    assert conv_tc._cache_get(key, (w,)) is None
    conv_tc._cache_put(key, (w,), "packed-B")
    assert conv_tc._cache_get(key, (w,)) == "packed-B"
    other = torch.randn(4, 4)                     # another tensor presenting the same key (id reuse) is refused
    assert conv_tc._cache_get(key, (other,)) is None


def test_cache_entry_dies_with_its_tensor():
    w, b = torch.randn(4, 4), torch.randn(4)
    key = ((id(w), id(b)), torch.float32, "t2")
    conv_tc._cache_put(key, (w, b), "packed")
    assert key in conv_tc._PACK_CACHE
    del b
    gc.collect()
    assert key not in conv_tc._PACK_CACHE


def test_cache_sees_data_reassignment_and_dtype_or_device_moves():
    """`w.data = ...` and Module.to()/half() (which reassign param.data) do not bump `_version`; the stamp
    also carries the storage address, device, dtype, shape and strides."""
    w = torch.nn.Parameter(torch.randn(4, 4))
    key = ((id(w),), torch.float32, "t3")
    conv_tc._cache_put(key, (w,), "packed")
    assert conv_tc._cache_get(key, (w,)) == "packed"
    v0 = w._version
    w.data = torch.randn(4, 4)                    # new storage, same Parameter object, same version
    assert w._version == v0
    assert conv_tc._cache_get(key, (w,)) is None
    conv_tc._cache_put(key, (w,), "packed2")
    m = torch.nn.Conv2d(4, 4, 3)
    key2 = ((id(m.weight),), torch.float32, "t4")
    conv_tc._cache_put(key2, (m.weight,), "packed3")
    m.half()                                      # _apply: param.data reassigned in place, dtype changes
    assert conv_tc._cache_get(key2, (m.weight,)) is None
    conv_tc.invalidate_pack_cache()
    assert not conv_tc._PACK_CACHE


def test_fp32_path_switch_and_layout_helpers_on_the_host():
    """Host-side logic of the fp32 route selection and the layout helpers (no GPU): bad names are refused, the shape
    rule of the tensor-core fp32 kernel, and the CPU fall-backs of the NHWC helpers are plain torch copies."""
    import pytest
    import torch
    from s2anet_b200 import alignconv, conv_tc
    assert alignconv._FORCE_SIMT_F32 is False
    alignconv.set_fp32_path("simt")
    assert alignconv._FORCE_SIMT_F32 is True
    alignconv.set_fp32_path("tf32x3")
    assert alignconv._FORCE_SIMT_F32 is False
    with pytest.raises(ValueError):
        alignconv.set_fp32_path("fp64")
    assert conv_tc.tf32x3_supported(256, 256) and conv_tc.tf32x3_supported(32, 32) and conv_tc.tf32x3_supported(96, 64)
    assert not conv_tc.tf32x3_supported(24, 32) and not conv_tc.tf32x3_supported(32, 48) and not conv_tc.tf32x3_supported(32, 288)
    x = torch.randn(2, 8, 3, 5)
    y = conv_tc._nhwc(x)                                  # CPU tensor: torch's own conversion
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(x, y)
    back = torch.empty_like(x)
    assert conv_tc.nchw_from_nhwc(y, back) is back and back.is_contiguous() and torch.equal(back, x)
    with pytest.raises(ValueError):
        conv_tc._nhwc(torch.zeros(3, 4))
