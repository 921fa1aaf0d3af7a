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
