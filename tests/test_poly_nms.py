"""DOTA result-merging NMS (second half): fp64 polygon IoU + greedy NMS.

reference: DOTA_devkit/polyiou/csrc/polyiou.cpp:9-126, DOTA_devkit/ResultMerge_multi_process.py:62-123.
Bar: polygon IoU BIT-EXACT (fp64, every operation individually rounded in the reference's order), keep lists identical.
The golden vectors were produced by the reference's own polyiou.cpp compiled in place (tests/golden/make_golden_poly.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from s2anet_b200.synth import random_quads as quads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poly_small.npz")


def same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def random_quads(n, seed, extent=300.0):
    return quads(n, np.random.default_rng(seed), extent=extent)


def random_dets(n, seed, extent=400.0):
    rng = np.random.default_rng(seed)
    m = max(1, n // 3)
    seeds = quads(m, rng, extent=extent)
    polys = np.concatenate([seeds, seeds + rng.normal(0, 2, seeds.shape), seeds + rng.normal(0, 12, seeds.shape)])[:n]
    if polys.shape[0] < n:
        polys = np.concatenate([polys, quads(n - polys.shape[0], rng, extent=extent)])
    scores = (rng.permutation(n) + 1.0) / (n + 1.0)
    return np.concatenate([polys, scores[:, None]], 1)


# ---------------------------------------------------------------- CPU: oracle pinned to the reference
def test_oracle_matches_golden_vectors():
    g = np.load(GOLD)
    with np.errstate(invalid="ignore", divide="ignore"):
        assert same_bits(O.poly_iou_pairs(g["p"], g["q"]), g["iou"])
    for thr, key in ((0.1, "keep_01"), (0.3, "keep_03"), (0.5, "keep_05")):
        assert np.array_equal(O.poly_nms(g["dets"], thr), g[key])


def test_known_answers():
    g = np.load(GOLD)
    iou = g["iou"]
    assert iou[0] == 1.0                      # identical squares
    assert iou[1] == 0.0 and iou[2] == 0.0 and iou[3] == 0.0    # edge / corner contact, far apart
    assert iou[4] == 0.25                     # 5 x 5 square inside a 10 x 10 one
    assert iou[5] == 1.0                      # winding does not matter
    assert np.isnan(iou[8])                   # 0 / 0 for two points -- and the NMS drops such a pair (NaN <= thr is false)


def test_oracle_matches_compiled_reference_when_present():
    if O.ref_polyiou() is None:
        pytest.skip("oracle/_ref/libref_polyiou.so not built")
    p = random_quads(3000, 11)
    q = p + np.random.default_rng(12).normal(0, 6, p.shape)
    assert same_bits(O.poly_iou_pairs(p, q), O.ref_poly_iou_pairs(p, q))
    d = random_dets(700, 13)
    for thr in (0.05, 0.3, 0.7):
        assert np.array_equal(O.poly_nms(d, thr), O.ref_py_cpu_nms_poly_fast(d, thr))


def test_nms_edge_cases_oracle():
    assert O.poly_nms(np.zeros((0, 9)), 0.3).size == 0
    one = random_dets(1, 1)
    assert np.array_equal(O.poly_nms(one, 0.3), [0])
    dup = np.repeat(random_dets(1, 2), 5, axis=0)
    dup[:, 8] = [0.1, 0.5, 0.3, 0.9, 0.2]
    assert np.array_equal(O.poly_nms(dup, 0.3), [3])           # five copies: only the best survives


# ---------------------------------------------------------------- GPU: the kernels through the C ABI
@pytest.mark.gpu
def test_gpu_poly_iou_bit_exact():
    from s2anet_b200.poly_nms import iou_poly, iou_poly_pairs
    g = np.load(GOLD)
    out = iou_poly_pairs(torch.from_numpy(g["p"]).cuda(), torch.from_numpy(g["q"]).cuda()).cpu().numpy()
    assert same_bits(out, g["iou"])
    p = random_quads(20000, 21)
    q = p + np.random.default_rng(22).normal(0, 6, p.shape)
    q[:4000] = random_quads(4000, 23)
    out = iou_poly_pairs(torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda()).cpu().numpy()
    assert same_bits(out, O.poly_iou_pairs(p, q))
    assert iou_poly(g["p"][4], g["q"][4]) == 0.25               # polyiou.iou_poly's own call shape


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 450, 3000])
@pytest.mark.parametrize("thr", [0.1, 0.5])
def test_gpu_poly_nms_keep_lists(n, thr):
    from s2anet_b200.poly_nms import poly_nms
    d = random_dets(n, 100 + n)
    keep = poly_nms(torch.from_numpy(d).cuda(), thr).cpu().numpy()
    assert np.array_equal(keep, O.poly_nms(d, thr))


@pytest.mark.gpu
def test_gpu_poly_nms_golden_and_dropin_surface():
    from s2anet_b200 import poly_nms as P
    g = np.load(GOLD)
    for thr, key in ((0.1, "keep_01"), (0.3, "keep_03"), (0.5, "keep_05")):
        keep = P.py_cpu_nms_poly_fast(g["dets"], thr)            # the reference's name: numpy in, list out
        assert isinstance(keep, list) and keep == [int(i) for i in g[key]]
    assert P.py_cpu_nms_poly_fast(np.zeros((0, 9)), 0.3) == []
    assert P.poly_nms(torch.zeros((0, 9), dtype=torch.float64, device="cuda"), 0.3).numel() == 0
    # two point polygons: their axis-aligned boxes do not overlap (w = h = 0), the polygon IoU is never asked: both stay
    z = np.zeros((2, 9))
    z[:, 8] = [0.9, 0.8]
    assert P.py_cpu_nms_poly_fast(z, 0.3) == [0, 1] and np.array_equal(O.poly_nms(z, 0.3), [0, 1])
    # two zero-area "diagonal line" polygons: the boxes overlap, the polygon IoU is 0 / 0 = NaN, and NaN <= thr is
    # false in the reference's `inds = np.where(hbb_ovr <= thresh)`: the later one is dropped
    d = np.zeros((2, 9))
    d[:, :8] = [0, 0, 10, 10, 0, 0, 10, 10]
    d[:, 8] = [0.9, 0.8]
    assert P.py_cpu_nms_poly_fast(d, 0.3) == [0] and np.array_equal(O.poly_nms(d, 0.3), [0])
    with pytest.raises(ValueError):
        P.poly_nms(torch.zeros((4, 5), dtype=torch.float64, device="cuda"), 0.3)
