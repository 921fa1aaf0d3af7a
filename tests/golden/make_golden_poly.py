"""Golden vectors for the DOTA result-merging NMS, produced by the REFERENCE's own polygon IoU
(DOTA_devkit/polyiou/csrc/polyiou.cpp, compiled in place into oracle/_ref/libref_polyiou.so) and
py_cpu_nms_poly_fast restated line by line on top of it (oracle.ref_py_cpu_nms_poly_fast):

    python tests/golden/make_golden_poly.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def adversarial():
    sq = np.array([0, 0, 10, 0, 10, 10, 0, 10], dtype=np.float64)
    cases = [
        (sq, sq),                                                    # identical
        (sq, sq + 10.0 * np.array([1, 0] * 4)),                      # sharing an edge
        (sq, sq + 10.0),                                             # sharing a corner
        (sq, sq + 50.0),                                             # far apart
        (sq, sq * 0.5 + 2.5),                                        # nested
        (sq, sq.reshape(4, 2)[::-1].reshape(-1)),                    # same square, clockwise
        (sq, np.array([5, -5, 15, 5, 5, 15, -5, 5], dtype=np.float64)),   # rotated by 45 degrees
        (sq, np.array([0, 0, 10, 0, 10, 0, 0, 0], dtype=np.float64)),     # zero-area polygon
        (np.zeros(8), np.zeros(8)),                                  # two points: 0 / 0
        (sq, sq + 1e-9),                                             # shifted by less than the reference's eps
        (sq * 1e4, sq * 1e4 + 3.0),                                  # large coordinates
        (sq, np.array([2, 2, 30, 3, 31, 9, 1, 8], dtype=np.float64)),     # general convex quadrilateral
        (np.array([0, 0, 10, 0, 0, 10, 10, 10], dtype=np.float64), sq),   # self-intersecting ("bow tie") polygon
    ]
    return np.stack([c[0] for c in cases]), np.stack([c[1] for c in cases])


def main():
    from oracle import build_oracle, oracle as O
    from s2anet_b200.synth import random_quads as quads
    build_oracle.build_oracle()
    assert build_oracle.build_ref_polyiou() is not None, "needs /root/reference"
    rng = np.random.default_rng(7)
    p = quads(243, rng)
    q = p + rng.normal(0, 5, p.shape)
    q[:60] = quads(60, rng)                                          # mostly disjoint pairs
    ap, aq = adversarial()
    p, q = np.concatenate([ap, p]), np.concatenate([aq, q])
    with np.errstate(invalid="ignore", divide="ignore"):
        iou = O.ref_poly_iou_pairs(p, q)
    # detections: 150 seeds, each with a jittered copy and a shifted copy; distinct scores
    seeds = quads(150, rng, extent=400.0)
    dets_p = np.concatenate([seeds, seeds + rng.normal(0, 2, seeds.shape), seeds + rng.normal(0, 12, seeds.shape)])
    scores = (rng.permutation(dets_p.shape[0]) + 1.0) / (dets_p.shape[0] + 1.0)
    dets = np.concatenate([dets_p, scores[:, None]], 1)
    out = dict(p=p, q=q, iou=iou, dets=dets)
    for thr in (0.1, 0.3, 0.5):
        out["keep_%02d" % int(thr * 10)] = O.ref_py_cpu_nms_poly_fast(dets, thr)
    np.savez_compressed(os.path.join(HERE, "poly_small.npz"), **out)
    print("pairs", p.shape[0], "nan", int(np.isnan(iou).sum()), "dets", dets.shape[0],
          "kept", [int(out["keep_%02d" % t].size) for t in (1, 3, 5)])


if __name__ == "__main__":
    main()
