"""Seeded synthetic inputs for the S2ANet hot path: DOTA-split-shaped
rotated boxes, FPN feature maps and refined anchors.  Pure numpy/torch-CPU generators so the
same inputs are seen by the CUDA path, the oracle and the golden-vector generator."""
import math

import numpy as np

STRIDES = (8, 16, 32, 64, 128)


def level_shapes(img=1024, strides=STRIDES):
    return [(img // s, img // s) for s in strides]


def clustered_boxes(n_seed=400, rep=5, seed=0, img=1024.0):
    """Config 1: n_seed seed boxes (centre U[0,img)^2, w,h logU[8,256], theta U[-pi/4,3pi/4)), each
    replicated `rep` times with jitter so that NMS really suppresses; distinct scores; labels in [0,15)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, img, (n_seed, 2))
    wh = np.exp(rng.uniform(math.log(8), math.log(256), (n_seed, 2)))
    th = rng.uniform(-math.pi / 4, 3 * math.pi / 4, (n_seed, 1))
    base = np.concatenate([c, wh, th], 1)
    b = np.repeat(base, rep, 0)
    n = b.shape[0]
    m = np.minimum(b[:, 2], b[:, 3])[:, None]
    b[:, 0:2] += rng.normal(0, 0.1, (n, 2)) * m
    b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (n, 2)))
    b[:, 4] += rng.normal(0, 0.1, n)
    scores = rng.uniform(0, 1, n) + np.arange(n) * 1e-7
    labels = rng.integers(0, 15, n)
    perm = rng.permutation(n)
    return (b[perm].astype(np.float32), scores[perm].astype(np.float32), labels[perm].astype(np.float32))


def adversarial_boxes():
    """Degenerate / touching / huge boxes that exercise every early-out of the IoU routine."""
    h = math.pi / 2
    rows = [
        (100, 100, 40, 20, 0.0), (100, 100, 40, 20, 0.0),            # identical
        (100, 100, 40, 20, h), (100, 100, 20, 40, 0.0),              # same rectangle, theta off by pi/2
        (100, 100, 40, 20, math.pi), (100, 100, 40, 20, -math.pi),   # theta off by pi / 2pi
        (140, 100, 40, 20, 0.0),                                     # touching edge of the first
        (120, 110, 40, 20, 0.0),                                     # half overlap
        (100, 100, 0, 20, 0.3), (100, 100, 1e-9, 1e-9, 0.3),         # zero / sub-epsilon area
        (100, 100, 1e4, 1e4, 0.7), (5000, 5000, 1e4, 3e3, -0.2),     # huge
        (100, 100, 40, 20, 1e-8), (100, 100, 40, 20, math.pi / 4),   # tiny angle, 45 deg
        (100.5, 100.25, 3, 300, 1.1), (100, 100, 300, 3, 1.1),       # thin, crossing
        (0, 0, 1, 1, 0.0), (0.5, 0.5, 1, 1, 0.0), (1, 1, 1, 1, 0.0), # unit squares (1/7 pair), corner touch
        (100, 100, 10, 10, 0.2), (100, 100, 80, 60, 0.9),            # containment
    ]
    return np.asarray(rows, np.float32)


def dota_like_gt(n, seed, img=1024.0):
    """Config 4 ground truth: centre U, long side logU[10,400], aspect U[1,8], theta U[-pi/4,3pi/4)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, img, (n, 2))
    long_side = np.exp(rng.uniform(math.log(10), math.log(400), n))
    aspect = rng.uniform(1, 8, n)
    th = rng.uniform(-math.pi / 4, 3 * math.pi / 4, n)
    return np.stack([c[:, 0], c[:, 1], long_side, long_side / aspect, th], 1).astype(np.float32)


def grid_anchors(H, W, stride, scale=4.0):
    """models/anchors.py:75-126: centre s*i + 0.5*(s-1), w = h = scale*s, theta = 0.  [H, W, 5]."""
    xs = np.arange(W, dtype=np.float32) * stride + 0.5 * (stride - 1)
    ys = np.arange(H, dtype=np.float32) * stride + 0.5 * (stride - 1)
    a = np.zeros((H, W, 5), np.float32)
    a[..., 0] = xs[None, :]
    a[..., 1] = ys[:, None]
    a[..., 2] = scale * stride
    a[..., 3] = scale * stride
    return a


def refined_anchors(B, H, W, stride, seed, scale=4.0):
    """Config 2: grid anchors perturbed like a FAM output: dxy ~ N(0, stride) px,
    w,h *= exp(N(0,0.5)), theta ~ U[-pi/4, 3pi/4).  [B, H, W, 5] float32."""
    rng = np.random.default_rng(seed)
    a = np.broadcast_to(grid_anchors(H, W, stride, scale), (B, H, W, 5)).copy()
    a[..., 0:2] += rng.normal(0, float(stride), (B, H, W, 2)).astype(np.float32)
    a[..., 2:4] *= np.exp(rng.normal(0, 0.5, (B, H, W, 2))).astype(np.float32)
    a[..., 4] = rng.uniform(-math.pi / 4, 3 * math.pi / 4, (B, H, W)).astype(np.float32)
    return a.astype(np.float32)


def all_level_anchors(B, seed, img=1024, strides=STRIDES):
    """Refined anchors of all five levels concatenated: [B, 21824, 5] at 1024^2."""
    parts = []
    for li, s in enumerate(strides):
        H = W = img // s
        parts.append(refined_anchors(B, H, W, s, seed * 16 + li).reshape(B, H * W, 5))
    return np.concatenate(parts, 1)


def random_quads(n, rng, extent=300.0, lo=8.0, hi=120.0):
    """n rotated rectangles as 8-number polygons, random orientation (both windings occur)."""
    c = rng.uniform(0, extent, (n, 2))
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    t = rng.uniform(-np.pi, np.pi, n)
    dx = np.stack([np.cos(t), np.sin(t)], 1)
    dy = np.stack([-np.sin(t), np.cos(t)], 1)
    corners = [c + s1 * dx * w[:, None] / 2 + s2 * dy * h[:, None] / 2 for s1, s2 in ((-1, -1), (1, -1), (1, 1), (-1, 1))]
    p = np.concatenate(corners, 1)
    flip = rng.random(n) < 0.5                  # clockwise vertex order for half of them
    p[flip] = p[flip].reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8)
    return p
