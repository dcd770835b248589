"""Seeded synthetic crowds for tests and bench.py.

No dataset can be fetched, so every workload is generated: head positions are a
mixture of a uniform background and Gaussian clusters (like a crowd photo),
continuous coordinates (no duplicates, hence no kNN / top-k ties), targets
U[0.3, 1] as produced by datasets/bay_dataset.py:93-96, predicted density
|N(0,1)| * N/M (non-negative like BL_VGG's abs output, models/baselines/BL.py).
Seeds are ``1000 * config + image_index``.
"""
import math

import numpy as np


def crowd_points(rng, n, width, height, dtype=np.float32, outside="redraw"):
    """[n,2] (col,row) image-pixel coordinates inside [0,W) x [0,H).

    ``outside``: what happens to cluster heads that land outside the image.  "redraw" (the default, every workload and
    test input): drawn again uniformly.  "clip": clamped onto the border -- the first version of this generator, kept
    ONLY because tests/golden/bl_*.npz and dmap_cases.npz were drawn with it (tests/golden/make_golden.py passes it, so
    that the committed fixtures regenerate bit for bit; tests/test_golden_pin.py checks that)."""
    if n == 0:
        return np.zeros((0, 2), dtype=dtype)
    k = int(rng.integers(3, 21))
    centres = rng.uniform([0, 0], [width, height], size=(k, 2))
    spread = rng.uniform(20.0, 150.0, size=k)
    weights = rng.dirichlet(np.ones(k))
    n_bg = int(round(0.15 * n))
    which = rng.choice(k, size=n - n_bg, p=weights)
    pts = centres[which] + rng.normal(size=(n - n_bg, 2)) * spread[which, None]
    bg = rng.uniform([0, 0], [width, height], size=(n_bg, 2))
    pts = np.concatenate([pts, bg], 0)
    if outside == "clip":
        pts[:, 0] = np.clip(pts[:, 0], 0, np.nextafter(np.float32(width), np.float32(0)))
        pts[:, 1] = np.clip(pts[:, 1], 0, np.nextafter(np.float32(height), np.float32(0)))
    else:
        # heads that fall outside the image are re-drawn uniformly (clipping would pile duplicates on the
        # border, and duplicates make kNN / top-k tie order implementation-defined)
        out = (pts[:, 0] < 0) | (pts[:, 0] >= width - 1) | (pts[:, 1] < 0) | (pts[:, 1] >= height - 1)
        pts[out] = rng.uniform([0, 0], [width - 1, height - 1], size=(int(out.sum()), 2))
    rng.shuffle(pts, axis=0)
    return pts.astype(dtype)


def log_uniform_count(rng, lo, hi):
    return int(round(math.exp(rng.uniform(math.log(lo), math.log(hi)))))


def bl_image(seed, n, width, height, stride, outside="redraw"):
    """One image's (points [n,2] f32, targets [n] f32, density [H',W'] f32, st_size)."""
    rng = np.random.default_rng(seed)
    pts = crowd_points(rng, n, width, height, outside=outside)
    targets = rng.uniform(0.3, 1.0, size=n).astype(np.float32)
    hp, wp = height // stride, width // stride
    dens = np.abs(rng.normal(size=(hp, wp))).astype(np.float32)
    dens *= np.float32(max(n, 1) / (hp * wp))
    return pts, targets, dens, float(min(width, height))


def bl_batch(config, counts, width, height, stride=8, first_image=0, outside="redraw"):
    """Batch for BL: lists of per-image points/targets, density [B,1,H',W'], st_sizes [B]."""
    pts, tgt, den, st = [], [], [], []
    for i, n in enumerate(counts):
        p, t, d, s = bl_image(1000 * config + first_image + i, n, width, height, stride, outside)
        pts.append(p)
        tgt.append(t)
        den.append(d)
        st.append(s)
    density = np.stack(den, 0)[:, None]
    return pts, tgt, density, np.asarray(st, dtype=np.float32)


def config_counts(config, batch=None):
    """Per-image point counts of BASELINE.json configs 1-3."""
    rng = np.random.default_rng(1000 * config + 999)
    if config == 1:
        return [200]
    if config == 2:
        return [log_uniform_count(rng, 50, 3000) for _ in range(batch or 8)]
    if config == 3:
        b = batch or 16
        counts = [log_uniform_count(rng, 500, 12000) for _ in range(b)]
        counts[0] = 12000
        return counts
    raise ValueError(config)


CONFIG_SHAPES = {1: (1024, 768), 2: (1024, 1024), 3: (2048, 1536)}  # (W, H) image pixels


def config4_images(n_images=64, seed=4000):
    """BASELINE config 4: ``n_images`` JHU-Crowd-shaped images as ((H, W), points [N,2]) -- sides U{512..2048}
    (utils/preprocess_data.py:536-539 bounds), head counts exp(U[ln 1, ln 25000]) plus one empty and one 3-head image,
    float64 points (the JHU path, preprocess_data.py:54-55) with every eighth image float32 (the QNRF path, :70)."""
    rng = np.random.default_rng(seed)
    images = []
    for i in range(n_images):
        h, w = int(rng.integers(512, 2049)), int(rng.integers(512, 2049))
        n = 0 if i == 0 else 3 if i == 1 else log_uniform_count(rng, 1, 25000)
        dtype = np.float32 if i % 8 == 7 else np.float64
        images.append(((h, w), crowd_points(np.random.default_rng(seed + i), n, w, h, dtype=dtype)))
    return images


CONFIG5_SHAPES = [(8, 64, 160, 160), (8, 256, 80, 80), (8, 512, 40, 40)]  # (B, C, H, W) of the three whitened layers
