"""Deterministic synthetic "S3DIS-shape" scenes.

No S3DIS data or network exists on the build/GPU boxes, so benches and tests use rooms made of
planar surfaces (floor, ceiling, walls, axis-aligned furniture boxes, boards) sampled at
~2500 pts/m² with 5 mm normal jitter, then pushed through the reference's own preprocessing
order: shift to min=0, voxel-dedupe at 0.04 m (first point per voxel), crop the N points nearest
to a seeded centre, shuffle, re-shift (reference: pytorch/util/data_util.py:45-76,
pytorch/util/voxelize.py:38-56).  Labels follow surface/object ids (13 classes) so that label
boundaries exist for the contrastive-boundary loss.
"""
import numpy as np

NUM_CLASSES = 13


def _rect(rng, origin, u, v, density, jitter, normal):
    area = np.linalg.norm(u) * np.linalg.norm(v)
    cnt = max(int(area * density), 4)
    a = rng.random((cnt, 1))
    b = rng.random((cnt, 1))
    pts = origin[None] + a * u[None] + b * v[None]
    pts = pts + rng.normal(0.0, jitter, (cnt, 1)) * normal[None]
    return pts


def _box(rng, lo, hi, density, jitter):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.eye(3)
    faces = [
        (lo, ex * d[0], ey * d[1], ez), (lo + ez * d[2], ex * d[0], ey * d[1], ez),
        (lo, ex * d[0], ez * d[2], ey), (lo + ey * d[1], ex * d[0], ez * d[2], ey),
        (lo, ey * d[1], ez * d[2], ex), (lo + ex * d[0], ey * d[1], ez * d[2], ex),
    ]
    return np.concatenate([_rect(rng, o, u, v, density, jitter, nrm) for o, u, v, nrm in faces], 0)


def make_scene(n_points, seed, voxel=0.04, density=2500.0, jitter=0.005):
    """-> coord (n,3) f32, feat (n,3) f32 rgb in [0,1], label (n,) int64.  n == n_points."""
    rng = np.random.default_rng(seed)
    scale = 1.0
    for _ in range(8):
        L, W = rng.uniform(4, 10, 2) * scale
        H = 3.0
        parts, labels = [], []

        def add(p, lab):
            parts.append(p)
            labels.append(np.full(len(p), lab, np.int64))

        ex, ey, ez = np.eye(3)
        zero = np.zeros(3)
        add(_rect(rng, zero, ex * L, ey * W, density, jitter, ez), 1)                 # floor
        add(_rect(rng, ez * H, ex * L, ey * W, density, jitter, ez), 0)               # ceiling
        add(_rect(rng, zero, ex * L, ez * H, density, jitter, ey), 2)                 # walls
        add(_rect(rng, ey * W, ex * L, ez * H, density, jitter, ey), 2)
        add(_rect(rng, zero, ey * W, ez * H, density, jitter, ex), 2)
        add(_rect(rng, ex * L, ey * W, ez * H, density, jitter, ex), 2)
        nbox = int(rng.integers(6, 13) * scale * scale)
        for bi in range(nbox):
            sz = rng.uniform([0.4, 0.4, 0.3], [2.0, 1.5, 1.8])
            lo = np.array([rng.uniform(0, max(L - sz[0], 0.1)), rng.uniform(0, max(W - sz[1], 0.1)), 0.0])
            add(_box(rng, lo, lo + sz, density, jitter), 3 + (bi % 8))                # furniture classes 3..10
        for bi in range(2):                                                            # boards on walls
            w, h = rng.uniform(1.0, 2.5), rng.uniform(0.8, 1.5)
            x0, z0 = rng.uniform(0, max(L - w, 0.1)), rng.uniform(0.8, 1.4)
            add(_rect(rng, np.array([x0, 0.02 + bi * (W - 0.04), z0]), ex * w, ez * h, density, jitter, ey), 11 + bi)
        coord = np.concatenate(parts, 0)
        label = np.concatenate(labels, 0)
        base = rng.random((NUM_CLASSES, 3))
        feat = np.clip(base[label] + rng.normal(0, 0.05, coord.shape), 0, 1)
        # reference preprocessing order (data_util.py:45-76)
        coord = coord - coord.min(0)
        key = np.floor(coord / voxel).astype(np.int64)
        key = key[:, 0] + key[:, 1] * 100003 + key[:, 2] * 100003 * 100003
        _, first = np.unique(key, return_index=True)
        first.sort()
        coord, feat, label = coord[first], feat[first], label[first]
        if coord.shape[0] >= n_points:
            break
        scale *= 1.5
    else:
        raise RuntimeError("could not generate enough points")
    centre = coord[rng.integers(coord.shape[0])]
    crop = np.argsort(np.sum((coord - centre) ** 2, 1))[:n_points]
    coord, feat, label = coord[crop], feat[crop], label[crop]
    perm = rng.permutation(n_points)
    coord, feat, label = coord[perm], feat[perm], label[perm]
    coord = coord - coord.min(0)
    return coord.astype(np.float32), feat.astype(np.float32), label.astype(np.int64)


def make_batch(n_scenes, n_points, seed):
    """Collated batch like pytorch/util/s3dis.py:94-130: points, features, point_labels, offset (int32 cumulative ends)."""
    if np.isscalar(n_points):
        n_points = [int(n_points)] * n_scenes
    cs, fs, ls = zip(*[make_scene(n, seed * 1000 + i) for i, n in enumerate(n_points)])
    offset = np.cumsum([c.shape[0] for c in cs]).astype(np.int32)
    return {
        "points": np.concatenate(cs, 0),
        "features": np.concatenate(fs, 0),
        "point_labels": np.concatenate(ls, 0),
        "offset": offset,
    }


def uniform_cube(n_points, seed, side=None):
    """Worst-case volumetric density variant for the KNN microbench."""
    rng = np.random.default_rng(seed)
    side = side if side is not None else (n_points / 4000.0) ** (1.0 / 3.0)
    return (rng.random((n_points, 3)) * side).astype(np.float32)
