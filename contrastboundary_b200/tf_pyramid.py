"""SURVEY §8(f) row 2 — the TF tree's input pyramid on the device.

Mirror of `Dataset.tf_segmentation_inputs_radius` (tensorflow/datasets/base.py:767-842), the caller of the TF-side
operators: for every level of the 5-level network it builds the radius neighbourhoods, the grid-subsampled points of the
next level, and the pooling / upsampling neighbourhoods between the two, cropped to `neighborhood_limits`
(`big_neighborhood_filter`, base.py:756-764).  The reference runs the ~13 CPU op calls of one sample inside
`tf.data.map` worker threads (base.py:96); here they are libcbops kernels on the current CUDA stream.  Host work per
level: the replay of libstdc++'s `unordered_map` iteration order over the occupied voxel keys (O(M); it defines the
reference's output point order) and the pooled point counts that size the next level's tensors.

    inputs = segmentation_inputs_radius(points, features, labels, lens, PyramidConfig())
"""
from dataclasses import dataclass, field
from typing import List

import torch

from . import tf_ops


@dataclass
class PyramidConfig:            # tensorflow/config/s3dis.py:76-87
    num_layers: int = 5
    first_subsampling_dl: float = 0.04
    density_parameter: float = 5.0
    neighborhood_limits: List[int] = field(default_factory=lambda: [26, 31, 38, 41, 39, 29])


def stack_batch_inds(stacks_len, tight=False):
    """flat point indices of every batch element, padded with the shadow index n = sum(stacks_len)
    (`tf_stack_batch_inds_while`, base.py:694-737): (B, max_len), plus one shadow column when no row is padded."""
    lens = stacks_len.to(torch.int64)
    b = lens.shape[0]
    n = int(lens.sum())
    mx = int(lens.max()) if b > 0 else 0
    start = torch.cumsum(lens, 0) - lens
    col = torch.arange(mx, device=lens.device).unsqueeze(0)
    inds = torch.where(col < lens.unsqueeze(1), start.unsqueeze(1) + col, torch.full_like(col, n))
    if not tight and n == mx * b:
        inds = torch.cat([inds, torch.full((b, 1), n, dtype=inds.dtype, device=inds.device)], 1)
    return inds.to(torch.int32)


def segmentation_inputs_radius(stacked_points, stacked_features, point_labels, stacks_lengths, cfg: PyramidConfig = None):
    """stacked_points (N,3) f32, stacks_lengths (B,) int32 — numpy or torch.  Returns the reference's `input_dict`
    (base.py:828-840) with CUDA tensors."""
    cfg = cfg or PyramidConfig()
    pts = tf_ops._to_cuda(stacked_points, torch.float32)
    lens = tf_ops._to_cuda(stacks_lengths, torch.int32)
    dev = pts.device
    batch_inds = torch.repeat_interleave(torch.arange(lens.shape[0], device=dev), lens.long())
    # batch weight of every point (base.py:776-779)
    weights = (lens.min().float() / lens.float())[batch_inds]
    dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2.0
    nl = cfg.num_layers
    input_points, input_neighbors, input_pools = [None] * nl, [None] * nl, [None] * nl
    input_upsamples, input_lens = [None] * nl, [None] * nl
    input_upsamples[0] = torch.zeros((0, 1), dtype=torch.int32, device=dev)
    lim = cfg.neighborhood_limits
    for dt in range(nl - 1):
        neighbors = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, r, limit=lim[dt])                       # base.py:796
        pool_pts, pool_lens = tf_ops.tf_batch_subsampling(pts, lens, 2 * dl)                                # :797
        pools = tf_ops.tf_batch_neighbors(pool_pts, pts, pool_lens, lens, r, limit=lim[dt])                 # :798
        ups = tf_ops.tf_batch_neighbors(pts, pool_pts, lens, pool_lens, 2 * r, limit=lim[dt])               # :799
        input_points[dt], input_neighbors[dt], input_pools[dt] = pts, neighbors, pools
        input_upsamples[dt + 1], input_lens[dt] = ups, lens
        pts, lens = pool_pts, pool_lens
        r *= 2
        dl *= 2
    input_points[nl - 1] = pts
    input_neighbors[nl - 1] = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, r, limit=lim[nl - 1])         # :816
    input_pools[nl - 1] = torch.zeros((0, 1), dtype=torch.int32, device=dev)
    input_lens[nl - 1] = lens
    return {
        "points": tuple(input_points), "neighbors": tuple(input_neighbors), "pools": tuple(input_pools),
        "upsamples": tuple(input_upsamples), "batches_len": tuple(input_lens),
        "features": tf_ops._to_cuda(stacked_features, torch.float32) if stacked_features is not None else None,
        "batch_weights": weights,
        "in_batches": stack_batch_inds(input_lens[0]), "out_batches": stack_batch_inds(input_lens[-1]),
        "point_labels": tf_ops._to_cuda(point_labels, torch.int64) if point_labels is not None else None,
    }


# ------------------------------------------------------------------------------------------------------
# calibration of the pyramid (tensorflow/datasets/base.py:158-294) — SURVEY 8(f) row 2
# ------------------------------------------------------------------------------------------------------
def calibrate_neighbors(batches, cfg: PyramidConfig = None, keep_ratio=0.8, samples_threshold=10000):
    """`neighborhood_limits` of every layer = the `keep_ratio` percentile of the neighbourhood sizes seen in the batches
    (Dataset.calibrate_neighbors, base.py:199-294).  `batches` yields (stacked_points (N,3), stacks_lengths (B,)).
    The reference runs its whole input pipeline with the limits set to the upper bound hist_n and counts the valid entries of
    every neighbour row on the host; here only the COUNTS are computed (cb_radius_count) and histogrammed on the device.
    Like the reference, neighbourhoods of hist_n or more points fall outside the histogram (base.py:267)."""
    import math
    cfg = cfg or PyramidConfig()
    nl = cfg.num_layers
    hist_n = int(math.ceil(4.0 / 3.0 * math.pi * (cfg.density_parameter + 1) ** 3))            # base.py:207
    hists = None
    for stacked_points, stacks_lengths in batches:
        pts = tf_ops._to_cuda(stacked_points, torch.float32)
        lens = tf_ops._to_cuda(stacks_lengths, torch.int32)
        if hists is None:
            hists = torch.zeros((nl, hist_n), dtype=torch.int64, device=pts.device)
        dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2.0
        for layer in range(nl):
            counts = tf_ops.radius_counts(pts, pts, lens, lens, r).long()
            hists[layer] += torch.bincount(counts[counts < hist_n], minlength=hist_n)[:hist_n]    # :266-268
            if layer + 1 < nl:
                pts, lens = tf_ops.tf_batch_subsampling(pts, lens, 2 * dl)
                r, dl = 2 * r, 2 * dl
        if int(hists.sum(1).min()) >= samples_threshold:                                         # :249
            break
    cumsum = torch.cumsum(hists.t(), 0)                                                           # :286
    percentiles = (cumsum < keep_ratio * cumsum[hist_n - 1, :].double()).sum(0)                   # :287
    return [int(v) for v in percentiles.cpu()]


def calibrate_batches(clouds, in_radius, batch_size, rng=None, n_samples=10000):
    """`batch_limit` (the point budget of a batch) such that batches hold `batch_size` input spheres on average
    (Dataset.calibrate_batches, base.py:158-197).  clouds: list of (n,3) arrays (the sub-sampled training clouds).  The sphere
    sizes come from cb_radius_count on the device; the proportional corrector is the reference's, driven by `rng`
    (numpy Generator; the reference uses the global np.random state)."""
    import numpy as np
    rng = rng or np.random.default_rng()
    N = (n_samples // len(clouds)) + 1
    sizes = []
    for cloud in clouds:
        pts = np.asarray(cloud, dtype=np.float32)
        pick = rng.choice(pts.shape[0], size=min(N, pts.shape[0]), replace=False)
        # (the reference adds noise to a COPY and then queries the un-noised picks, base.py:170-173)
        q = np.ascontiguousarray(pts[pick])
        one = np.array([len(q)], np.int32), np.array([pts.shape[0]], np.int32)
        sizes += tf_ops.radius_counts(q, pts, one[0], one[1], in_radius).cpu().tolist()
    sizes = np.sort(np.asarray(sizes))
    lim = float(sizes[-1] * batch_size)                                                           # :179
    acc, max_b = 0, 0
    for i, s in enumerate(sizes):                                                                 # :181-187
        acc += s
        if acc > lim:
            max_b = i
            break
    estim_b = 0.0
    for i in range(10000):                                                                        # :189-196
        rand_shapes = rng.choice(sizes, size=max_b, replace=False)
        b = np.sum(np.cumsum(rand_shapes) < lim)
        estim_b += (b - estim_b) / min(i + 1, 100)
        lim += 10.0 * (batch_size - estim_b)
    return lim
