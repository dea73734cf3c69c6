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
