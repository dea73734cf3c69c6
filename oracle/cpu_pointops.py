"""oracle/cpu_pointops.py — TEST INFRASTRUCTURE.

The reference's operator API (pytorch/lib/pointops/functions/pointops.py) on CPU tensors, backed
by the C restatement in pointops_oracle.c.  Lets (a) the REAL reference model code run on CPU in
the build container (tests/golden/make_golden_model.py injects this module as
`lib.pointops.functions.pointops`), and (b) the restated model (oracle/ref_model.py) run on the GPU
box's host cores for bench.py --impl reference / cpu_baseline.
Only the functions the reference model calls are provided (pointops.py:27,45,79,164)."""
import numpy as np
import torch

import oracle as O


def furthestsampling(xyz, offset, new_offset):
    """pointops.py:10-27"""
    idx = O.furthestsampling(xyz.detach().numpy(), offset.numpy(), new_offset.numpy())
    return torch.from_numpy(idx)


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """pointops.py:30-45 — returns (idx int32, sqrt(dist2))"""
    if new_xyz is None:
        new_xyz = xyz
    idx, d2 = O.knnquery(int(nsample), xyz.detach().numpy(), new_xyz.detach().numpy(),
                         np.asarray(offset, dtype=np.int32), np.asarray(new_offset, dtype=np.int32))
    return torch.from_numpy(idx), torch.sqrt(torch.from_numpy(d2))


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """pointops.py:79-100"""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    n, m, c = xyz.shape[0], new_xyz.shape[0], feat.shape[1]
    grouped_xyz = xyz[idx.view(-1).long(), :].view(m, nsample, 3)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(1)
    grouped_feat = feat[idx.view(-1).long(), :].view(m, nsample, c)
    if use_xyz:
        return torch.cat((grouped_xyz, grouped_feat), -1)
    return grouped_feat


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.py:164-178"""
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=1, keepdim=True)
    weight = dist_recip / norm
    new_feat = torch.zeros(new_xyz.shape[0], feat.shape[1])
    for i in range(k):
        new_feat = new_feat + feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return new_feat
