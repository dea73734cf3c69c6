"""oracle/gpu_pointops.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's operator API (pytorch/lib/pointops/functions/pointops.py) on CUDA tensors, backed by
the reference's OWN kernels compiled unmodified for sm_100a (oracle/_ref/pointops_cuda.so, built by
oracle/build_ref.sh).  Together with oracle/ref_model.py this is the "stock pointops CUDA build"
baseline B1 of BASELINE.md §3 (the >=10x denominator of the north star): /root/reference does not
exist on the GPU box, so the reference's python wrappers are restated here around the reference's
compiled extension.  Only the functions the reference model calls are provided
(pointops.py:27,45,79,164); none of this repo's kernels is on this path.
"""
import torch

import oracle as O

_mod = None


def _ext():
    global _mod
    if _mod is None:
        _mod = O.ref_pointops_cuda()
    return _mod


def furthestsampling(xyz, offset, new_offset):
    """pointops.py:10-27 (host loop over the offsets included: it syncs, as the reference does)"""
    oh = offset.tolist()
    n_max, prev = oh[0], oh[0]
    for e in oh[1:]:
        n_max = max(n_max, e - prev)
        prev = e
    b = offset.shape[0]
    m = int(new_offset[b - 1].item())
    idx = torch.zeros(m, dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    _ext().furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx)
    return idx


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """pointops.py:30-45 — (idx int32, sqrt(dist2))"""
    if new_xyz is None:
        new_xyz = xyz
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    _ext().knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, d2)
    return idx, torch.sqrt(d2)


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """pointops.py:79-100"""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    m, c = new_xyz.shape[0], feat.shape[1]
    flat = idx.view(-1).long()
    grouped_xyz = xyz[flat, :].view(m, nsample, 3) - new_xyz.unsqueeze(1)
    grouped_feat = feat[flat, :].view(m, nsample, c)
    if use_xyz:
        return torch.cat((grouped_xyz, grouped_feat), -1)
    return grouped_feat


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.py:164-178 (python loop over k, as the reference)"""
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    recip = 1.0 / (dist + 1e-8)
    weight = recip / torch.sum(recip, dim=1, keepdim=True)
    out = torch.zeros((new_xyz.shape[0], feat.shape[1]), dtype=feat.dtype, device=feat.device)
    for i in range(k):
        out = out + feat[idx[:, i].long(), :] * weight[:, i].unsqueeze(-1)
    return out
