"""Fused operators that have no single counterpart in the reference API (they replace short
sequences of reference calls).  Each cites the reference sequence it replaces."""
import ctypes as C

import torch

from . import _lib as L


def knn_gather(nsample, xyz, new_xyz, feat, offset, new_offset):
    """idx, dist2, grouped = fused [knnquery -> feat[idx]]  (reference pointops.py:88-94).

    idx (m,K) int32, dist2 (m,K) f32 SQUARED distances, grouped (m,K,c) f32.  One kernel searches
    and streams the K feature rows through shared memory with TMA bulk copies (cb_knn_gather)."""
    if new_xyz is None:
        new_xyz = xyz
    L.require_cuda(xyz, new_xyz, feat, offset, new_offset)
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    n, m, b, c = xyz.shape[0], new_xyz.shape[0], offset.shape[0], feat.shape[1]
    dev = xyz.device
    idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
    dist2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
    grouped = torch.empty((m, nsample, c), dtype=torch.float32, device=dev)
    ws = L.workspace(L.lib().cb_knn_workspace_bytes(n, m, b), dev, "knn")
    rc = L.lib().cb_knn_gather(C.c_int(m), C.c_int(int(nsample)), C.c_int(c), L.ptr(xyz), C.c_int(n), L.ptr(new_xyz),
                               L.ptr(feat), L.ptr(offset), L.ptr(new_offset), C.c_int(b), L.ptr(idx), L.ptr(dist2),
                               L.ptr(grouped), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_knn_gather")
    return idx, dist2, grouped


def grid_build(xyz, offset, nsample_hint=16, m_max=None):
    """Build the uniform search grid of a support set once (cb_grid_build).  Returns the workspace tensor
    to pass to knn_gather_grid / knn_query_grid (valid until the next call that reuses it)."""
    n, b = xyz.shape[0], offset.shape[0]
    m_max = n if m_max is None else m_max
    nbytes = L.lib().cb_knn_workspace_bytes(n, m_max, b)
    ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=xyz.device)
    rc = L.lib().cb_grid_build(L.ptr(xyz), C.c_int(n), L.ptr(offset), C.c_int(b), C.c_int(int(nsample_hint)), L.ptr(ws),
                               C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_grid_build")
    return ws


def knn_gather_grid(grid, nsample, xyz, new_xyz, feat, offset, new_offset, out=None):
    """fused KNN+gather on a prebuilt grid (the timed north-star kernel)"""
    if new_xyz is None:
        new_xyz = xyz
    n, m, b, c = xyz.shape[0], new_xyz.shape[0], offset.shape[0], feat.shape[1]
    dev = xyz.device
    if out is None:
        out = (torch.empty((m, nsample), dtype=torch.int32, device=dev), torch.empty((m, nsample), dtype=torch.float32, device=dev),
               torch.empty((m, nsample, c), dtype=torch.float32, device=dev))
    idx, dist2, grouped = out
    rc = L.lib().cb_knn_gather_grid(C.c_int(m), C.c_int(int(nsample)), C.c_int(c), L.ptr(xyz), C.c_int(n), L.ptr(new_xyz),
                                    L.ptr(feat), L.ptr(offset), L.ptr(new_offset), C.c_int(b), L.ptr(idx), L.ptr(dist2),
                                    L.ptr(grouped), L.ptr(grid), C.c_size_t(grid.numel()), L.stream())
    L.check(rc, "cb_knn_gather_grid")
    return idx, dist2, grouped
