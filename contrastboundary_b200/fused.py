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
