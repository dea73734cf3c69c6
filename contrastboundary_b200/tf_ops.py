"""Drop-in surface for the reference's TF-side operator API (`tensorflow/ops/tf_ops.py`, class
TF_OPS): the same callables on torch CUDA tensors / numpy arrays (TensorFlow is not part of this
stack), backed by libcbops.so.

    from contrastboundary_b200 import tf_ops
    sub_pts, sub_lens = tf_ops.tf_batch_subsampling(points, batches_len, sampleDl)      # tf_ops.py:158
    neighbors        = tf_ops.tf_batch_neighbors(queries, supports, q_lens, s_lens, r)  # tf_ops.py:165
    pts[, feats][, labels] = tf_ops.grid_subsampling(points, features, labels, sampleDl)  # tf_ops.py:82
    idx              = tf_ops.tf_knn_search(query_pts, support_pts, k)                  # tf_ops.py:117

Semantics follow the reference's CPU kernels: per-scene LENGTHS (not cumulative offsets), shadow
index = number of supports, row width = batch-wide max count, output voxel order = the reference's
(std::unordered_map iteration order).  No CPU compute path: tensors must be / are moved to CUDA.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _dev():
    if not torch.cuda.is_available():
        raise L.CbopsError("contrastboundary_b200.tf_ops needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(a, dtype):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device=_dev(), dtype=dtype).contiguous()


def _offsets(lens):
    return torch.cumsum(lens.to(torch.int64), 0).to(torch.int32).contiguous()


def _subsample(points, lens, dl, features=None):
    """-> (sub_points (M,3), sub_features or None, sub_lens (b) int32, point_cell (n), perm, order info)"""
    dev = points.device
    n, b = points.shape[0], lens.shape[0]
    fdim = 0 if features is None else features.shape[1]
    off = _offsets(lens)
    lib = L.lib()
    lib.cb_grid_subsample_workspace_bytes.restype = C.c_size_t
    ws = L.workspace(lib.cb_grid_subsample_workspace_bytes(C.c_int(n), C.c_int(b), C.c_int(fdim)), dev, "gs")
    cxyz = torch.empty((n, 3), dtype=torch.float32, device=dev)
    cfeat = torch.empty((n, max(fdim, 1)), dtype=torch.float32, device=dev) if fdim else None
    ckey = torch.empty(n, dtype=torch.int64, device=dev)
    cfirst = torch.empty(n, dtype=torch.int32, device=dev)
    pcell = torch.empty(n, dtype=torch.int32, device=dev)
    ncell = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.cb_grid_subsample_cells(L.ptr(points), C.c_int(n), L.ptr(off), C.c_int(b), C.c_float(float(dl)), L.ptr(features),
                                     C.c_int(fdim), L.ptr(cxyz), L.ptr(cfeat), L.ptr(ckey), L.ptr(cfirst), L.ptr(pcell),
                                     L.ptr(ncell), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_grid_subsample_cells")
    m = int(ncell.item())                                    # the op's output size is data dependent (as in the reference)
    keys_h = ckey[:m].cpu().numpy().astype(np.uint64)
    first_h = np.ascontiguousarray(cfirst[:m].cpu().numpy())
    perm_h = np.zeros(max(m, 1), np.int32)
    counts_h = np.zeros(b, np.int32)
    rc = lib.cb_unordered_map_order(keys_h.ctypes.data_as(C.c_void_p), first_h.ctypes.data_as(C.c_void_p), C.c_int(m), C.c_int(b),
                                    perm_h.ctypes.data_as(C.c_void_p), counts_h.ctypes.data_as(C.c_void_p))
    L.check(rc, "cb_unordered_map_order")
    perm = torch.from_numpy(perm_h[:m]).to(dev)
    out_xyz = torch.empty((m, 3), dtype=torch.float32, device=dev)
    out_feat = torch.empty((m, fdim), dtype=torch.float32, device=dev) if fdim else None
    rc = lib.cb_grid_subsample_permute(C.c_int(m), C.c_int(fdim), L.ptr(perm), L.ptr(cxyz), L.ptr(cfeat), L.ptr(out_xyz),
                                       L.ptr(out_feat), L.stream())
    L.check(rc, "cb_grid_subsample_permute")
    return out_xyz, out_feat, torch.from_numpy(counts_h).to(dev), pcell, perm, m


def tf_batch_subsampling(points, batches_len, sampleDl):
    """grid subsampling for stacked clouds [BxN, 3] -> (sub_points, sub_batches_len)   (tf_ops.py:158-161)"""
    points = _to_cuda(points, torch.float32)
    lens = _to_cuda(batches_len, torch.int32)
    out_xyz, _, counts, _, _, _ = _subsample(points, lens, sampleDl)
    return out_xyz, counts


batch_grid_subsampling = tf_batch_subsampling


def grid_subsampling(points, features=None, labels=None, sampleDl=0.1, verbose=0):
    """CPP-wrapper flavour (tf_ops.py:82-102): one cloud, numpy in / numpy out; features -> voxel mean,
    labels -> per-column majority vote."""
    as_numpy = isinstance(points, np.ndarray)
    p = _to_cuda(points, torch.float32)
    n = p.shape[0]
    f = None if features is None else _to_cuda(features, torch.float32).reshape(n, -1)
    lens = torch.tensor([n], dtype=torch.int32, device=p.device)
    out_xyz, out_feat, _, pcell, perm, m = _subsample(p, lens, sampleDl, f)
    outs = [out_xyz]
    if f is not None:
        outs.append(out_feat)
    if labels is not None:
        lab = _to_cuda(labels, torch.int64).reshape(n, -1)
        lab_h = None
        cols = []
        for col in range(lab.shape[1]):
            l = lab[:, col]
            lo = int(l.min().item())
            ncls = int(l.max().item()) - lo + 1
            cnt = torch.zeros((m, ncls), dtype=torch.int32, device=p.device)
            cnt.index_put_((pcell.long(), l - lo), torch.ones(n, dtype=torch.int32, device=p.device), accumulate=True)
            mx = cnt.max(1, keepdim=True)[0]
            win = cnt.argmax(1) + lo
            tied = ((cnt == mx).sum(1) > 1).nonzero().flatten()
            if tied.numel():                                   # the reference's tie rule is its hash-map iteration order
                if lab_h is None:
                    lab_h, pc_h = lab.cpu().numpy(), pcell.cpu().numpy()
                win_h = win.cpu().numpy()
                for c in tied.cpu().numpy():
                    seq = np.ascontiguousarray(lab_h[pc_h == c, col].astype(np.int32))
                    win_h[c] = L.lib().cb_label_vote_host(seq.ctypes.data_as(C.c_void_p), C.c_int(len(seq)))
                win = torch.from_numpy(win_h).to(p.device)
            cols.append(win[perm.long()].to(torch.int32))
        out_lab = torch.stack(cols, 1)
        outs.append(out_lab[:, 0] if np.ndim(labels) == 1 or (hasattr(labels, "dim") and labels.dim() == 1) else out_lab)
    if as_numpy:
        outs = [o.cpu().numpy() for o in outs]
    return outs[0] if len(outs) == 1 else tuple(outs)


def tf_batch_neighbors(queries, supports, q_batches, s_batches, radius, limit=None, exact_width=False):
    """radius neighbours for stacked clouds -> int32 (Nq, max_count) padded with Ns   (tf_ops.py:165-168).
    `limit` (optional) = neighborhood_limits[layer]: return only the nearest `limit` columns, which is what
    the caller keeps anyway (datasets/base.py:762).  Width contract: with `limit` the result is EXACTLY `limit`
    columns wide (columns past the batch-wide max count hold the shadow index Ns) and no device->host read
    happens; `exact_width=True` reproduces the reference's `neighbors[:, :limit]` shape, min(max_count, limit),
    at the price of one host sync."""
    q = _to_cuda(queries, torch.float32)
    s = _to_cuda(supports, torch.float32)
    same = (queries is supports)
    if same:
        q = s
    ql, sl = _to_cuda(q_batches, torch.int32), _to_cuda(s_batches, torch.int32)
    qo, so = _offsets(ql), _offsets(sl)
    nq, ns, b = q.shape[0], s.shape[0], ql.shape[0]
    lib = L.lib()
    ws = L.workspace(lib.cb_knn_workspace_bytes(ns, nq, b), q.device, "knn")
    counts = torch.empty(max(nq, 1), dtype=torch.int32, device=q.device)
    mx = torch.zeros(1, dtype=torch.int32, device=q.device)
    rc = lib.cb_radius_count(C.c_int(nq), L.ptr(q), C.c_int(ns), L.ptr(s), L.ptr(qo), L.ptr(so), C.c_int(b), C.c_float(float(radius)),
                             L.ptr(counts), L.ptr(mx), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_radius_count")
    if limit is None:
        width = int(mx.item())
    else:
        width = min(int(limit), int(mx.item())) if exact_width else int(limit)
    out = torch.empty((nq, width), dtype=torch.int32, device=q.device)
    if width > 0 and nq > 0:
        rc = lib.cb_radius_fill(C.c_int(nq), C.c_int(width), L.ptr(q), C.c_int(ns), L.ptr(s), L.ptr(qo), L.ptr(so), C.c_int(b),
                                C.c_float(float(radius)), L.ptr(out), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
        L.check(rc, "cb_radius_fill")
    return out


batch_neighbors = tf_batch_neighbors


def radius_counts(queries, supports, q_batches, s_batches, radius):
    """number of supports within `radius` of every query (strict, nanoflann arithmetic) -> int32 (Nq): the row lengths of
    tf_batch_neighbors without materialising the rows (what the reference's calibration needs of them:
    `np.sum(neighb_mat < neighb_mat.shape[0], axis=1)`, datasets/base.py:266, and `len(neighb)` of :173-175)"""
    q = _to_cuda(queries, torch.float32)
    s = q if queries is supports else _to_cuda(supports, torch.float32)
    ql, sl = _to_cuda(q_batches, torch.int32), _to_cuda(s_batches, torch.int32)
    qo, so = _offsets(ql), _offsets(sl)
    nq, ns, b = q.shape[0], s.shape[0], ql.shape[0]
    lib = L.lib()
    ws = L.workspace(lib.cb_knn_workspace_bytes(ns, nq, b), q.device, "knn")
    counts = torch.empty(max(nq, 1), dtype=torch.int32, device=q.device)
    mx = torch.zeros(1, dtype=torch.int32, device=q.device)
    rc = lib.cb_radius_count(C.c_int(nq), L.ptr(q), C.c_int(ns), L.ptr(s), L.ptr(qo), L.ptr(so), C.c_int(b), C.c_float(float(radius)),
                             L.ptr(counts), L.ptr(mx), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_radius_count")
    return counts[:nq]


def tf_knn_search(query_pts, support_pts, k):
    """[B,M,3] queries, [B,N,3] supports -> int32 [B,M,k] indices local to each batch element (tf_ops.py:117-129)"""
    q = _to_cuda(query_pts, torch.float32)
    s = _to_cuda(support_pts, torch.float32)
    B, M, _ = q.shape
    N = s.shape[1]
    ql = torch.full((B,), M, dtype=torch.int32, device=q.device)
    sl = torch.full((B,), N, dtype=torch.int32, device=q.device)
    qo, so = _offsets(ql), _offsets(sl)
    lib = L.lib()
    qf, sf = q.reshape(-1, 3).contiguous(), s.reshape(-1, 3).contiguous()
    ws = L.workspace(lib.cb_knn_workspace_bytes(B * N, B * M, B), q.device, "knn")
    counts = torch.empty(B * M, dtype=torch.int32, device=q.device)
    mx = torch.zeros(1, dtype=torch.int32, device=q.device)
    rc = lib.cb_radius_count(C.c_int(B * M), L.ptr(qf), C.c_int(B * N), L.ptr(sf), L.ptr(qo), L.ptr(so), C.c_int(B), C.c_float(0.0),
                             L.ptr(counts), L.ptr(mx), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_radius_count")
    out = torch.empty((B * M, k), dtype=torch.int32, device=q.device)
    rc = lib.cb_radius_fill(C.c_int(B * M), C.c_int(int(k)), L.ptr(qf), C.c_int(B * N), L.ptr(sf), L.ptr(qo), L.ptr(so), C.c_int(B),
                            C.c_float(1e18), L.ptr(out), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_radius_fill")
    base = (torch.arange(B, device=q.device, dtype=torch.int32) * N).repeat_interleave(M).unsqueeze(1)
    return (out - base).view(B, M, k)


_FUNCS = {
    "grid_subsampling": grid_subsampling, "tf_batch_subsampling": tf_batch_subsampling,
    "tf_batch_neighbors": tf_batch_neighbors, "tf_knn_search": tf_knn_search,
}


def get_tf_func(key):
    """same accessor as the reference's `ops.get_tf_func(key)` (tensorflow/ops/tf_ops.py:26-73)"""
    return _FUNCS[key]
