"""SURVEY §8(f) row 1 — the reference's host-side batch preparation on the device: `voxelize`
(pytorch/util/voxelize.py:38-56), `data_prepare` (pytorch/util/data_util.py:45-92) and the collate of
pytorch/util/s3dis.py:94-130, same names and arguments, over libcbops' cb_voxelize / cb_data_prepare.

The reference runs these in NumPy on dataloader workers (an argsort over 10^5..10^6 points per cloud, twice).  Here
one cloud is ~15 kernel launches and a whole batch is prepared without a device->host read: every cloud is appended
to the batch buffers at a row offset that lives on the device.  `prepare_batch` reads the B+1 offsets back ONCE at the
end (the network's host code needs the scene sizes), `prepare_batch_async` leaves even that to the caller.

What is bit-exact and what is defined modulo NumPy's RNG / unstable sort: csrc/dataprep.cu header."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if not t.is_cuda:
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _coord(x):
    t = _dev(x)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64 if t.dtype in (torch.int64,) else torch.float32)
    return t


def _ws(n, device):
    lib = L.lib()
    lib.cb_data_prepare_workspace_bytes.restype = C.c_size_t
    return L.workspace(lib.cb_data_prepare_workspace_bytes(C.c_int(n)), device, "dataprep")


def voxelize(coord, voxel_size=0.05, hash_type="fnv", mode=0, seed=None):
    """voxelize.py:38-56.  mode 1 (val): (idx_sort, count) — point indices sorted by voxel key, points per voxel.
    mode 0 (train): idx_unique, one point per voxel in ascending key order; the reference draws the point with
    np.random, here it is the first point of the voxel in input order (seed None) or a hashed draw (seed int)."""
    if hash_type != "fnv":
        raise NotImplementedError("voxelize: only the FNV64-1A hash (the default, the one data_prepare uses) is built")
    c = _coord(coord)
    n = c.shape[0]
    dev = c.device
    idx_sort = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    nvox = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _ws(n, dev)
    rc = L.lib().cb_voxelize(L.ptr(c), C.c_int(int(c.dtype == torch.float64)), C.c_int(n), C.c_double(float(voxel_size)),
                             L.ptr(idx_sort), L.ptr(count), L.ptr(nvox), None, L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_voxelize")
    nv = int(nvox.item())
    count = count[:nv]
    if mode != 0:
        return idx_sort.long(), count.long()
    start = torch.cumsum(count.long(), 0) - count.long()
    if seed is None:
        pick = torch.zeros_like(start)
    else:
        g = torch.Generator(device=dev).manual_seed(int(seed))
        pick = torch.randint(0, int(count.max().item()), (nv,), generator=g, device=dev) % count.long()   # voxelize.py:50
    return idx_sort.long()[start + pick]


class _Batch:
    """device buffers of one batch under construction"""

    def __init__(self, capacity, fdim, b, device):
        self.points = torch.empty((capacity, 3), dtype=torch.float32, device=device)
        self.features = torch.empty((capacity, fdim), dtype=torch.float32, device=device)
        self.labels = torch.empty((capacity,), dtype=torch.int64, device=device)
        self.index = torch.empty((capacity,), dtype=torch.int32, device=device)
        self.offsets = torch.zeros(b + 1, dtype=torch.int32, device=device)
        self.capacity = capacity


def _prepare_into(batch, i, coord, feat, label, split, voxel_size, voxel_max, shuffle_index, seed, pick, feat_div, centre):
    c = _coord(coord)
    f = _dev(feat, c.dtype)
    lab = _dev(label, torch.int64)
    n = c.shape[0]
    ws = _ws(n, c.device)
    if centre is None:
        centre = -2 if "train" in split else -1                        # data_util.py:59-63
    rc = L.lib().cb_data_prepare(
        L.ptr(c), L.ptr(f), C.c_int(int(c.dtype == torch.float64)), C.c_int(f.shape[1]), L.ptr(lab), C.c_int(n),
        C.c_double(float(voxel_size or 0.0)), C.c_int(int(voxel_max or 0)), C.c_int(1 if pick == "random" else 0),
        C.c_int(int(centre)), C.c_int(1 if shuffle_index else 0), C.c_ulonglong(int(seed) & (2 ** 64 - 1)), C.c_float(feat_div),
        C.c_void_p(batch.offsets.data_ptr() + 4 * i), C.c_int(batch.capacity), L.ptr(batch.points), L.ptr(batch.features),
        L.ptr(batch.labels), L.ptr(batch.index), L.ptr(ws), C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_data_prepare")


def _cap(n, voxel_max):
    return min(n, int(voxel_max)) if voxel_max else n


def data_prepare(coord, feat, label, split="train", voxel_size=0.04, voxel_max=None, transform=None, shuffle_index=False,
                 origin="min", seed=0, pick="first", centre=None):
    """data_util.py:45-92 for one cloud -> (coord, feat, label, xyz) CUDA tensors (float32, float32 in [0,1], int64,
    float32).  Extra keywords: seed / pick ('first' | 'random') / centre (index into the voxelised cloud) pin down what
    the reference takes from np.random."""
    if transform:
        coord, feat, label = transform(coord, feat, label)
    if origin != "min":
        raise NotImplementedError("data_prepare: only origin='min' (the reference's default and only use) is built")
    n = int(coord.shape[0])
    b = _Batch(_cap(n, voxel_max), int(feat.shape[1]), 1, _coord(coord).device)
    _prepare_into(b, 0, coord, feat, label, split, voxel_size, voxel_max, shuffle_index, seed, pick, 255.0, centre)
    m = int(b.offsets[1].item())
    return b.points[:m], b.features[:m], b.labels[:m], b.points[:m]


def prepare_batch_async(clouds, split="train", voxel_size=0.04, voxel_max=None, shuffle_index=False, seed=0, pick="first",
                        feat_div=255.0):
    """clouds: list of (coord (n,3), feat (n,d), label (n)) — NumPy arrays or tensors, float32 or float64 coordinates.
    Enqueues the preparation of the whole batch on the current stream and returns the padded device buffers
    {points, features, point_labels, index, offsets (B+1, device int32)} without any host synchronisation."""
    cap = sum(_cap(int(c[0].shape[0]), voxel_max) for c in clouds)
    dev = _coord(clouds[0][0]).device
    batch = _Batch(cap, int(clouds[0][1].shape[1]), len(clouds), dev)
    for i, (coord, feat, label) in enumerate(clouds):
        _prepare_into(batch, i, coord, feat, label, split, voxel_size, voxel_max, shuffle_index, seed * 1000003 + i, pick, feat_div, None)
    return {"points": batch.points, "features": batch.features, "point_labels": batch.labels, "index": batch.index,
            "offsets": batch.offsets}


def prepare_batch(clouds, batch_limits=None, **kw):
    """prepare_batch_async + the collate of s3dis.py:94-130: ONE device->host read (the B+1 offsets) for the whole batch.
    batch_limits: drop the trailing clouds once the running point count exceeds it (collate_default :112-118).
    -> dict(points, features, point_labels, offset (int32 cumulative ends, device), offset_host (python list))."""
    r = prepare_batch_async(clouds, **kw)
    offs = r["offsets"].tolist()                         # the single synchronisation of the batch
    ends = offs[1:]
    if batch_limits is not None:
        keep = 0
        for e in ends:
            if e > batch_limits:
                break
            keep += 1
        ends = ends[:max(keep, 0)]
    m = ends[-1] if ends else 0
    return {"points": r["points"][:m], "features": r["features"][:m], "point_labels": r["point_labels"][:m],
            "offset": r["offsets"][1:1 + len(ends)], "offset_host": ends, "index": r["index"][:m]}
