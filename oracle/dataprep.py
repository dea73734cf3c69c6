"""oracle/dataprep.py — TEST INFRASTRUCTURE.  NumPy restatement of the reference's batch preparation
(pytorch/util/voxelize.py:4-16,38-56; pytorch/util/data_util.py:45-92), the checker of
contrastboundary_b200/dataprep.py.

Two modes:
  * rng="numpy"  — the reference's own calls in the reference's own order (np.argsort on the keys, np.random.randint
    for the per-voxel pick, np.random.randint for the centre, np.random.shuffle): bit-identical to the reference under
    the same np.random.seed.  Pinned by tests/golden/dataprep_ref.npz (made from the REAL functions imported from
    /root/reference by tests/golden/make_golden_dataprep.py), checked in tests/test_oracle_cpu.py.
  * rng=None     — the deterministic variant the CUDA path implements: stable sort, first point of every voxel (or the
    picks passed in), centre given / middle, crop ties by index, no shuffle.  Same arithmetic, line by line.
Arithmetic runs in the dtype of `coord` (the reference: NumPy 1.x value-based casting keeps `coord / np.array(voxel_size)`
in coord's dtype; under NumPy >= 2 pass voxel_size as a scalar of that dtype to get the same).
"""
import numpy as np


def fnv_hash_vec(arr):
    """voxelize.py:4-16 (FNV64-1A over the columns)"""
    arr = arr.astype(np.uint64)
    h = np.full(arr.shape[0], 14695981039346656037, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(arr.shape[1]):
            h = h * np.uint64(1099511628211)
            h = np.bitwise_xor(h, arr[:, j])
    return h


def voxel_keys(coord, voxel_size):
    """voxelize.py:39,43"""
    vs = coord.dtype.type(voxel_size)
    return fnv_hash_vec(np.floor(coord / vs))


def voxelize(coord, voxel_size=0.05, mode=0, rng=None, picks=None):
    key = voxel_keys(coord, voxel_size)
    idx_sort = np.argsort(key) if rng == "numpy" else np.argsort(key, kind="stable")      # voxelize.py:45
    key_sort = key[idx_sort]
    _, count = np.unique(key_sort, return_counts=True)
    if mode != 0:
        return idx_sort, count
    start = np.cumsum(np.insert(count, 0, 0)[0:-1])
    if rng == "numpy":
        r = np.random.randint(0, count.max(), count.size) % count                         # voxelize.py:50
    elif picks is not None:
        r = np.asarray(picks) % count
    else:
        r = 0
    return idx_sort[start + r]


def data_prepare(coord, feat, label, split="train", voxel_size=0.04, voxel_max=None, shuffle_index=False, rng=None, centre=None,
                 picks=None):
    """data_util.py:45-92 (origin='min', no transform) -> coord f32, feat f32 (/255), label i64, index (input row of every
    output row)"""
    coord = coord.copy()
    index = np.arange(coord.shape[0])
    if voxel_size:
        coord -= np.min(coord, 0)                                                         # :54-55
        uniq = voxelize(coord, voxel_size, 0, rng, picks)
        coord, feat, label, index = coord[uniq], feat[uniq], label[uniq], index[uniq]
    n = label.shape[0]
    if rng == "numpy" and "train" in split and voxel_max and n > voxel_max:
        init_idx = np.random.randint(n)                                                   # :59-60
    elif centre is not None:
        init_idx = min(int(centre), n - 1)
    else:
        init_idx = n // 2                                                                 # :63
    coord_init = coord[init_idx]
    if voxel_max and n > voxel_max:
        d2 = np.sum(np.square(coord - coord_init), 1)                                     # :67
        order = np.argsort(d2) if rng == "numpy" else np.argsort(d2, kind="stable")
        crop = order[:voxel_max]
        coord, feat, label, index = coord[crop], feat[crop], label[crop], index[crop]
    if shuffle_index and rng == "numpy":
        shuf = np.arange(coord.shape[0])
        np.random.shuffle(shuf)                                                           # :69-71
        coord, feat, label, index = coord[shuf], feat[shuf], label[shuf], index[shuf]
    coord = coord - np.min(coord, 0)                                                      # :75-76
    return coord.astype(np.float32), (feat.astype(np.float32) / np.float32(255.0)), label.astype(np.int64), index
