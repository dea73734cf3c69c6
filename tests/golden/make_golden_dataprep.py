"""Generate tests/golden/dataprep_ref.npz from the REAL reference functions (pytorch/util/voxelize.py,
pytorch/util/data_util.py, imported unmodified from /root/reference; its `SharedArray` import — dataset shared-memory
I/O, unused here — is stubbed).

    python tests/golden/make_golden_dataprep.py

Cases: cases.raw_cloud(dtype) for float32 and float64 coordinates, with and without the voxel_max crop.  Under NumPy >= 2
`coord / np.array(voxel_size)` would promote float32 clouds to float64 (NEP 50); the reference's pinned NumPy 1.x keeps
coord's dtype, so voxel_size is passed as a scalar of coord's dtype — the same arithmetic the reference ran.
np.random.seed(cases.DATAPREP_SEED) precedes every call."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402


def import_reference():
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
    sys.path.insert(0, "/root/reference/pytorch")
    from util import data_util, voxelize
    return voxelize, data_util


def main():
    vox, du = import_reference()
    res = {}
    for dt in ("f32", "f64"):
        coord, feat, label = cases.raw_cloud(dt)
        vs = coord.dtype.type(cases.DATAPREP_VOXEL)
        c0 = coord - coord.min(0)
        key = vox.fnv_hash_vec(np.floor(c0 / np.array(vs)))
        idx_sort, count = vox.voxelize(c0, vs, mode=1)
        res[f"{dt}/keys"] = key
        res[f"{dt}/unique_keys"] = key[idx_sort][np.cumsum(np.insert(count, 0, 0)[:-1])]
        res[f"{dt}/count"] = count
        for tag, vmax, split, shuf in (("val_nocrop", None, "val", False), ("val_crop", cases.DATAPREP_VOXEL_MAX, "val", False),
                                       ("train_crop", cases.DATAPREP_VOXEL_MAX, "train", True)):
            np.random.seed(cases.DATAPREP_SEED)
            c, f, l, xyz = du.data_prepare(coord.copy(), feat.copy(), label.copy(), split=split, voxel_size=vs, voxel_max=vmax,
                                           shuffle_index=shuf)
            res[f"{dt}/{tag}/coord"] = c.numpy()
            res[f"{dt}/{tag}/feat"] = f.numpy()
            res[f"{dt}/{tag}/label"] = l.numpy()
    out = os.path.join(ROOT, "tests", "golden", "dataprep_ref.npz")
    np.savez_compressed(out, **res)
    print("wrote", out, {k: v.shape for k, v in res.items() if k.endswith("coord") or k.endswith("count")})


if __name__ == "__main__":
    main()
