"""Generate tests/golden/tfops_ref_cpu.npz from the REFERENCE's own TF-side C++ cores
(oracle/_ref/libref_cpu.so, libref_cpy.so — compiled unmodified from /root/reference by
oracle/build_ref.sh).  Runs in the build container (no GPU needed):

    python tests/golden/make_golden_cpu.py

Inputs are regenerated from tests/golden/cases.py seeds; only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import oracle  # noqa: E402


def main():
    assert oracle.have_ref_cpu(), "run oracle/build_ref.sh first (needs /root/reference)"
    res = {}
    # config 1
    p, f, l = cases.tf_config1()
    sp, sf, sl = oracle.ref_grid_subsampling(p, f, l, 0.08)
    res["c1/sub_points"], res["c1/sub_features"], res["c1/sub_labels"] = sp, sf, sl
    lens = np.array([sp.shape[0]], np.int32)
    res["c1/neighbors"] = oracle.ref_batch_neighbors(sp, sp, lens, lens, 0.1)
    # pyramid of a config-3 sphere + a second scene in the batch
    p2 = cases.tf_config1()[0][:3000]
    pts = np.concatenate([cases.tf_sphere()[0], p2], 0)
    lens = np.array([15000, 3000], np.int32)
    dl, r = 0.04 * 2, 0.04 * 5 / 2
    for lvl in range(3):
        res[f"pyr/{lvl}/neighbors"] = oracle.ref_batch_neighbors(pts, pts, lens, lens, r)
        pool_pts, pool_lens = oracle.ref_batch_grid_subsampling(pts, lens, dl)
        res[f"pyr/{lvl}/pool_pts"], res[f"pyr/{lvl}/pool_lens"] = pool_pts, pool_lens
        res[f"pyr/{lvl}/pools"] = oracle.ref_batch_neighbors(pool_pts, pts, pool_lens, lens, r)
        res[f"pyr/{lvl}/upsamples"] = oracle.ref_batch_neighbors(pts, pool_pts, lens, pool_lens, 2 * r)
        pts, lens = pool_pts, pool_lens
        dl *= 2
        r *= 2
    out = os.path.join(ROOT, "tests", "golden", "tfops_ref_cpu.npz")
    res = {k: (v.astype(np.int16) if v.dtype == np.int32 and v.size and v.max() < 32767 else v) for k, v in res.items()}
    np.savez_compressed(out, **res)
    print("wrote", out, {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
