"""SURVEY §8(f) row 2 / config 3 plumbing: the 5-level input pyramid of the TF tree (tensorflow/datasets/base.py:767-842)
for a batch of 4 spheres of ~15000 points — libcbops on the B200 vs the reference's own C++ operators (oracle/_ref/
libref_cpu.so, single thread as the TF op runs them) and the C++ restatement.  Developer / documentation tool.
    python tools/bench_pyramid.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import oracle  # noqa: E402
from contrastboundary_b200 import tf_pyramid  # noqa: E402


def cpu_pyramid(pts, lens, cfg, nb, sub):
    dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2
    for lvl in range(cfg.num_layers - 1):
        nb(pts, pts, lens, lens, r)
        pp, pl = sub(pts, lens, 2 * dl)
        nb(pp, pts, pl, lens, r)
        nb(pts, pp, lens, pl, 2 * r)
        pts, lens, dl, r = pp, pl, dl * 2, r * 2
    nb(pts, pts, lens, lens, r)


def main():
    base = cases.tf_sphere()[0]
    rng = np.random.default_rng(0)
    clouds = [base + rng.normal(0, 1e-3, base.shape).astype(np.float32) for _ in range(4)]
    pts = np.concatenate(clouds, 0).astype(np.float32)
    lens = np.array([len(c) for c in clouds], np.int32)
    cfg = tf_pyramid.PyramidConfig()
    for _ in range(2):
        tf_pyramid.segmentation_inputs_radius(pts, None, None, lens, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        tf_pyramid.segmentation_inputs_radius(pts, None, None, lens, cfg)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    res = {"workload": "5-level radius pyramid, 4 clouds x %d points" % len(base), "gpu_ms": gpu_ms}
    if oracle.have_ref_cpu():
        t0 = time.perf_counter()
        cpu_pyramid(pts, lens, cfg, oracle.ref_batch_neighbors, oracle.ref_batch_grid_subsampling)
        res["reference_cpu_ms_1thread"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    cpu_pyramid(pts, lens, cfg, oracle.batch_neighbors, oracle.batch_grid_subsampling)
    res["oracle_cpu_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
