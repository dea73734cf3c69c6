"""SURVEY §8(f) row 1 — device-side voxelize / data_prepare / collate (contrastboundary_b200/dataprep.py) against
(a) the NumPy restatement oracle/dataprep.py in its deterministic mode: bit-exact rows, and (b) the golden vectors made
by the reference's OWN functions (tests/golden/dataprep_ref.npz): voxel keys, occupied voxels, counts, voxel order."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from oracle import dataprep as O  # noqa: E402

pytestmark = pytest.mark.gpu
VS, VMAX = cases.DATAPREP_VOXEL, cases.DATAPREP_VOXEL_MAX


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_voxelize_matches_reference(golden_dir, dt):
    from contrastboundary_b200 import dataprep as P
    g = np.load(os.path.join(golden_dir, "dataprep_ref.npz"))
    coord, _, _ = cases.raw_cloud(dt)
    c0 = coord - coord.min(0)
    idx_sort, count = P.voxelize(c0, VS, mode=1)
    idx_sort, count = _np(idx_sort), _np(count)
    assert np.array_equal(count, g[f"{dt}/count"])                       # points per voxel, in the reference's voxel order
    keys = g[f"{dt}/keys"]
    ks = keys[idx_sort]
    assert np.all(ks[1:] >= ks[:-1])                                     # sorted by the reference's FNV key
    assert np.array_equal(np.sort(idx_sort), np.arange(len(coord)))      # a permutation
    same = ks[1:] == ks[:-1]
    assert np.all(idx_sort[1:][same] > idx_sort[:-1][same])              # stable inside a voxel
    o_sort, o_count = O.voxelize(c0, VS, mode=1)
    assert np.array_equal(idx_sort, o_sort) and np.array_equal(count, o_count)
    uniq = _np(P.voxelize(c0, VS, mode=0))                               # one point per voxel, ascending key
    assert np.array_equal(keys[uniq], g[f"{dt}/unique_keys"])
    uniq_r = _np(P.voxelize(c0, VS, mode=0, seed=5))
    assert np.array_equal(keys[uniq_r], g[f"{dt}/unique_keys"]) and not np.array_equal(uniq, uniq_r)


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("vmax", [None, VMAX])
def test_data_prepare_bit_exact_vs_restatement(golden_dir, dt, vmax):
    from contrastboundary_b200 import dataprep as P
    coord, feat, label = cases.raw_cloud(dt)
    c, f, l, xyz = P.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=vmax)
    oc, of, ol, oi = O.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=vmax)
    assert c.dtype == torch.float32 and f.dtype == torch.float32 and l.dtype == torch.int64
    assert np.array_equal(_np(c).view(np.uint32), oc.view(np.uint32))
    assert np.array_equal(_np(f).view(np.uint32), of.view(np.uint32))
    assert np.array_equal(_np(l), ol)
    # against the reference's own output (same np.random-free parts): same number of points, same bounding box origin,
    # and — without crop — the same voxel per row (the reference keeps a random point of the voxel, we keep the first)
    g = np.load(os.path.join(golden_dir, "dataprep_ref.npz"))
    tag = "val_nocrop" if vmax is None else "val_crop"
    rc = g[f"{dt}/{tag}/coord"]
    assert rc.shape == tuple(c.shape)
    assert np.array_equal(_np(c).min(0), rc.min(0))
    if vmax is None:
        assert np.abs(_np(c) - rc).max() < 2 * VS                        # row v of both outputs lies in voxel v


def test_data_prepare_given_centre_and_hashed_pick():
    from contrastboundary_b200 import dataprep as P
    coord, feat, label = cases.raw_cloud("f32")
    c, f, l, _ = P.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=VMAX, centre=123)
    oc, of, ol, _ = O.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=VMAX, centre=123)
    assert np.array_equal(_np(c).view(np.uint32), oc.view(np.uint32)) and np.array_equal(_np(l), ol)
    # hashed per-voxel pick + hashed centre + shuffle (train split): a valid sample of the same distribution
    kw = dict(split="train", voxel_size=VS, voxel_max=VMAX, shuffle_index=True, seed=9, pick="random")
    b = P.prepare_batch([(coord, feat, label)], **kw)
    c2, index = b["points"], _np(b["index"])
    assert c2.shape == (VMAX, 3) and float(c2.min()) == 0.0
    keys0 = O.voxel_keys(coord - coord.min(0), VS)
    assert len(np.unique(keys0[index])) == VMAX                          # one point per (original) voxel
    assert np.array_equal(_np(b["point_labels"]), label[index])
    unshuffled = P.prepare_batch([(coord, feat, label)], **dict(kw, shuffle_index=False))
    assert not np.array_equal(_np(unshuffled["index"]), index)
    assert np.array_equal(np.sort(_np(unshuffled["index"])), np.sort(index))      # the shuffle is a permutation of the crop
    b3 = P.prepare_batch([(coord, feat, label)], **kw)
    assert torch.equal(c2, b3["points"])                                  # counter-based RNG: reproducible


def test_prepare_batch_collate():
    from contrastboundary_b200 import dataprep as P
    clouds = [cases.raw_cloud("f32", n, seed) for n, seed in ((30000, 77), (9000, 78), (20000, 79))]
    b = P.prepare_batch(clouds, split="val", voxel_size=VS, voxel_max=VMAX)
    assert b["offset"].dtype == torch.int32 and b["offset"].tolist() == b["offset_host"]
    prev = 0
    for (coord, feat, label), end in zip(clouds, b["offset_host"]):
        oc, of, ol, _ = O.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=VMAX)
        assert end - prev == len(oc)
        assert np.array_equal(_np(b["points"][prev:end]).view(np.uint32), oc.view(np.uint32))
        assert np.array_equal(_np(b["features"][prev:end]).view(np.uint32), of.view(np.uint32))
        assert np.array_equal(_np(b["point_labels"][prev:end]), ol)
        prev = end
    lim = b["offset_host"][1] + 10                                        # collate_default's batch_limits (s3dis.py:112-118)
    b2 = P.prepare_batch(clouds, batch_limits=lim, split="val", voxel_size=VS, voxel_max=VMAX)
    assert b2["offset_host"] == b["offset_host"][:2] and b2["points"].shape[0] == b["offset_host"][1]


def test_prepared_batch_trains():
    """the device-prepared batch feeds the hot path directly"""
    from contrastboundary_b200 import dataprep as P, engine, model
    clouds = [cases.raw_cloud("f32", 30000, 80 + i) for i in range(2)]
    b = P.prepare_batch(clouds, split="train", voxel_size=VS, voxel_max=4096, shuffle_index=True, seed=3, pick="random")
    b["point_labels"] = b["point_labels"] % 13
    ts = engine.TrainStep(model.CBLConfig(), "cuda")
    loss = ts.step(b)
    assert torch.isfinite(loss).all()


def test_data_prepare_large_cloud_properties():
    """10^6 raw points (a full S3DIS room): size-independent properties"""
    from contrastboundary_b200 import dataprep as P
    rng = np.random.default_rng(1)
    n = 1_000_000
    coord = (rng.random((n, 3)) * np.array([8.0, 6.0, 3.0])).astype(np.float32)
    coord[:, 2] = np.round(coord[:, 2] * 2) / 2                           # layered surfaces -> many points per voxel
    feat = rng.integers(0, 256, (n, 3)).astype(np.float32)
    label = rng.integers(0, 13, n)
    vmax = 80000                                                           # the reference yaml's voxel_max
    c, f, l, _ = P.data_prepare(coord, feat, label, split="val", voxel_size=VS, voxel_max=vmax)
    assert c.shape == (vmax, 3)
    c = _np(c)
    assert np.array_equal(c.min(0), np.zeros(3, np.float32))
    k = O.voxel_keys(coord - coord.min(0), VS)
    nvox = len(np.unique(k))
    assert nvox > vmax
    # rows come out nearest-first around the centre: distances to row 0 are non-decreasing
    d2 = ((c - c[0]) ** 2).sum(1)
    assert d2[0] == 0 and np.all(np.diff(d2) >= -1e-4)
    assert float(f.max()) <= 1.0 and float(f.min()) >= 0.0
