"""SURVEY §8(f) row 3 on the GPU: test-time boundary masks and full-resolution room neighbours
(pytorch/model/basic_operators.py:69-97, pytorch/tool/test.py:250-257,392-428)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import boundary as oboundary  # noqa: E402

pytestmark = pytest.mark.gpu


def test_boundary_mask_matches_reference_goldens(golden_dir):
    from contrastboundary_b200 import boundary_eval as be
    g = np.load(os.path.join(golden_dir, "boundary_ref.npz"))
    for case in range(3):
        lab = torch.from_numpy(g[f"{case}/labels"]).cuda()
        idx = torch.from_numpy(g[f"{case}/idx"]).cuda()
        valid = lab >= 0
        b, p = be.get_boundary_mask(lab, neighbor_idx=idx, get_plain=True)
        assert np.array_equal(b.cpu().numpy(), g[f"{case}/bound"]) and np.array_equal(p.cpu().numpy(), g[f"{case}/plain"])
        bv, pv = be.get_boundary_mask(lab, neighbor_idx=idx, valid_mask=valid, get_plain=True)
        assert np.array_equal(bv.cpu().numpy(), g[f"{case}/bound_valid"]) and np.array_equal(pv.cpu().numpy(), g[f"{case}/plain_valid"])
        c = be.get_boundary_mask(lab, neighbor_idx=idx, valid_mask=valid, get_cnt=True)
        assert np.array_equal(c.cpu().numpy(), g[f"{case}/cnt"])


@pytest.mark.parametrize("kr", [16, 64])
def test_full_resolution_room_boundary_iou(kr):
    """a 200 000-point room: neighbours vs the oracle on a sample of queries, masks vs the oracle, histograms vs numpy"""
    from contrastboundary_b200 import boundary_eval as be
    from contrastboundary_b200 import synthetic
    n = 200000
    xyz, _, lab = synthetic.make_scene(n, 77)
    rng = np.random.default_rng(3)
    label = lab.astype(np.int64)
    label[rng.random(n) < 0.02] = 255                                   # ignore label, as S3DIS' unlabeled points
    pred = np.where(rng.random(n) < 0.8, label, rng.integers(0, 13, n)).astype(np.int64)
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    idx = be.room_neighbors(t(xyz), kr)
    off = np.array([n], np.int32)
    qs = np.sort(rng.choice(n, 200, replace=False))
    oi, _ = oracle.knnquery(kr, xyz, xyz[qs], off, np.array([len(qs)], np.int32))
    assert np.array_equal(idx.cpu().numpy()[qs], oi)
    ob, op = oboundary.get_boundary_mask(label, idx.cpu().numpy(), get_plain=True)
    res = be.boundary_iou(t(pred), t(label), t(xyz), krs=(kr,))[kr]
    for name, mask in (("bound", ob), ("plain", op)):
        p_, l_ = pred[mask].copy(), label[mask]
        p_[l_ == 255] = 255
        inter = p_[p_ == l_]
        ai = np.bincount(inter[inter < 13], minlength=13)[:13]
        ao = np.bincount(p_[p_ < 13], minlength=13)[:13]
        at = np.bincount(l_[l_ < 13], minlength=13)[:13]
        assert np.array_equal(res[f"{name}-i"].cpu().numpy(), ai)
        assert np.array_equal(res[f"{name}-u"].cpu().numpy(), ao + at - ai)
        assert np.array_equal(res[f"{name}-t"].cpu().numpy(), at)
    assert 0 < ob.sum() < n and 0 < op.sum() < n


def test_voting_crops_follow_reference_loop():
    """regular_crops vs the loop of pytorch/tool/test.py:199-216 restated in NumPy with the same initial potentials"""
    from contrastboundary_b200 import synthetic, voting
    coord = synthetic.make_scene(9000, 77)[0].astype(np.float64)
    pot = np.random.default_rng(3).random(len(coord)) * 1e-3
    vmax = 4000
    # reference loop (test.py:199-216)
    coord_p, idx_uni, ref = pot.copy(), np.array([]), []
    while idx_uni.size != len(coord):
        init_idx = np.argmin(coord_p)
        dist = np.sum(np.power(coord - coord[init_idx], 2), 1)
        idx_crop = np.argsort(dist, kind="stable")[:vmax]
        d = dist[idx_crop]
        coord_p[idx_crop] += np.square(1 - d / np.max(d))
        ref.append(idx_crop)
        idx_uni = np.unique(np.concatenate((idx_uni, idx_crop)))
    ours = list(voting.regular_crops(torch.from_numpy(coord).cuda(), vmax, potentials=torch.from_numpy(pot).cuda()))
    assert len(ours) == len(ref) and len(ref) >= 3
    for (idx_crop, c_sub), r in zip(ours, ref):
        assert np.array_equal(np.sort(idx_crop.cpu().numpy()), np.sort(r))      # same crop (order inside equal distances aside)
        assert float(c_sub.min()) == 0.0


def test_vote_room_accumulates_every_point():
    from contrastboundary_b200 import synthetic, voting
    rng = np.random.default_rng(5)
    base = synthetic.make_scene(6000, 78)[0]
    coord = np.concatenate([base, base + rng.normal(0, 0.004, base.shape).astype(np.float32)])   # ~2 points per voxel
    feat = rng.integers(0, 256, coord.shape).astype(np.float32)
    seen = []

    def model(inputs):                       # logits = a one-hot of the crop's scene slot: counts how often a point is predicted
        n = inputs["points"].shape[0]
        assert inputs["offset"][-1] == n and float(inputs["features"].max()) <= 1.0
        seen.append(n)
        return torch.ones(n, 13, device="cuda"), None
    cum = voting.vote_room(model, torch.from_numpy(coord).cuda(), torch.from_numpy(feat).cuda(), 13, voxel_size=0.04, voxel_max=3000,
                           generator=torch.Generator(device="cuda").manual_seed(0))
    assert cum.shape == (len(coord), 13)
    assert float(cum.min()) >= 1.0                                              # every point of the room received at least one vote
    assert len(seen) >= 1 and sum(seen) == int(cum[:, 0].sum())
