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
