"""The two independently written restatements of the TF tree's a13 / a14 operators (oracle/tf_model.py in torch,
oracle/tf_convnet_np.py in NumPy float64) must agree with each other — neither is pinned by TensorFlow itself
("parity unpinned")."""
import numpy as np
import torch

from oracle import tf_convnet_np as R
from oracle import tf_model as T


def _case(seed, n=300, n0=500, k=12, c=16):
    rng = np.random.default_rng(seed)
    q, s = rng.random((n, 3)), rng.random((n0, 3))
    idx = rng.integers(0, n0 + 1, (n, k))
    idx[:, -3:] = np.where(rng.random((n, 3)) < 0.5, n0, idx[:, -3:])       # shadow entries
    return q, s, idx, rng.standard_normal((n0, c)), rng.standard_normal((c, 3)), rng.standard_normal(c)


def test_adaptive_weight_restatements_agree():
    q, s, idx, f, w, b = _case(0)
    P = {"x.fc_1.weight": w, "x.fc_1.bias": b, "x.pool_bn.weight": np.ones(16), "x.pool_bn.bias": np.zeros(16)}
    a = R.adaptive_weight(P, "x", q, s, idx, f, 0.3, 1e-6)
    tq, ts_, tf_, tw, tb = (torch.from_numpy(v) for v in (q, s, f, w, b))
    agg = T.adaptive_weight(tq, ts_, torch.from_numpy(idx), tf_, tw, tb, 0.3)
    y = torch.relu((agg - agg.mean(0)) / torch.sqrt(agg.var(0, unbiased=False) + 1e-6)).numpy()
    assert np.abs(a - y).max() < 1e-6          # tf_model.py counts the neighbours in float32 (cnt + 1e-5), as TF does


def test_contrast_loss_restatements_agree():
    rng = np.random.default_rng(1)
    n, k, d = 400, 20, 72
    feat = rng.standard_normal((n, d))
    idx = rng.integers(0, n, (n, k))
    idx[:, 0] = np.arange(n)
    idx[:, -4:] = np.where(rng.random((n, 4)) < 0.5, n, idx[:, -4:])
    cls = rng.integers(0, 4, n)
    a = R.contrast_loss(feat, idx, cls, None, 0.1)
    b = float(T.contrast_loss(torch.from_numpy(feat), torch.from_numpy(idx), torch.from_numpy(cls), 1.0, 0.1))
    assert abs(a - b) < 1e-12 * max(1.0, abs(b)) and a > 0


def test_hard_labels_and_nearest_index_brute_force():
    rng = np.random.default_rng(2)
    p0 = rng.random((600, 3)).astype(np.float32)
    p2 = p0[::40] + 0.01
    inp = {"points": [p0, p0[::4], p2], "batches_len": [np.array([600]), np.array([150]), np.array([15])],
           "point_labels": rng.integers(0, 5, 600), "upsamples": [None, np.zeros((600, 1), np.int64)], "pools": [rng.integers(0, 601, (150, 6))]}
    up, cls = R.head_geometry(inp, [0.2, 0.3], 5)
    assert up[2].shape == (600,) and cls[2].shape == (15,) and (up[2] <= 15).all()
    d = ((p0[:, None] - p2[None]) ** 2).sum(-1)
    assert np.array_equal(np.where(d.min(1) < 0.09, d.argmin(1), 15), up[2])
