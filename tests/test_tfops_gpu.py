"""GPU parity of the TF-side operators (grid subsampling, radius neighbours, knn_batch) against the
golden vectors produced by the REFERENCE's own C++ cores (tests/golden/tfops_ref_cpu.npz) and the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402
from test_oracle_cpu import rows_equal_mod_ties  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tf_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "tfops_ref_cpu.npz"))


def test_config1_grid_subsampling_and_neighbors(tf_golden):
    from contrastboundary_b200 import tf_ops
    p, f, l = cases.tf_config1()
    sp, sf, sl = tf_ops.grid_subsampling(p, f, l, 0.08)
    assert np.array_equal(sp, tf_golden["c1/sub_points"])            # bit-exact barycentres in the reference's order
    assert np.array_equal(sf, tf_golden["c1/sub_features"])
    assert np.array_equal(sl, tf_golden["c1/sub_labels"])
    lens = np.array([sp.shape[0]], np.int32)
    nb = tf_ops.tf_batch_neighbors(sp, sp, lens, lens, 0.1).cpu().numpy()
    assert rows_equal_mod_ties(nb, tf_golden["c1/neighbors"].astype(np.int32), sp, sp)


def test_pyramid_levels(tf_golden):
    from contrastboundary_b200 import tf_ops
    p2 = cases.tf_config1()[0][:3000]
    pts = np.concatenate([cases.tf_sphere()[0], p2], 0)
    lens = np.array([15000, 3000], np.int32)
    dl, r = 0.08, 0.1
    for lvl in range(3):
        g = lambda k: tf_golden[f"pyr/{lvl}/{k}"]  # noqa: E731
        nb = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, r).cpu().numpy()
        assert rows_equal_mod_ties(nb, g("neighbors").astype(np.int32), pts, pts), f"neighbors lvl {lvl}"
        pool_pts, pool_lens = tf_ops.tf_batch_subsampling(pts, lens, dl)
        pool_pts, pool_lens = pool_pts.cpu().numpy(), pool_lens.cpu().numpy()
        assert np.array_equal(pool_pts, g("pool_pts")) and np.array_equal(pool_lens, g("pool_lens"))
        pools = tf_ops.tf_batch_neighbors(pool_pts, pts, pool_lens, lens, r).cpu().numpy()
        assert rows_equal_mod_ties(pools, g("pools").astype(np.int32), pool_pts, pts), f"pools lvl {lvl}"
        ups = tf_ops.tf_batch_neighbors(pts, pool_pts, lens, pool_lens, 2 * r).cpu().numpy()
        assert rows_equal_mod_ties(ups, g("upsamples").astype(np.int32), pts, pool_pts), f"upsamples lvl {lvl}"
        # the caller crops to a neighbourhood limit (datasets/base.py:762): fused crop gives the same columns
        lim = 20
        cropped = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, r, limit=lim).cpu().numpy()
        assert rows_equal_mod_ties(cropped, nb[:, :lim], pts, pts)
        pts, lens, dl, r = pool_pts, pool_lens, dl * 2, r * 2


def test_random_clouds_vs_oracle():
    from contrastboundary_b200 import tf_ops
    rng = np.random.default_rng(5)
    for trial in range(3):
        lens = rng.integers(1, 3000, 4).astype(np.int32)
        lens[trial % 4] = 1                                       # a single-point scene
        pts = (rng.random((int(lens.sum()), 3)) * rng.uniform(0.5, 3.0)).astype(np.float32)
        dl = float(rng.uniform(0.05, 0.3))
        a = tf_ops.tf_batch_subsampling(pts, lens, dl)
        b = oracle.batch_grid_subsampling(pts, lens, dl)
        assert np.array_equal(a[0].cpu().numpy(), b[0]) and np.array_equal(a[1].cpu().numpy(), b[1])
        nb = tf_ops.tf_batch_neighbors(b[0], pts, b[1], lens, 1.5 * dl).cpu().numpy()
        assert rows_equal_mod_ties(nb, oracle.batch_neighbors(b[0], pts, b[1], lens, 1.5 * dl), b[0], pts)
        feats = rng.random((int(lens[1]), 5)).astype(np.float32)
        labs = rng.integers(0, 4, (int(lens[1]), 2)).astype(np.int32)
        seg = pts[lens[0]:lens[0] + lens[1]]
        ga = tf_ops.grid_subsampling(seg, feats, labs, dl)
        gb = oracle.grid_subsampling(seg, feats, labs, dl)
        assert all(np.array_equal(x, y) for x, y in zip(ga, gb))


def test_knn_batch():
    from contrastboundary_b200 import tf_ops
    rng = np.random.default_rng(6)
    sup = rng.random((3, 500, 3)).astype(np.float32)
    qry = rng.random((3, 120, 3)).astype(np.float32)
    idx = tf_ops.tf_knn_search(qry, sup, 5).cpu().numpy()
    ref = oracle.knn_batch(sup, qry, 5)
    assert np.array_equal(idx, ref)


def _pyramid_level(seed=3):
    from contrastboundary_b200 import tf_ops
    p2 = cases.tf_config1()[0][:3000]
    pts = np.concatenate([cases.tf_sphere()[0][:6000], p2], 0)
    lens = np.array([6000, 3000], np.int32)
    nb = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, 0.1, limit=26)
    return torch.from_numpy(pts).cuda(), nb


def test_adaptive_weight_matches_restatement():
    """a13 forward + backward against oracle/tf_model.py, a line-by-line restatement that is itself pinned by the executed
    reference source (tests/test_convnet_cpu.py); the forward is also compared with that golden directly, below"""
    from contrastboundary_b200 import tf_model
    from oracle import tf_model as otf
    pts, nb = _pyramid_level()
    torch.manual_seed(0)
    for c in (72, 144, 32):
        feat = torch.randn(pts.shape[0], c, device="cuda")
        w = torch.randn(c, 3, device="cuda") * 0.5
        b = torch.randn(c, device="cuda") * 0.1
        g = torch.randn(pts.shape[0], c, device="cuda")
        res = []
        for fn in (otf.adaptive_weight, tf_model.adaptive_weight):
            f2, w2, b2 = (x.clone().requires_grad_(True) for x in (feat, w, b))
            out = fn(pts, pts, nb, f2, w2, b2, 0.1)
            out.backward(g)
            res.append((out.detach(), f2.grad, w2.grad, b2.grad))
        for a, r in zip(res[1], res[0]):
            assert float((a - r).abs().max()) <= 2e-5 * float(r.abs().max()) + 1e-7


def test_adaptive_weight_matches_executed_reference_source():
    """a13 against tests/golden/tf_ops_ref.npz: vectors produced by EXECUTING the reference's own AdaptiveWeight source
    (tensorflow/models/local_aggregation_operators.py:316-500, adapt.yaml) on a NumPy stand-in for the TF-1 API
    (tests/golden/make_golden_tf_ops.py) — the CUDA kernel reproduces the aggregation that enters pool_bn, and the module-level
    relu(pool_bn(.)) reproduces the operator's output."""
    import os
    from contrastboundary_b200 import tf_model
    g = np.load(os.path.join(ROOT, "tests", "golden", "tf_ops_ref.npz"))
    for name in ("self", "pool"):
        k = lambda s: g[f"aw/{name}/{s}"]
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
        q, sup, feat = t32(k("query")), t32(k("support")), t32(k("features"))
        nb = torch.from_numpy(k("neighbors").astype(np.int32)).cuda()
        w, b = t32(k("fc_weight").T), t32(k("fc_bias"))                          # TF kernel (3, c) -> (c, 3)
        agg = tf_model.adaptive_weight(q, sup, nb, feat, w, b, float(k("radius")))
        ref = torch.from_numpy(k("aggregated")).cuda()
        assert float((agg.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
        y = torch.relu(torch.nn.functional.batch_norm(agg, None, None, t32(k("bn_gamma")), t32(k("bn_beta")), True, 0.0, 1e-3))
        out = torch.from_numpy(k("output")).cuda()
        assert float((y.double() - out).abs().max()) <= 1e-4 * max(1.0, float(out.abs().max()))


def test_tf_contrast_loss_matches_restatement():
    """a14 forward + backward against oracle/tf_model.py, a restatement that is itself pinned by the executed reference source
    (tests/test_convnet_cpu.py); the forward is also compared with that golden directly, below"""
    from contrastboundary_b200 import tf_model
    from oracle import tf_model as otf
    pts, nb = _pyramid_level()
    n = pts.shape[0]
    torch.manual_seed(1)
    labels = (pts[:, 0] * 3).long() % 5                      # spatially coherent labels -> boundaries exist
    for d in (72, 32):
        feat = torch.randn(n, d, device="cuda")
        res = []
        for fn in (otf.contrast_loss, tf_model.tf_contrast_loss):
            f2 = feat.clone().requires_grad_(True)
            loss = fn(f2, nb, labels, 1.0, 0.1)
            loss.backward()
            res.append((float(loss), f2.grad))
        assert abs(res[1][0] - res[0][0]) <= 1e-5 * abs(res[0][0])
        assert float((res[1][1] - res[0][1]).abs().max()) <= 1e-4 * float(res[0][1].abs().max())


def test_tf_contrast_head_matches_executed_reference_source():
    """a14 against tests/golden/tf_ops_ref.npz: the reference's contrast head (heads/head.py:462-807) EXECUTED on the NumPy TF
    stand-in at every stage of a 5-level pyramid (tests/golden/make_golden_tf_ops.py): the CUDA label votes reproduce its hard
    sub-scene labels exactly, the CUDA soft-NN loss reproduces its loss at every stage (incl. the stages whose loss is 0)."""
    from contrastboundary_b200 import convnet, tf_model
    g = np.load(os.path.join(ROOT, "tests", "golden", "tf_ops_ref.npz"))
    pts = [torch.from_numpy(g[f"net/points/{l}"].astype(np.float32)).cuda() for l in range(5)]
    lens = [torch.from_numpy(g[f"cbl/batches_len/{l}"].astype(np.int32)).cuda() for l in range(5)]
    nbs = [torch.from_numpy(g[f"net/neighbors/{l}"].astype(np.int32)).cuda() for l in range(5)]
    pools0 = torch.from_numpy(g["net/pools/0"].astype(np.int32)).cuda()
    labels = torch.from_numpy(g["cbl/point_labels"].astype(np.int64)).cuda()
    r_sample = [float(r) for r in g["cbl/r_sample"]]
    nonzero = 0
    for i in range(5):
        ref_lab = g[f"cbl/scene_label/{i}"]
        if i == 0:
            cls = labels.to(torch.int32)
        elif i == 1:
            cls = convnet.label_vote_idx(pools0, labels, 13, pts[0].shape[0])
        else:
            cls = convnet.label_vote_radius(pts[i], pts[0], lens[i], lens[0], r_sample[i - 1], labels, 13)
        assert np.array_equal(cls.cpu().numpy().astype(np.int64), ref_lab), f"hard sub-scene labels, stage {i}"
        feat = torch.from_numpy(g[f"cbl/latent/{i}"].astype(np.float32)).cuda()
        ref = float(g[f"cbl/loss/{i}"])
        loss = float(tf_model.tf_contrast_loss(feat, nbs[i], cls, 1.0, 0.1))
        assert abs(loss - ref) <= 1e-4 * max(abs(ref), 1e-3), (i, loss, ref)
        nonzero += ref > 0
    assert nonzero >= 3


def test_segmentation_inputs_radius_pyramid():
    """the whole 5-level input pyramid (datasets/base.py:767-842) built on the device vs the same
    composition of the CPU oracle operators, level by level"""
    from contrastboundary_b200 import tf_pyramid
    p2 = cases.tf_config1()[0][:3000]
    pts = np.concatenate([cases.tf_sphere()[0], p2], 0)
    lens = np.array([15000, 3000], np.int32)
    labels = (np.arange(len(pts)) % 13).astype(np.int64)
    cfg = tf_pyramid.PyramidConfig()
    out = tf_pyramid.segmentation_inputs_radius(pts, pts.copy(), labels, lens, cfg)
    dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2
    cp, cl = pts, lens
    for lvl in range(cfg.num_layers):
        lim = cfg.neighborhood_limits[lvl]
        assert np.array_equal(out["points"][lvl].cpu().numpy(), cp), f"points lvl {lvl}"
        assert np.array_equal(out["batches_len"][lvl].cpu().numpy(), cl)
        nb = oracle.batch_neighbors(cp, cp, cl, cl, r)[:, :lim]
        assert rows_equal_mod_ties(_pad_cols(out["neighbors"][lvl].cpu().numpy(), nb.shape[1]), nb, cp, cp), f"neighbors lvl {lvl}"
        if lvl == cfg.num_layers - 1:
            assert out["pools"][lvl].shape == (0, 1)
            break
        pp, pl = oracle.batch_grid_subsampling(cp, cl, 2 * dl)
        pools = oracle.batch_neighbors(pp, cp, pl, cl, r)[:, :lim]
        ups = oracle.batch_neighbors(cp, pp, cl, pl, 2 * r)[:, :lim]
        assert rows_equal_mod_ties(_pad_cols(out["pools"][lvl].cpu().numpy(), pools.shape[1]), pools, pp, cp), f"pools lvl {lvl}"
        assert rows_equal_mod_ties(_pad_cols(out["upsamples"][lvl + 1].cpu().numpy(), ups.shape[1]), ups, cp, pp), f"ups lvl {lvl}"
        cp, cl, dl, r = pp, pl, dl * 2, r * 2
    assert out["upsamples"][0].shape == (0, 1)
    # batch weights and the stacked batch index matrices (base.py:776-779, 694-737)
    w = out["batch_weights"].cpu().numpy()
    assert np.allclose(w[:15000], 3000 / 15000) and np.allclose(w[15000:], 1.0)
    ib = out["in_batches"].cpu().numpy()
    assert ib.shape == (2, 15000) and ib[0, -1] == 14999 and ib[1, 2999] == 17999 and ib[1, 3000] == 18000
    eq = tf_pyramid.stack_batch_inds(torch.tensor([3, 3], dtype=torch.int32, device="cuda")).cpu().numpy()
    assert np.array_equal(eq, np.array([[0, 1, 2, 6], [3, 4, 5, 6]], np.int32))      # extra shadow column when no row is padded


def _pad_cols(a, width):
    """the fused crop returns exactly `limit` columns; the oracle's matrix is narrower when no row has that many
    neighbours — the extra columns must then be pure shadow padding"""
    if a.shape[1] <= width:
        return a
    assert (a[:, width:] == a.max()).all()
    return a[:, :width]


def test_calibrate_neighbors_and_batches():
    """tf_pyramid.calibrate_neighbors / calibrate_batches (tensorflow/datasets/base.py:158-294) against the same procedure
    evaluated with the CPU restatement of the reference's radius search"""
    import math
    import oracle
    from contrastboundary_b200 import synthetic, tf_pyramid
    cfg = tf_pyramid.PyramidConfig()
    batches = []
    for s in range(2):
        clouds = [synthetic.make_scene(5000, 900 + 10 * s + i)[0] for i in range(2)]
        batches.append((np.concatenate(clouds).astype(np.float32), np.array([5000, 5000], np.int32)))
    limits = tf_pyramid.calibrate_neighbors(batches, cfg, keep_ratio=0.8, samples_threshold=10 ** 9)
    hist_n = int(math.ceil(4 / 3 * math.pi * (cfg.density_parameter + 1) ** 3))
    hists = np.zeros((cfg.num_layers, hist_n), np.int64)
    for pts, lens in batches:
        dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2
        for layer in range(cfg.num_layers):
            nb = oracle.batch_neighbors(pts, pts, lens, lens, r)
            counts = np.sum(nb < nb.shape[0], axis=1)                       # base.py:266
            hists[layer] += np.bincount(counts, minlength=hist_n)[:hist_n]
            if layer + 1 < cfg.num_layers:
                pts, lens = oracle.batch_grid_subsampling(pts, lens, 2 * dl)
                r, dl = 2 * r, 2 * dl
    cumsum = np.cumsum(hists.T, axis=0)
    ref = np.sum(cumsum < (0.8 * cumsum[hist_n - 1, :]), axis=0)            # base.py:286-287
    assert limits == [int(v) for v in ref], (limits, ref)
    assert all(5 < v < 80 for v in limits)
    clouds = [synthetic.make_scene(20000, 950 + i)[0] for i in range(2)]
    lim = tf_pyramid.calibrate_batches(clouds, in_radius=1.0, batch_size=4, rng=np.random.default_rng(0), n_samples=400)
    # the proportional corrector converges to a budget of ~batch_size average spheres
    sizes = []
    for c in clouds:
        d = np.random.default_rng(1).choice(len(c), 100, replace=False)
        sizes += [int((((c - c[i]) ** 2).sum(1) < 1.0).sum()) for i in d]
    assert 2.5 * np.mean(sizes) < lim < 6.0 * np.mean(sizes), (lim, np.mean(sizes))


def test_batch_neighbors_width_contract_and_wide_rows():
    """tf_batch_neighbors: `limit` pads to exactly `limit` columns (no host read), exact_width=True reproduces the reference's
    `neighbors[:, :limit]` shape min(max_count, limit) (datasets/base.py:762); rows wider than 256 (dense clouds during
    neighbourhood calibration) take the exact scan"""
    import oracle
    from contrastboundary_b200 import synthetic, tf_ops
    pts = synthetic.make_scene(3000, 41)[0]
    lens = np.array([3000], np.int32)
    ref = oracle.batch_neighbors(pts, pts, lens, lens, 0.1)
    mx = ref.shape[1]
    padded = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, 0.1, limit=mx + 7).cpu().numpy()
    assert padded.shape == (3000, mx + 7) and (padded[:, mx:] == 3000).all()
    exact = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, 0.1, limit=mx + 7, exact_width=True).cpu().numpy()
    assert exact.shape == ref.shape
    assert np.array_equal(np.sort(exact, 1), np.sort(ref, 1))               # same sets (order inside equal distances aside)
    cropped = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, 0.1, limit=5, exact_width=True).cpu().numpy()
    assert cropped.shape == (3000, 5)
    wide_ref = oracle.batch_neighbors(pts, pts, lens, lens, 0.4)
    assert wide_ref.shape[1] > 256
    wide = tf_ops.tf_batch_neighbors(pts, pts, lens, lens, 0.4).cpu().numpy()
    assert wide.shape == wide_ref.shape
    assert np.array_equal((wide < 3000).sum(1), (wide_ref < 3000).sum(1))
    rows = np.random.default_rng(0).choice(3000, 50, replace=False)
    for r in rows:
        assert np.array_equal(np.sort(wide[r]), np.sort(wide_ref[r]))
