"""The two independently written restatements of the TF tree's a13 / a14 operators (oracle/tf_model.py in torch,
oracle/tf_convnet_np.py in NumPy float64) must agree with each other, and — for AdaptiveWeight, the resnet backbone, the
segmentation head and the contrast head — with vectors produced by executing the reference's own TF source on a NumPy
stand-in of the TF-1 API (tests/golden/make_golden_tf_ops.py)."""
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


def test_adaptive_weight_restatements_match_executed_reference_source():
    """tests/golden/tf_ops_ref.npz comes from EXECUTING the reference's own AdaptiveWeight source
    (tensorflow/models/local_aggregation_operators.py:316-500, adapt.yaml) on a NumPy stand-in for the TF-1 API
    (tests/golden/make_golden_tf_ops.py): both restatements reproduce the aggregation (what enters pool_bn) and the
    operator's output relu(pool_bn(.)) — row a13 is pinned by reference code that ran, not only by reading it."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    for name in ("self", "pool"):
        k = lambda s: g[f"aw/{name}/{s}"]
        q, sup, idx, feat = k("query").astype(np.float64), k("support").astype(np.float64), k("neighbors").astype(np.int64), k("features")
        w, b, radius = k("fc_weight").T.copy(), k("fc_bias"), float(k("radius"))      # TF kernel (3, c) -> (c, 3)
        assert (idx == len(sup)).any() and (idx < len(sup)).all(1).any()              # shadow entries and full rows both occur
        agg = T.adaptive_weight(torch.from_numpy(q), torch.from_numpy(sup), torch.from_numpy(idx), torch.from_numpy(feat),
                                torch.from_numpy(w), torch.from_numpy(b), radius).numpy()
        assert np.abs(agg - k("aggregated")).max() < 1e-6 * np.abs(k("aggregated")).max()     # tf_model.py counts in float32
        c = feat.shape[1]
        P = {"x.fc_1.weight": w, "x.fc_1.bias": b, "x.pool_bn.weight": k("bn_gamma"), "x.pool_bn.bias": k("bn_beta")}
        out = R.adaptive_weight(P, "x", q, sup, idx, feat, radius, 1e-3)
        assert out.shape == (len(q), c)
        assert np.abs(out - k("output")).max() < 1e-10 * max(1.0, np.abs(k("output")).max())


def test_backbone_and_seg_head_restatement_match_executed_reference_source():
    """The reference's resnet_backbone (backbone/resnet.py:307-420: input conv, simple block, strided / plain bottlenecks with
    AdaptiveWeight, ind_max_pool shortcuts) and resnet_scene_segmentation_head (heads/seg_head.py:31-110), EXECUTED on the NumPy
    TF stand-in over a 5-level pyramid built by the CPU oracle of the reference's C++ operators: the float64 restatement that
    the CUDA network is tested against (oracle/tf_convnet_np.py, tests/test_convnet_gpu.py) reproduces every stage feature.
    The TF variable names go through the product's checkpoint-name converter (convnet.tf_variable_to_state_dict)."""
    import os
    import types
    from contrastboundary_b200 import convnet
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    fdim, dl, density, ratio, depth = g["net/config"]
    cfg = types.SimpleNamespace(bn_eps=1e-3, first_features_dim=int(fdim), first_subsampling_dl=float(dl), density_parameter=float(density),
                                num_layers=5, depth=int(depth))
    P, used = {}, 0
    for key in g.files:
        if not key.startswith("net/var/"):
            continue
        m = convnet.tf_variable_to_state_dict(key[len("net/var/"):])
        assert m is not None, key
        P[m[0]] = g[key].T.copy() if m[1] else g[key]
        used += 1
    assert used == 200 and "resnet_backbone.res.1.strided_bottleneck.conv2.fc_1.weight" in P
    ref_keys = set(convnet.ConvNetSeg(convnet.ConvNetConfig()).state_dict().keys())
    assert {k for k in P if "res.0.bottleneck0" in k or "up_conv" in k or "res1_" in k} <= ref_keys      # same names as the product's modules
    inp = {k: [g[f"net/{k}/{l}"].astype(np.float64 if k == "points" else np.int64) for l in range(5)] for k in ("points", "neighbors", "pools", "upsamples")}
    inp["features"] = g["net/features"]
    feats = R.backbone(P, inp, cfg)
    f_out = R.seg_head_features(P, inp, feats, cfg)
    for l in range(5):
        ref = g[f"net/F/{l}"]
        assert feats[l].shape == ref.shape and np.abs(feats[l] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), l
    for l in range(4):
        ref = g[f"net/F_up/{l}"]
        assert np.abs(f_out[l] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), l


def test_multiscale_head_restatement_matches_executed_reference_source():
    """The reference's multiscale_head (heads/head.py:338-460) EXECUTED on the NumPy TF stand-in with the reference's own
    head-config object (config/head.py multiscale_1 = '||Ua-concat-latent') on the executed backbone's features: per-stage latent
    MLP, nearest upsampling to the input points through the cross-stage radius search (some input points find no coarse point
    in range and gather the zero row), concat, linear classifier, mean sparse softmax cross-entropy.  The restatement
    reproduces latents, logits and loss; the variable names go through the product's checkpoint-name converter."""
    import os
    import types
    from contrastboundary_b200 import convnet
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    P = {}
    for key in g.files:
        if key.startswith("multi/var/"):
            m = convnet.tf_variable_to_state_dict(key[len("multi/var/"):])
            assert m is not None, key
            P[m[0]] = g[key].T.copy() if m[1] else g[key]
    ref_keys = set(convnet.ConvNetSeg(convnet.ConvNetConfig()).state_dict().keys())
    assert len(P) == 5 * 5 + 2 and set(P) <= ref_keys
    inp = {k: [g[f"net/{k}/{l}"] for l in range(5)] for k in ("points", "upsamples", "pools")}
    inp["batches_len"] = [g[f"cbl/batches_len/{l}"] for l in range(5)]
    inp["point_labels"] = g["cbl/point_labels"]
    up_idx0, _ = R.head_geometry(inp, [float(r) for r in g["multi/r_sample"]], 13)
    assert sum(int((up_idx0[i] == len(inp["points"][i])).sum()) for i in range(2, 5)) > 50          # the zero-row path is exercised
    f_out = [g[f"net/F_up/{l}"] for l in range(4)] + [g["net/F/4"]]
    cfg = types.SimpleNamespace(bn_eps=1e-6, num_layers=5)
    logits, latents, xen = R.multiscale_head(P, f_out, up_idx0, inp["point_labels"], cfg)
    for i in range(5):
        ref = g[f"multi/latent/{i}"]
        assert np.abs(latents[i] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), i
    assert np.abs(logits - g["multi/logits"]).max() < 1e-9 * np.abs(g["multi/logits"]).max()
    assert abs(xen - float(g["multi/loss"])) < 1e-12 * float(g["multi/loss"]) + 1e-12


def test_whole_network_restatement_matches_executed_reference_model():
    """The reference's OWN builder — models/build_models.py SceneSegModel with its own config object (config/s3dis.py Conv
    '|multi-Ua-concat-latent|contrast-Ua-softnn-latent-label-l2-w.1' + config/s3dis/adapt.yaml; only first_features_dim is
    reduced) — EXECUTED end to end on the NumPy TF stand-in: backbone, segmentation head, build_head (its load_config +
    apply_head_ops for both heads), build_loss.  The float64 restatement that the CUDA network is tested against
    (oracle/tf_convnet_np.forward) reproduces the logits, every entry of the loss dictionary, the L2 term and the total."""
    import os
    import types
    from contrastboundary_b200 import convnet
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    fdim, dl, density, ratio, depth, eps, wd, ncls = g["full/config"]
    assert (ratio, depth, eps, wd, ncls) == (2, 1, 1e-6, 1e-3, 13)                     # adapt.yaml / s3dis.py values reached the model
    P, l2 = {}, 0.0
    decay_rule = lambda k: k.endswith("weights.weight") or k.endswith("linear.weight")     # ConvNetSeg.decay_parameters()
    for key in g.files:
        if key.startswith("full/var/"):
            m = convnet.tf_variable_to_state_dict(key[len("full/var/"):])
            assert m is not None, key
            P[m[0]] = g[key].T.copy() if m[1] else g[key]
            if decay_rule(m[0]):
                l2 += wd * 0.5 * float((g[key] ** 2).sum())
    cfg = types.SimpleNamespace(bn_eps=float(eps), first_features_dim=int(fdim), first_subsampling_dl=float(dl), density_parameter=float(density),
                                num_layers=5, depth=int(depth), r_sample=[float(r) for r in g["full/r_sample"]], num_classes=int(ncls),
                                contrast=True, contrast_temperature=None, contrast_weight=0.1)
    inp = {k: [g[f"net/{k}/{l}"].astype(np.float64 if k == "points" else np.int64) for l in range(5)] for k in ("points", "neighbors", "pools", "upsamples")}
    inp["batches_len"] = [g[f"cbl/batches_len/{l}"] for l in range(5)]
    inp["point_labels"], inp["features"] = g["cbl/point_labels"], g["net/features"]
    logits, losses, latents, _ = R.forward(P, inp, cfg)
    ref = g["full/logits"]
    assert np.abs(logits - ref).max() < 1e-8 * np.abs(ref).max()
    names = ["seg"] + [f"softnn-up{i}" for i in range(5)]
    for v, n in zip(losses, names):
        r = float(g["full/loss/" + n])
        assert abs(v - r) < 1e-9 * max(1.0, abs(r)), (n, v, r)
    assert sum(float(g["full/loss/" + n]) > 0 for n in names) >= 5
    assert abs(l2 - float(g["full/loss/l2_loss"])) < 1e-12 * float(g["full/loss/l2_loss"])   # the optimiser's weight-decay set, as an L2 loss
    assert abs(losses.sum() + l2 - float(g["full/loss/loss"])) < 1e-9 * float(g["full/loss/loss"])


def test_weight_decay_set_matches_executed_reference_source():
    """Which variables the reference puts an L2 weight loss on (recorded while its source ran with weight_decay > 0:
    basic_operators.py:126-129,371-379) == the parameters ConvNetSeg.decay_parameters() hands to the optimiser's weight decay
    (l2_loss = sum(w^2) / 2, so d/dw = weight_decay * w: the same update)."""
    import os
    from contrastboundary_b200 import convnet
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    l2 = set(str(n) for n in g["l2_loss_variables"])
    names = [k[len("net/var/"):] for k in g.files if k.startswith("net/var/")] + [k[len("multi/var/"):] for k in g.files if k.startswith("multi/var/")]
    assert len(l2) == 34 and l2 <= set(names)
    model = convnet.ConvNetSeg(convnet.ConvNetConfig())
    decay, rest = model.decay_parameters()
    decay_ids = {id(p) for p in decay}
    param_names = {n: id(p) for n, p in model.named_parameters()}
    checked = 0
    for tf_name in names:
        key, _ = convnet.tf_variable_to_state_dict(tf_name)
        if key not in param_names:                      # running statistics are buffers, not parameters
            assert key.endswith(("running_mean", "running_var")), key
            continue
        assert (param_names[key] in decay_ids) == (tf_name in l2), (tf_name, key)
        checked += 1
    assert checked > 120


def test_contrast_head_restatements_match_executed_reference_source():
    """a14: the reference's contrast head (heads/head.py:462-807 — sample_labels 'label', collect_labels / get_scene_label 'max',
    solve_samples_mask, calc_dist 'l2', soft-NN) EXECUTED on the NumPy TF stand-in at every stage of a 5-level pyramid; its
    cross-stage radius search is served by the CPU oracle of the reference's C++ neighbour operator.  The restatements
    reproduce the hard sub-scene labels exactly and the loss of every stage (including the stages where no point has both a
    positive and a negative neighbour: loss 0)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    inp = {k: [g[f"net/{k}/{l}"] for l in range(5)] for k in ("points", "neighbors", "pools", "upsamples")}
    inp["batches_len"] = [g[f"cbl/batches_len/{l}"] for l in range(5)]
    inp["point_labels"] = g["cbl/point_labels"]
    _, cls = R.head_geometry(inp, [float(r) for r in g["cbl/r_sample"]], 13)
    nonzero = 0
    for i in range(5):
        ref_lab = g[f"cbl/scene_label/{i}"]
        assert np.array_equal(np.asarray(cls[i]).astype(np.int64), ref_lab), f"hard sub-scene labels, stage {i}"
        feat, nb, ref = g[f"cbl/latent/{i}"], inp["neighbors"][i].astype(np.int64), float(g[f"cbl/loss/{i}"])
        a = R.contrast_loss(feat, nb, np.asarray(cls[i]), None, 0.1)
        b = float(T.contrast_loss(torch.from_numpy(feat), torch.from_numpy(nb), torch.from_numpy(np.asarray(cls[i]).astype(np.int64)), 1.0, 0.1))
        assert abs(a - ref) < 1e-10 * max(1.0, abs(ref)) and abs(b - ref) < 1e-10 * max(1.0, abs(ref)), (i, a, b, ref)
        nonzero += ref > 0
    assert nonzero >= 3


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


# ---- the torch module composition of contrastboundary_b200/convnet.py against the NumPy restatement, in float64 on the CPU -------
def _cpu_pyramid(pts, feats, labels, lens, cfg):
    """tf_pyramid.segmentation_inputs_radius with the CPU restatements of the reference operators (oracle/tfops_oracle.cpp)"""
    import oracle
    dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2
    nl, lim = cfg.num_layers, cfg.neighborhood_limits
    P, N, PO, UP, LE = [None] * nl, [None] * nl, [None] * nl, [None] * nl, [None] * nl
    UP[0] = np.zeros((0, 1), np.int32)
    for l in range(nl - 1):
        nb = oracle.batch_neighbors(pts, pts, lens, lens, r)[:, :lim[l]]
        pp, pl = oracle.batch_grid_subsampling(pts, lens, 2 * dl)
        po = oracle.batch_neighbors(pp, pts, pl, lens, r)[:, :lim[l]]
        up = oracle.batch_neighbors(pts, pp, lens, pl, 2 * r)[:, :lim[l]]
        P[l], N[l], PO[l], UP[l + 1], LE[l] = pts, nb, po, up, lens
        pts, lens, r, dl = pp, pl, 2 * r, 2 * dl
    P[nl - 1], LE[nl - 1] = pts, lens
    N[nl - 1] = oracle.batch_neighbors(pts, pts, lens, lens, r)[:, :lim[nl - 1]]
    PO[nl - 1] = np.zeros((0, 1), np.int32)
    return {"points": P, "neighbors": N, "pools": PO, "upsamples": UP, "batches_len": LE, "features": feats, "point_labels": labels}


def test_convnet_modules_match_float64_restatement_and_finite_differences(monkeypatch):
    """convnet.ConvNetSeg / ConvNetLoss with the libcbops operators swapped for plain-torch twins (the CUDA operators cannot run
    here), float64, CPU: (1) logits and every loss term equal the NumPy restatement to 1e-6 — the module wiring (radii, strides,
    shortcuts, upsampling, heads) is the restatement's; (2) autograd equals float64 central differences of the RESTATEMENT's loss."""
    from contrastboundary_b200 import convnet, linear_ops, synthetic, tf_pyramid
    sizes = [1600, 1300]
    scenes = [synthetic.make_scene(n, 300 + i) for i, n in enumerate(sizes)]
    pts = np.concatenate([s[0] for s in scenes])
    feats = np.concatenate([np.ones((len(pts), 1), np.float32), np.concatenate([s[1] for s in scenes]), pts[:, 2:3]], 1)
    labels = np.concatenate([s[2] for s in scenes])
    cfg = convnet.ConvNetConfig()
    inp = _cpu_pyramid(pts, feats, labels, np.array(sizes, np.int32), tf_pyramid.PyramidConfig())
    monkeypatch.setattr(convnet, "adaptive_weight", lambda q, s, nb, f, w, b, r: T.adaptive_weight(q, s, nb, f, w, b, r))
    monkeypatch.setattr(convnet, "ind_max_pool", lambda x, inds: torch.cat([x, x.min(0, keepdim=True)[0].detach()], 0)[inds.long()].max(1)[0])
    monkeypatch.setattr(convnet, "tf_contrast_loss", lambda f, nb, c, t, w: T.contrast_loss(f, nb, c.long(), t, w))
    monkeypatch.setattr(linear_ops, "FUSED_BN", False)
    torch.manual_seed(1)
    model = convnet.ConvNetSeg(cfg).double().train()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g).double())
    as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double() if a.dtype.kind == "f" else torch.from_numpy(np.ascontiguousarray(a))
    tin = {k: ([as_t(a) for a in v] if isinstance(v, list) else as_t(v)) for k, v in inp.items()}
    P0 = {k: v.detach().numpy().astype(np.float64) for k, v in model.state_dict().items() if v.dtype.is_floating_point}
    rl, rloss, _, (up0, cls) = R.forward(P0, inp, cfg)
    geo = {"up_idx0": [None] + [torch.from_numpy(u) for u in up0[1:]], "cls": [torch.from_numpy(np.asarray(c)).int() for c in cls]}
    logits, sl = model(tin, geo)
    loss = convnet.ConvNetLoss(cfg)(logits, tin["point_labels"], sl)
    # (1e-6, not 1e-12: oracle/tf_model.py keeps TF's float32 neighbour count `cnt + 1e-5`)
    assert np.abs(logits.detach().numpy() - rl).max() < 1e-6 * np.abs(rl).max()
    np.testing.assert_allclose(loss.detach().numpy(), rloss, rtol=1e-6, atol=1e-12)
    loss.sum().backward()
    params = dict(model.named_parameters())
    rng = np.random.default_rng(0)
    for gname, sel, rel in (("1x1 kernels", lambda n: n.endswith("weights.weight"), 2e-8), ("fc_1", lambda n: ".fc_1." in n, 2e-9),
                            ("batch norm", lambda n: ".bn." in n or ".pool_bn." in n, 2e-8)):
        names = [n for n in params if sel(n)]
        d = {n: rng.standard_normal(params[n].shape) for n in names}
        eps = rel / np.sqrt(sum(float((v ** 2).sum()) for v in d.values())) * np.sqrt(sum(float((P0[n] ** 2).sum()) for n in names))
        lp = R.forward({**P0, **{n: P0[n] + eps * d[n] for n in names}}, inp, cfg)[1].sum()
        lm = R.forward({**P0, **{n: P0[n] - eps * d[n] for n in names}}, inp, cfg)[1].sum()
        fd = (lp - lm) / (2 * eps)
        an = sum(float((params[n].grad.numpy() * d[n]).sum()) for n in names)
        assert abs(fd - an) <= 2e-3 * abs(fd), (gname, fd, an)


def test_product_modules_loaded_from_the_executed_reference_model(monkeypatch):
    """The PRODUCT's modules (convnet.ConvNetSeg / ConvNetLoss; libcbops operators swapped for plain-torch twins because CUDA is
    absent here; float64) with the variables of the reference's own executed SceneSegModel loaded through
    convnet.load_tf_variables (TF names -> state_dict keys, (in, out) kernels transposed): same logits, same loss entries as
    the reference model produced (tests/golden/tf_ops_ref.npz 'full/*')."""
    import os
    from contrastboundary_b200 import convnet, linear_ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    fdim, dl, density, ratio, depth, eps, wd, ncls = g["full/config"]
    cfg = convnet.ConvNetConfig(first_features_dim=int(fdim), depth=int(depth), bottleneck_ratio=int(ratio), first_subsampling_dl=float(dl),
                                density_parameter=float(density), bn_eps=float(eps), weight_decay=float(wd), num_classes=int(ncls))
    assert np.allclose(cfg.r_sample, g["full/r_sample"])
    monkeypatch.setattr(convnet, "adaptive_weight", lambda q, s, nb, f, w, b, r: T.adaptive_weight(q, s, nb, f, w, b, r))
    monkeypatch.setattr(convnet, "ind_max_pool", lambda x, inds: torch.cat([x, x.min(0, keepdim=True)[0].detach()], 0)[inds.long()].max(1)[0])
    monkeypatch.setattr(convnet, "tf_contrast_loss", lambda f, nb, c, t, w: T.contrast_loss(f, nb, c.long(), t, w))
    monkeypatch.setattr(linear_ops, "FUSED_BN", False)
    model = convnet.ConvNetSeg(cfg).double().train()
    variables = {k[len("full/var/"):]: g[k] for k in g.files if k.startswith("full/var/")}
    unused = convnet.load_tf_variables(model, variables)
    assert unused == [], unused
    trainable = {n for n, _ in model.named_parameters()}
    loaded = {convnet.tf_variable_to_state_dict(n)[0] for n in variables}
    assert trainable <= loaded                                       # every parameter of the product came from a reference variable
    as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double() if a.dtype.kind == "f" else torch.from_numpy(np.ascontiguousarray(a))
    inp = {k: [as_t(g[f"net/{k}/{l}"]) for l in range(5)] for k in ("points", "neighbors", "pools", "upsamples")}
    inp["batches_len"] = [as_t(g[f"cbl/batches_len/{l}"]) for l in range(5)]
    inp["point_labels"], inp["features"] = as_t(g["cbl/point_labels"]), as_t(g["net/features"])
    np_inp = {k: ([a.numpy() for a in v] if isinstance(v, list) else v.numpy()) for k, v in inp.items()}
    up0, cls = R.head_geometry(np_inp, cfg.r_sample, cfg.num_classes)
    geo = {"up_idx0": [None] + [torch.from_numpy(u) for u in up0[1:]], "cls": [torch.from_numpy(np.asarray(c)).int() for c in cls]}
    logits, sl = model(inp, geo)
    ref = g["full/logits"]
    assert np.abs(logits.detach().numpy() - ref).max() < 1e-6 * np.abs(ref).max()          # 1e-6: the float32 neighbour count of tf_model.py
    names = ["seg"] + [f"softnn-up{i}" for i in range(5)]
    loss_t = convnet.ConvNetLoss(cfg)(logits, inp["point_labels"], sl)
    np.testing.assert_allclose(loss_t.detach().numpy(), [float(g["full/loss/" + n]) for n in names], rtol=1e-6, atol=1e-12)
    # backward: autograd through the product's modules == central differences of the EXECUTED reference model (the golden script
    # re-ran the reference's builder at +-h along directions seeded by the TF variable names; a ladder of steps, because the loss is
    # only piecewise smooth)
    import zlib
    loss_t.sum().backward()
    params = dict(model.named_parameters())
    groups = {"kernels": lambda n: n.endswith("/weights") and "/fc_1/" not in n, "fc_1": lambda n: "/fc_1/" in n,
              "batch_norm": lambda n: n.endswith(("/gamma", "/beta"))}
    for gname, sel in groups.items():
        an = 0.0
        for tf_name in sorted(n for n in variables if sel(n)):
            key, transpose = convnet.tf_variable_to_state_dict(tf_name)
            d = np.random.default_rng(zlib.crc32(tf_name.encode())).standard_normal(variables[tf_name].shape)
            an += float((params[key].grad.numpy() * (d.T if transpose else d)).sum())
        ladder = g[f"full/dd/{gname}"]
        assert min(abs(an - fd) / abs(fd) for fd in ladder) < 2e-3, (gname, an, ladder)
