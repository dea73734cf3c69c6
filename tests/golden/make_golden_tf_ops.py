"""Golden vectors for the TF tree's ConvNet (BASELINE config 3), produced by EXECUTING the reference's own source files,
unmodified, from /root/reference — TensorFlow itself is not installable in the build container, so `tensorflow` is the NumPy
stand-in of the TF-1 API in tf_numpy_shim.py (eager meaning of every op the files call, float64, variables in a scope-keyed
dictionary):

    tensorflow/models/local_aggregation_operators.py   AdaptiveWeight (config/s3dis/adapt.yaml)                      row a13
    tensorflow/models/basic_operators.py, utils.py      conv1d_1x1, batch_norm, ind_max_pool, ind_closest_pool, dense_layer, mlps, ...
    tensorflow/models/backbone/resnet.py                resnet_backbone (input conv, simple block, strided / plain bottlenecks)
    tensorflow/models/heads/seg_head.py                 resnet_scene_segmentation_head (sep_head)
    tensorflow/models/heads/head.py                     multiscale_head ('||Ua-concat-latent') + cross-entropy,
                                                        contrast_head (label sampling, hard sub-scene labels, soft-NN)       row a14
    tensorflow/config/head.py                           the reference's own head-config objects (multiscale_1, contrast_0)

The 5-level input pyramid and the heads' cross-stage radius searches come from the CPU oracle of the reference's C++ operators
(oracle/, itself pinned against the reference's compiled C++).

    python tests/golden/make_golden_tf_ops.py        -> tests/golden/tf_ops_ref.npz

Consumers: tests/test_convnet_cpu.py (both restatements == these vectors; the set of L2-regularised variables ==
ConvNetSeg.decay_parameters(); TF variable names through convnet.tf_variable_to_state_dict) and tests/test_tfops_gpu.py (the CUDA
AdaptiveWeight kernel, label votes and soft-NN loss == these vectors); the CUDA network is tested against the restatement
(tests/test_convnet_gpu.py)."""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CB_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_reference_models():
    """the reference's tensorflow/models/*.py as package `refmodels`, with `tensorflow` := the NumPy shim; the package's own
    __init__ (which pulls in the whole model zoo) is not executed"""
    import tf_numpy_shim as shim
    sys.modules["tensorflow"] = shim
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "tensorflow", "models")]
    sys.modules["refmodels"] = pkg
    lao = importlib.import_module("refmodels.local_aggregation_operators")
    return shim, lao


def pyramid(points, lens, dl, density, limits):
    """the reference's input pyramid (tensorflow/datasets/base.py:767-842) composed from the CPU oracle of its C++ operators"""
    import oracle
    pts, lens = points.astype(np.float32), np.asarray(lens, np.int32)
    r = dl * density / 2.0
    nl = len(limits)
    P, NB, PO, UP = [None] * nl, [None] * nl, [None] * nl, [np.zeros((0, 1), np.int32)] + [None] * (nl - 1)
    LENS = [None] * nl
    for l in range(nl - 1):
        LENS[l] = lens
        NB[l] = oracle.batch_neighbors(pts, pts, lens, lens, r)[:, :limits[l]]
        pool_pts, pool_lens = oracle.batch_grid_subsampling(pts, lens, 2 * dl)
        PO[l] = oracle.batch_neighbors(pool_pts, pts, pool_lens, lens, r)[:, :limits[l]]
        UP[l + 1] = oracle.batch_neighbors(pts, pool_pts, lens, pool_lens, 2 * r)[:, :limits[l]]
        P[l] = pts
        pts, lens, r, dl = pool_pts, pool_lens, 2 * r, 2 * dl
    P[nl - 1], LENS[nl - 1] = pts, lens
    NB[nl - 1] = oracle.batch_neighbors(pts, pts, lens, lens, r)[:, :limits[nl - 1]]
    PO[nl - 1] = np.zeros((0, 1), np.int32)
    return {"points": P, "neighbors": NB, "pools": PO, "upsamples": UP, "batches_len": LENS}


def scene(n, seed):
    from contrastboundary_b200 import synthetic
    return synthetic.make_scene(n, seed)[0].astype(np.float64)


def main():
    import oracle
    shim, lao = load_reference_models()
    out = {}
    cfg = types.SimpleNamespace(adaptive_weight=types.SimpleNamespace(          # config/s3dis/adapt.yaml:19-26
        local_input_feature="dp", reduction="mean", shared_channels=1, fc_num=1, weight_softmax=False, output_conv=False))
    rng = np.random.default_rng(7)
    cases = {"self": (800, 800, 0.14, 24), "pool": (600, 200, 0.12, 72)}
    for name, (n0, n, radius, fdim) in cases.items():
        sup = scene(n0, 11).astype(np.float32)
        qry = sup if name == "self" else sup[rng.choice(n0, n, replace=False)]
        lens_s, lens_q = np.array([n0], np.int32), np.array([len(qry)], np.int32)
        nb = oracle.batch_neighbors(qry, sup, lens_q, lens_s, radius).astype(np.int64)      # shadow index = n0 (reference C++ semantics)
        nb = nb[:, :26]                                                                       # neighborhood_limits[0] of the shipped config
        feat = rng.standard_normal((n0, fdim))
        shim.reset(seed=3)
        res = lao.AdaptiveWeight(cfg, qry.astype(np.float64), sup.astype(np.float64), nb, feat, scope="aw", radius=radius, out_fdim=fdim,
                                 is_training=True, init="xavier", weight_decay=0, activation_fn="relu", bn=True, bn_momentum=0.98, bn_eps=1e-3)
        v, taps = shim.variables(), shim.taps()
        out[f"aw/{name}/query"], out[f"aw/{name}/support"], out[f"aw/{name}/neighbors"] = qry, sup, nb.astype(np.int32)
        out[f"aw/{name}/features"], out[f"aw/{name}/radius"] = feat, np.float64(radius)
        out[f"aw/{name}/fc_weight"] = v["aw/fc_1/weights"]                # (3, fdim): TF kernels are (in, out)
        out[f"aw/{name}/fc_bias"] = v["aw/fc_1/biases"]
        out[f"aw/{name}/bn_gamma"], out[f"aw/{name}/bn_beta"] = v["aw/pool_bn/gamma"], v["aw/pool_bn/beta"]
        out[f"aw/{name}/aggregated"] = taps["aw/pool_bn/input"]           # what enters pool_bn: the aggregation itself
        out[f"aw/{name}/output"] = res                                    # relu(pool_bn(aggregated))
        print(name, "neighbors", nb.shape, "shadow fraction %.2f" % float((nb == n0).mean()), "output", res.shape)
    # ---- the backbone and the segmentation head: backbone/resnet.py:307-420 + heads/seg_head.py:31-110, executed -----------------------
    resnet = importlib.import_module("refmodels.backbone.resnet")
    seg = importlib.import_module("refmodels.heads.seg_head")
    fdim, dl, density, limits = 4, 0.04, 5.0, [12, 14, 16, 16, 14]
    cfg2 = types.SimpleNamespace(adaptive_weight=cfg.adaptive_weight, local_aggreagtion="adaptive_weight", sep_head=True, arch_up="", num_layers=5,
                                 num_classes=13)
    pts = np.concatenate([scene(900, 21), scene(700, 22) + np.array([30.0, 0, 0])]).astype(np.float32)
    pyr = pyramid(pts, [900, 700], dl, density, limits)
    feat_in = rng.standard_normal((len(pts), 5))
    inputs = {k: [np.asarray(a, np.float64) if k == "points" else np.asarray(a, np.int64) for a in v] for k, v in pyr.items()}
    lens0 = [900, 700]
    shim.reset(seed=5)
    F = resnet.resnet_backbone(cfg2, inputs, feat_in, base_radius=dl * density, base_fdim=fdim, bottleneck_ratio=2, depth=1, is_training=True,
                               init="xavier", weight_decay=0, activation_fn="relu", bn=True, bn_momentum=0.98, bn_eps=1e-3)
    F_up, head = seg.resnet_scene_segmentation_head(cfg2, inputs, F, base_fdim=fdim, is_training=True, init="xavier", weight_decay=0,
                                                    activation_fn="relu", bn=True, bn_momentum=0.98, bn_eps=1e-3)
    assert head is None and len(F) == 5 and len(F_up) == 4
    for k, v in pyr.items():
        for l, a in enumerate(v):
            out[f"net/{k}/{l}"] = a
    out["net/features"] = feat_in
    out["net/config"] = np.array([fdim, dl, density, 2, 1], np.float64)      # first_features_dim, dl, density, bottleneck_ratio, depth
    for l in range(5):
        out[f"net/F/{l}"] = F[l]
    for l in range(4):
        out[f"net/F_up/{l}"] = F_up[l]
    for name, value in shim.variables().items():
        out["net/var/" + name] = value
    print("backbone + seg head: level sizes", [len(p) for p in pyr["points"]], "variables", len(shim.variables()))
    # which variables carry an L2 weight loss when weight_decay > 0 (basic_operators.py:126-129,371-379): the same code once more
    shim.reset(seed=5)
    F2 = resnet.resnet_backbone(cfg2, inputs, feat_in, base_radius=dl * density, base_fdim=fdim, bottleneck_ratio=2, depth=1, is_training=True,
                                init="xavier", weight_decay=1e-3, activation_fn="relu", bn=True, bn_momentum=0.98, bn_eps=1e-3)
    seg.resnet_scene_segmentation_head(cfg2, inputs, F2, base_fdim=fdim, is_training=True, init="xavier", weight_decay=1e-3,
                                       activation_fn="relu", bn=True, bn_momentum=0.98, bn_eps=1e-3)
    assert all(np.array_equal(a, b) for a, b in zip(F, F2))
    l2_names = list(shim.taps()["l2_loss_variables"])

    # ---- a14: the contrast head (heads/head.py:462-807) on the same pyramid: label sampling ('label'), hard sub-scene labels
    #      (get_scene_label 'max': pools for stage 1, a radius search among the level-0 points beyond), soft-NN on l2 distances --------------
    from contrastboundary_b200 import synthetic
    import oracle
    head = importlib.import_module("refmodels.heads.head")
    labels0 = np.concatenate([synthetic.make_scene(900, 21)[2], synthetic.make_scene(700, 22)[2]]).astype(np.int64)
    r_sample = [dl * density * 2 ** i / 2.0 * 1.5 for i in range(4)]          # any radii work: they are inputs of both sides

    def radius_search(queries, supports, q_len, s_len, radius, device=None):   # what ops.get_tf_func('radius') wraps: the reference's C++
        return oracle.batch_neighbors(np.asarray(queries, np.float32), np.asarray(supports, np.float32), np.asarray(q_len, np.int32),
                                      np.asarray(s_len, np.int32), float(radius)).astype(np.int64)
    sys.modules["ops"] = types.SimpleNamespace(get_tf_func=lambda name: radius_search)
    d = 32
    latents = [rng.standard_normal((len(p), d)) for p in pyr["points"]]
    class GraphTensor(np.ndarray):          # TF-1 graph tensors compare by identity (head.py:146 asserts `pts_from == pts_to`)
        def __eq__(self, other):
            return self is other
        __hash__ = None
    stage = [{"p_out": inputs["points"][i].view(GraphTensor), "latent": latents[i], "f_out": latents[i]} for i in range(5)]
    cinputs = {"neighbors": inputs["neighbors"], "point_labels": labels0, "points": inputs["points"], "stage_list": {"up": stage, "down": stage},
               "sample_idx": {"down": inputs["pools"], "up": inputs["upsamples"]}, "batches_len": [np.asarray(x, np.int64) for x in pyr["batches_len"]],
               "_glb": {}}
    ccfg = types.SimpleNamespace(search="radius", sample="grid", ignored_labels=[], num_classes=13, num_layers=5, r_sample=r_sample, debug=False)
    hcfg = types.SimpleNamespace(sample="label", dist="l2", margin="", mask="", contrast_aug="", weight=0.1)
    out["cbl/point_labels"], out["cbl/r_sample"] = labels0, np.asarray(r_sample)
    for i in range(5):
        samples = head.contrast_head.sample_labels(cinputs, "up", i, "label", "latent", ccfg, name=f"up{i}/sample")
        res = head.contrast_head.contrast(latents[i], (*samples, "up", i), "softnn", cinputs, hcfg, ccfg, name=f"up{i}/softnn")
        scene_lab = head.get_scene_label(cinputs, "up", i, "latent", ccfg, reduction="max", extend=False, infer="")
        out[f"cbl/latent/{i}"] = latents[i]
        out[f"cbl/batches_len/{i}"] = np.asarray(pyr["batches_len"][i], np.int32)
        out[f"cbl/loss/{i}"] = np.float64(res["loss"])
        out[f"cbl/scene_label/{i}"] = labels0 if scene_lab is None else np.asarray(scene_lab).reshape(-1).astype(np.int64)
        print("contrast head stage", i, "loss", float(res["loss"]), "points with pos and neg:", int(np.asarray(res["logits"]).shape[0]))

    # ---- the multi-scale head '||Ua-concat-latent' (heads/head.py:338-460) with the reference's OWN head-config object
    #      (config/head.py: multiscale_1) on the executed backbone's features: latent_i = mlp(f_out_i), nearest upsampling to U0,
    #      concat, linear classifier, sparse softmax cross-entropy --------------------------------------------------------------------------
    cpkg = types.ModuleType("refconfig")
    cpkg.__path__ = [os.path.join(REF, "tensorflow", "config")]
    sys.modules["refconfig"] = cpkg
    hc = importlib.import_module("refconfig.head")
    mcfg = hc.main_dict["multiscale_1"]
    assert mcfg._ops == "||Ua-concat-latent" and hc.contrast_dict["contrast_0"]._ops == "softnn|latent|label|l2||w.1|Ua"

    class GlobalConfig:                      # config/base.py: a missing attribute reads as ''
        def __getattr__(self, name):
            return ""
    gc = GlobalConfig()
    gc.__dict__.update(num_layers=5, num_classes=13, first_features_dim=fdim, init="xavier", weight_decay=1e-3, bn_momentum=0.99, bn_eps=1e-6,
                       activation="relu", search="radius", sample="grid", ignored_labels=[], debug=False,
                       r_sample=[0.6 * dl * 2 ** (i + 1) for i in range(4)])   # config/s3dis.py:87, shrunk so that some
    #                                                   level-0 points find NO level-i point in range (the zero-row gather path)
    f_out = [F_up[0], F_up[1], F_up[2], F_up[3], F[4]]
    mstage = [{"p_out": stage[i]["p_out"], "f_out": f_out[i]} for i in range(5)]
    minputs = dict(cinputs, stage_list={"up": mstage, "down": mstage}, _glb={})
    shim.reset(seed=9)
    with shim.variable_scope("multi"):
        hd = head.multiscale_head()(minputs, mcfg, gc, True)
    out["multi/logits"], out["multi/loss"] = hd["logits"]["seg"], np.float64(hd["loss"]["seg"])
    out["multi/r_sample"] = np.asarray(gc.r_sample)
    for i in range(5):
        out[f"multi/latent/{i}"] = mstage[i]["latent"]
    for name, value in shim.variables().items():
        out["multi/var/" + name] = value
    l2_names += list(shim.taps()["l2_loss_variables"])
    out["l2_loss_variables"] = np.array(sorted(l2_names))
    shadow = {k: int((np.asarray(v) == len(pyr["points"][int(k.split("-")[0][-1])])).sum()) for k, v in minputs["_glb"].items() if "sample_neighbor" in k}
    print("multi-scale head: loss", float(hd["loss"]["seg"]), "logits", hd["logits"]["seg"].shape, "variables", sorted(shim.variables())[:4], "...",
          "level-0 points without a level-i point in range:", shadow)
    # ---- the WHOLE model: the reference's own builder (models/build_models.py SceneSegModel), driven by its own config object
    #      (config/s3dis.py Conv '|multi-Ua-concat-latent|contrast-Ua-softnn-latent-label-l2-w.1' + config/s3dis/adapt.yaml), on the
    #      same pyramid: backbone -> seg head -> build_head (load_config + apply_head_ops for both heads) -> build_loss -----------------------
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "tensorflow"))          # the reference resolves 'config/s3dis/adapt.yaml' relative to its tree
    sys.path.insert(0, os.path.join(REF, "tensorflow"))
    sys.modules["termcolor"] = types.SimpleNamespace(colored=lambda s_, *a, **k: s_)       # utils/logger.py colours its prints
    try:
        import config as refcfg
        import models as refmodels_full
        rcfg = refcfg.Config(refcfg.load_config(dataset_name="s3dis", cfg_name="conv_multi-Ua-concat-latent_contrast-Ua-softnn-latent-label-l2-w.1"))
        assert rcfg.arch_out == ["multi-Ua-concat-latent", "contrast-Ua-softnn-latent-label-l2-w.1"] and rcfg.depth == 1 and rcfg.bn_eps == 1e-6
        rcfg.update({"first_features_dim": 4}, exclude=[])                                 # 72 in adapt.yaml: smaller vectors, same code
        finputs = {"points": [np.asarray(p, np.float64).view(GraphTensor) for p in pyr["points"]],
                   "neighbors": inputs["neighbors"], "pools": inputs["pools"], "upsamples": inputs["upsamples"],
                   "batches_len": [np.asarray(x, np.int64) for x in pyr["batches_len"]], "features": feat_in, "point_labels": labels0,
                   "in_batches": None, "out_batches": None}
        shim.reset(seed=13)
        model = refmodels_full.SceneSegModel(finputs, True, rcfg, scope=None, verbose=False)
    finally:
        os.chdir(cwd)
    # ---- the input pyramid itself: datasets/base.py Dataset.tf_segmentation_inputs_radius (:767-842) executed, its two custom ops
    #      (ops.get_tf_func 'grid' / 'radius') served by the CPU oracle of the reference's C++ -------------------------------------------------
    def batch_subsampling(points, lens_, sampleDl=0.1, **_):
        p_, l_ = oracle.batch_grid_subsampling(np.asarray(points, np.float32), np.asarray(lens_, np.int32), float(sampleDl))
        return p_, l_
    sys.modules["ops"] = types.SimpleNamespace(get_tf_func=lambda name, verbose=False: {"grid": batch_subsampling, "radius": radius_search}[name])
    sys.modules.pop("datasets", None)
    base = importlib.import_module("datasets.base")
    ds = object.__new__(base.Dataset)
    ds.config, ds.verbose, ds.neighborhood_limits = rcfg, False, limits
    binds = np.repeat(np.arange(2), lens0)
    ref_pyr = ds.tf_segmentation_inputs_radius(pts, feat_in, labels0, np.asarray(lens0, np.int32), binds)
    for key in ("points", "neighbors", "pools", "upsamples", "batches_len"):
        for l in range(5):
            assert np.array_equal(np.asarray(ref_pyr[key][l]), np.asarray(pyr[key][l])), (key, l)     # == the arrays stored as net/*
    out["pyr/in_batches"], out["pyr/out_batches"] = np.asarray(ref_pyr["in_batches"]), np.asarray(ref_pyr["out_batches"])
    out["pyr/batch_weights"] = np.asarray(ref_pyr["batch_weights"], np.float64)
    print("pyramid: datasets/base.py == the oracle composition at all 5 levels; in_batches", out["pyr/in_batches"].shape, "out_batches",
          out["pyr/out_batches"].shape)
    out["full/config"] = np.array([rcfg.first_features_dim, rcfg.first_subsampling_dl, rcfg.density_parameter, rcfg.bottleneck_ratio, rcfg.depth,
                                   rcfg.bn_eps, rcfg.weight_decay, rcfg.num_classes], np.float64)
    out["full/r_sample"] = np.asarray(rcfg.r_sample, np.float64)
    out["full/logits"] = model.head_dict["result"]["seg"]["logits"]
    for k_, v_ in model.loss_dict.items():
        out["full/loss/" + k_] = np.float64(v_)
    for name, value in shim.variables().items():
        out["full/var/" + name] = value
    print("whole model:", {k_: round(float(v_), 6) for k_, v_ in model.loss_dict.items()}, "variables", len(shim.variables()))

    # ---- gradients of the executed model by central differences (the stand-in has no autodiff): directional derivatives of the sum of
    #      the head losses (the loss vector the product back-propagates; the L2 term lives in its optimiser) along seeded directions ----------
    import zlib
    base_vars = {k_: v_.copy() for k_, v_ in shim.variables().items()}
    groups = {"kernels": lambda n: n.endswith("/weights") and "/fc_1/" not in n, "fc_1": lambda n: "/fc_1/" in n,
              "batch_norm": lambda n: n.endswith(("/gamma", "/beta"))}

    def head_loss(values):
        os.chdir(os.path.join(REF, "tensorflow"))
        try:
            shim.reset(seed=13)
            shim.preset(values)
            m_ = refmodels_full.SceneSegModel(dict(finputs), True, rcfg, scope=None, verbose=False)
        finally:
            os.chdir(cwd)
        return float(m_.loss_dict["loss"]) - float(m_.loss_dict["l2_loss"])
    assert abs(head_loss(base_vars) - (float(model.loss_dict["loss"]) - float(model.loss_dict["l2_loss"]))) < 1e-12
    for gname, sel in groups.items():
        names = sorted(n for n in base_vars if sel(n))
        direction = {n: np.random.default_rng(zlib.crc32(n.encode())).standard_normal(base_vars[n].shape) for n in names}
        scale = np.sqrt(sum(float((base_vars[n] ** 2).sum()) for n in names)) / np.sqrt(sum(float((d_ ** 2).sum()) for d_ in direction.values()))
        vals = []
        for rel in (2e-7, 2e-8, 2e-9):           # a ladder: the loss is piecewise smooth (ReLU kinks), small steps see fewer kinks
            h = rel * scale
            lp = head_loss({**base_vars, **{n: base_vars[n] + h * direction[n] for n in names}})
            lm = head_loss({**base_vars, **{n: base_vars[n] - h * direction[n] for n in names}})
            vals.append((lp - lm) / (2 * h))
        out[f"full/dd/{gname}"] = np.asarray(vals, np.float64)
        print("directional derivative", gname, len(names), "variables, steps 2e-7 / 2e-8 / 2e-9:", vals)
    np.savez_compressed(os.path.join(HERE, "tf_ops_ref.npz"), **out)
    print("wrote tests/golden/tf_ops_ref.npz", sum(a.nbytes for a in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
