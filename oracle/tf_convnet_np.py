"""oracle/tf_convnet_np.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Second, independent restatement of the TF tree's ConvNet + CBL (BASELINE configs[2]) in NumPy float64, written from the
reference source (tensorflow/models/build_models.py:160-212, backbone/resnet.py, local_aggregation_operators.py:316-500,
heads/seg_head.py:31-95, heads/head.py:25-49,117-195,338-460,462-807, basic_operators.py:134-241,381-460) without
looking at contrastboundary_b200/convnet.py's math: forward and loss only.  Gradients of the CUDA path are checked against
it by central finite differences of this float64 loss (tests/test_convnet_gpu.py) — no autograd on this side.

PARITY UNPINNED: TensorFlow cannot be imported in the build container or on the GPU box, so neither this file nor
oracle/tf_model.py has been run against the reference itself; they are two separately written readings of the same
source that must agree with each other and with the kernels.

Weights come as a dict keyed by the product model's state_dict names (kernels stored (out, in)); the pyramid as NumPy
arrays (points / neighbors / pools / upsamples / batches_len per level, features, point_labels).
"""
import numpy as np

F64 = np.float64


def batch_norm(x, gamma, beta, eps):
    """tf.layers.batch_normalization(training=True): batch mean / biased variance (basic_operators.py:134-152)"""
    mean = x.mean(0)
    var = ((x - mean) ** 2).mean(0)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta


def conv1d_1x1(P, pre, x, eps, act=True, bn=True):
    """basic_operators.py:195-241"""
    y = x @ P[pre + ".weights.weight"].T
    if (pre + ".weights.bias") in P:
        y = y + P[pre + ".weights.bias"]
    if bn:
        y = batch_norm(y, P[pre + ".bn.weight"], P[pre + ".bn.bias"], eps)
    return np.maximum(y, 0.0) if act else y


def adaptive_weight(P, pre, query, support, idx, feat, radius, eps):
    """local_aggregation_operators.py:316-500 with adapt.yaml (dp / fc_num 1 / shared_channels 1 / mean / no softmax)"""
    n0 = support.shape[0]
    shadow_f = np.concatenate([feat, np.zeros_like(feat[:1])], 0)              # :370
    nf = shadow_f[idx]                                                          # :372  (n, K, c)
    shadow_p = np.concatenate([support, np.zeros_like(support[:1])], 0)        # :378
    rel = (shadow_p[idx] - query[:, None, :]) / radius                         # :379-382
    cw = rel @ P[pre + ".fc_1.weight"].T + P[pre + ".fc_1.bias"]              # :426-430
    agg = (cw * nf).sum(1)                                                      # :456-464
    pad = idx.max()                                                             # :466
    cnt = (idx < pad).sum(1, keepdims=True).astype(F64) + 1e-5                  # :467-470
    agg = agg / cnt
    y = batch_norm(agg, P[pre + ".pool_bn.weight"], P[pre + ".pool_bn.bias"], eps)   # :485-487
    return np.maximum(y, 0.0)                                                   # :488-489 (fdim == out_fdim: no output conv)


def ind_max_pool(x, inds):
    """basic_operators.py:155-172"""
    xs = np.concatenate([x, x.min(0, keepdims=True)], 0)
    return xs[inds].max(1)


def closest(x, idx0):
    """ind_closest_pool / tf_gather with a zero shadow row (basic_operators.py:175-192,381-409)"""
    return np.concatenate([x, np.zeros_like(x[:1])], 0)[idx0]


def bottleneck(P, pre, inp, layer, feat, radius, strided, eps):
    """resnet.py:94-300"""
    x = conv1d_1x1(P, pre + ".conv1", feat, eps)
    if strided:
        x = adaptive_weight(P, pre + ".conv2", inp["points"][layer + 1], inp["points"][layer], inp["pools"][layer], x, radius, eps)
        sc = ind_max_pool(feat, inp["pools"][layer])
    else:
        x = adaptive_weight(P, pre + ".conv2", inp["points"][layer], inp["points"][layer], inp["neighbors"][layer], x, radius, eps)
        sc = feat
    x = conv1d_1x1(P, pre + ".conv3", x, eps, act=False)
    if (pre + ".shortcut.weights.weight") in P:
        sc = conv1d_1x1(P, pre + ".shortcut", sc, eps, act=False)
    return np.maximum(x + sc, 0.0)


def _sqdist32(q, s):
    """nanoflann's float32 accumulation ((dx*dx) + dy*dy) + dz*dz (nanoflann.hpp:432-440)"""
    d = (q[:, None, :].astype(np.float32) - s[None, :, :].astype(np.float32))
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def head_geometry(inp, r_sample, ncls):
    """nearest level-i point within r_sample[i-1] of every level-0 point (head.py:133-176, kr = 1) and the hard sub-scene
    labels (head.py:25-49 reduction 'max', :117-131) — brute force per scene"""
    pts, lens, labels = inp["points"], inp["batches_len"], inp["point_labels"]
    nl = len(pts)
    up_idx0, cls = [None] * nl, [None] * nl
    cls[0] = labels.astype(np.int64)
    s0 = np.concatenate([[0], np.cumsum(lens[0])])
    for i in range(1, nl):
        si = np.concatenate([[0], np.cumsum(lens[i])])
        ni = pts[i].shape[0]
        if i == 1:
            up_idx0[i] = inp["upsamples"][1][:, 0].astype(np.int64)
            lab = np.concatenate([labels, [-1]])[inp["pools"][0]]              # tf_gather shadow -1 -> one_hot zero
            hist = np.stack([(lab == c).sum(1) for c in range(ncls)], 1)
            cls[i] = hist.argmax(1)
            continue
        r2 = np.float32(r_sample[i - 1]) * np.float32(r_sample[i - 1])
        up = np.full(pts[0].shape[0], ni, np.int64)
        ci = np.zeros(ni, np.int64)
        for b in range(len(lens[0])):
            q0, qi = pts[0][s0[b]:s0[b + 1]], pts[i][si[b]:si[b + 1]]
            if len(qi) == 0 or len(q0) == 0:
                continue
            d2 = _sqdist32(q0, qi)                                              # queries = level 0, supports = level i
            near = d2.argmin(1)
            ok = d2[np.arange(len(q0)), near] < r2
            up[s0[b]:s0[b + 1]] = np.where(ok, near + si[b], ni)
            d2t = _sqdist32(qi, q0)                                             # queries = level i, supports = level 0
            within = d2t < r2
            lab0 = labels[s0[b]:s0[b + 1]]
            hist = np.stack([(within & (lab0[None, :] == c)).sum(1) for c in range(ncls)], 1)
            ci[si[b]:si[b + 1]] = hist.argmax(1)
        up_idx0[i], cls[i] = up, ci
    return up_idx0, cls


def contrast_loss(feat, neighbors, cls, temperature, weight):
    """contrast_head with 'softnn|latent|label|l2': head.py:549-590 (neighbours without the self column, hard labels,
    valid = not shadow), :641-662 (pos / neg / point masks), :180-185 (l2 with max(., eps)), :750-771 (soft nearest
    neighbour: -log(pos / (pos + neg) + eps)), :805-806 (mean over the boundary points, x weight)"""
    eps = 1e-12                                                                  # basic_operators _eps
    n = feat.shape[0]
    idx = neighbors[:, 1:]
    valid = idx < n
    nb_cls = np.concatenate([cls, [-1]])[idx]
    posneg = nb_cls == cls[:, None]
    pos_mask, neg_mask = posneg & valid, (~posneg) & valid
    point = pos_mask.any(1) & neg_mask.any(1)
    if not point.any():
        return 0.0
    f = np.concatenate([feat, np.zeros_like(feat[:1])], 0)
    d = np.sqrt(np.maximum(((feat[:, None, :] - f[idx]) ** 2).sum(-1), eps))
    d = -d[point]
    if temperature is not None:
        d = d / temperature
    d = d - d.max(-1, keepdims=True)
    e = np.exp(d)
    pos = (e * pos_mask[point]).sum(-1)
    neg = (e * neg_mask[point]).sum(-1)
    return float((-np.log(pos / (pos + neg) + eps)).mean() * weight)


def backbone(P, inp, cfg):
    """resnet_backbone (backbone/resnet.py:307-420): input conv, simple block, per level [strided bottleneck +] depth x
    bottleneck -> list of the num_layers stage features.  P / inp as float64."""
    eps = cfg.bn_eps
    r = cfg.first_subsampling_dl * cfg.density_parameter                        # build_models.py:186
    pts = inp["points"]
    x = conv1d_1x1(P, "resnet_backbone.res1_input_conv", np.asarray(inp["features"], F64), eps)
    x = adaptive_weight(P, "resnet_backbone.res1_simple_block", pts[0], pts[0], inp["neighbors"][0], x, r, eps)
    feats = []
    for l in range(cfg.num_layers):
        pre = f"resnet_backbone.res.{l}"
        if l > 0:
            x = bottleneck(P, pre + ".strided_bottleneck", inp, l - 1, x, r * 2 ** (l - 1), True, eps)
        for i in range(cfg.depth):
            x = bottleneck(P, pre + f".bottleneck{i}", inp, l, x, r * 2 ** l, False, eps)
        feats.append(x)
    return feats


def seg_head_features(P, inp, feats, cfg):
    """resnet_scene_segmentation_head with sep_head (heads/seg_head.py:58-95): nearest upsampling + concat + 1x1 conv,
    -> [F_up[0] (finest) .. F_up[3], feats[4]]"""
    eps = cfg.bn_eps
    f_up = []
    x = feats[4]
    for j in range(4):
        lvl = 4 - j
        x = closest(x, inp["upsamples"][lvl][:, 0])
        x = conv1d_1x1(P, f"resnet_scene_segmentation_head.up_conv{j}", np.concatenate([x, feats[lvl - 1]], 1), eps)
        f_up.append(x)
    return list(reversed(f_up)) + [feats[4]]


def multiscale_head(P, f_out, up_idx0, labels, cfg):
    """multiscale_head '||Ua-concat-latent' (heads/head.py:338-425): latent_i = relu(bn(f_out_i W_i)) (mlps_by_ops '1mlp', :268-273),
    nearest upsampling of every latent to U0 (a level-0 point without a level-i point in range gathers a zero row), concat, linear
    classifier (:279-282), mean sparse softmax cross-entropy (:200-214) -> (logits, latents, loss)"""
    eps = cfg.bn_eps
    latents = [conv1d_1x1(P, f"multiscale.mlp.{i}", f_out[i], eps) for i in range(cfg.num_layers)]
    cols = [latents[0]] + [closest(latents[i], up_idx0[i]) for i in range(1, cfg.num_layers)]
    logits = np.concatenate(cols, 1) @ P["multiscale.linear.weight"].T + P["multiscale.linear.bias"]
    labels = np.asarray(labels).astype(np.int64)
    z = logits - logits.max(1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(1, keepdims=True))
    return logits, latents, float(-logp[np.arange(len(labels)), labels].mean())


def forward(P, inp, cfg):
    """P: {name: float64 array}; inp: pyramid dict of NumPy arrays; cfg: contrastboundary_b200.convnet.ConvNetConfig-like
    -> (logits (n0, ncls), loss vector [xen, cbl_0..cbl_4], latents)"""
    P = {k: np.asarray(v, F64) for k, v in P.items()}
    eps = cfg.bn_eps
    f = cfg.first_features_dim
    r = cfg.first_subsampling_dl * cfg.density_parameter                        # build_models.py:186
    pts = [np.asarray(p, F64) for p in inp["points"]]
    inp = dict(inp, points=pts)
    feats = backbone(P, inp, cfg)
    f_out = seg_head_features(P, inp, feats, cfg)
    # multiscale head
    up_idx0, cls = head_geometry({k: ([np.asarray(a) for a in v] if isinstance(v, (list, tuple)) else np.asarray(v)) for k, v in inp.items()
                                  if k in ("points", "batches_len", "point_labels", "upsamples", "pools")}, cfg.r_sample, cfg.num_classes)
    logits, latents, xen = multiscale_head(P, f_out, up_idx0, np.asarray(inp["point_labels"]), cfg)
    losses = [xen]
    if cfg.contrast:
        for i in range(cfg.num_layers):
            losses.append(contrast_loss(latents[i], np.asarray(inp["neighbors"][i]), np.asarray(cls[i]), cfg.contrast_temperature,
                                        cfg.contrast_weight))
    return logits, np.asarray(losses, F64), latents, (up_idx0, cls)
