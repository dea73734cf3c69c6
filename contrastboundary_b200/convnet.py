"""BASELINE configs[2] — the TF tree's "ConvNet": a ResNet of AdaptiveWeight local aggregations + nearest-upsample
segmentation head + multi-scale head + contrastive boundary loss, on the device-built radius pyramid.

Host-side mirror of (paths relative to the reference's tensorflow/ directory)
    models/build_models.py:160-212            SceneSegModel (backbone -> seg head -> heads -> loss)
    models/backbone/resnet.py:40-480          simple_block / bottleneck / strided_bottleneck / resnet_backbone
    models/local_aggregation_operators.py:316-500   AdaptiveWeight with config/s3dis/adapt.yaml (dp, one FC, mean, no softmax)
    models/heads/seg_head.py:13-110           nearest_upsample_block / resnet_scene_segmentation_head (sep_head: F_up only)
    models/heads/head.py:338-460              multiscale_head '||Ua-concat-latent' (per-stage latent MLP, nearest upsample to
                                              U0, concat, linear classifier, cross entropy)
    models/heads/head.py:462-807              contrast_head 'softnn|latent|label|l2||w.1|Ua'
    config/s3dis.py:17-160, config/s3dis/adapt.yaml   (conv_0 = '|multi-Ua-concat-latent|contrast-Ua-softnn-latent-label-l2-w.1')
as torch modules whose hot operators are libcbops kernels: cb_adaptive_weight_* (the aggregation), cb_ind_max_pool_*,
cb_label_vote_* (hard sub-scene labels), cb_cbl_*_ex (the loss), tall-skinny tensor-core linears and fused BatchNorm+ReLU for
the 1x1 convolutions; the input pyramid comes from tf_pyramid.segmentation_inputs_radius (cb_radius_*, cb_grid_subsample_*).
TensorFlow is not part of this stack (it is not installable here): variable scopes became module names
(`resnet_backbone.res2_strided_bottleneck.conv2.local_aggregation.fc_1` ...), kernels are stored (out, in) as torch does.

PARITY: no TensorFlow in the build container or on the GPU box.  AdaptiveWeight, the resnet backbone, the segmentation
head, the multi-scale head + cross-entropy and the contrast head (hard sub-scene labels, soft-NN loss) are pinned by vectors
from the reference's own TF source executed on a NumPy stand-in of the TF-1 API (tests/golden/make_golden_tf_ops.py ->
tf_ops_ref.npz); the gradients are checked against two independent restatements of the reference source — oracle/tf_model.py (operators, torch) and oracle/tf_convnet_np.py (the
whole network and loss, NumPy float64) — plus a finite-difference check of the gradients against the float64 restatement
(tests/test_convnet_gpu.py).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib as L
from . import tf_ops
from .linear_ops import BatchNorm1d, Linear, bn_act, flush_bn_counters
from .tf_model import adaptive_weight, tf_contrast_loss


@dataclass
class ConvNetConfig:            # config/s3dis.py Default + config/s3dis/adapt.yaml
    num_classes: int = 13
    in_features_dim: int = 5                # '1-rgb-Z'
    first_features_dim: int = 72
    num_layers: int = 5
    depth: int = 1
    bottleneck_ratio: int = 2
    first_subsampling_dl: float = 0.04
    density_parameter: float = 5.0
    neighborhood_limits: List[int] = field(default_factory=lambda: [26, 31, 38, 41, 39, 29])
    bn_momentum: float = 0.99               # TF decay; torch momentum = 1 - decay
    bn_eps: float = 1e-6
    weight_decay: float = 1e-3
    contrast: bool = True                   # 'contrast-Ua-softnn-latent-label-l2-w.1'
    contrast_weight: float = 0.1
    contrast_temperature: Optional[float] = None
    # optimiser (adapt.yaml)
    base_learning_rate: float = 0.02
    momentum: float = 0.98
    grad_norm: float = 100.0

    @property
    def r_sample(self):                     # config/s3dis.py:87
        return [self.first_subsampling_dl * 2 ** (i + 1) for i in range(self.num_layers - 1)]


# ------------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------------
class _IndMaxPoolFn(Function):
    @staticmethod
    def forward(ctx, x, inds):
        x = x.contiguous()
        inds = inds.contiguous()
        n1, c = x.shape
        n2, k = inds.shape
        out = torch.empty((n2, c), dtype=torch.float32, device=x.device)
        arg = torch.empty((n2, c), dtype=torch.uint8, device=x.device)
        colmin = x.min(0)[0].contiguous() if n1 > 0 else x.new_zeros(c)        # the reference's shadow row (basic_operators.py:168)
        L.call("cb_ind_max_pool_forward", n2, k, c, n1, x, colmin, inds, out, arg, L.stream())
        ctx.save_for_backward(inds, arg)
        ctx.n1 = n1
        return out

    @staticmethod
    def backward(ctx, g):
        inds, arg = ctx.saved_tensors
        n2, k = inds.shape
        c = g.shape[1]
        gx = torch.zeros((ctx.n1, c), dtype=torch.float32, device=g.device)
        L.call("cb_ind_max_pool_backward", n2, k, c, inds, arg, g.contiguous(), gx, L.stream())
        return gx, None


def ind_max_pool(x, inds):
    """basic_operators.py:155-172"""
    if x.shape[1] % 4 or inds.shape[1] >= 255:
        xs = torch.cat([x, x.min(0, keepdim=True)[0]], 0)
        return xs[inds.long()].max(1)[0]
    return _IndMaxPoolFn.apply(x, inds)


def closest_pool(x, idx0):
    """ind_closest_pool (basic_operators.py:175-192) / tf_gather with a zero shadow row: rows x[idx0], zeros where idx0 is the
    shadow index len(x)"""
    return torch.cat([x, x.new_zeros(1, x.shape[1])], 0)[idx0.long()]


def label_vote_idx(label_idx, target, ncls, n_valid):
    cls = torch.empty(label_idx.shape[0], dtype=torch.int32, device=target.device)
    L.call("cb_label_vote_idx", label_idx.shape[0], label_idx.shape[1], ncls, n_valid, label_idx.contiguous(), target.contiguous(), cls,
           L.stream())
    return cls


def label_vote_radius(queries, supports, q_lens, s_lens, radius, target, ncls):
    q, s = queries.contiguous(), supports.contiguous()
    qo, so = tf_ops._offsets(q_lens.int()), tf_ops._offsets(s_lens.int())
    lib = L.lib()
    ws = L.workspace(lib.cb_knn_workspace_bytes(s.shape[0], 0, qo.shape[0]), q.device, "knn")
    cls = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    rc = lib.cb_label_vote_radius(C.c_int(q.shape[0]), L.ptr(q), C.c_int(s.shape[0]), L.ptr(s), L.ptr(qo), L.ptr(so), C.c_int(qo.shape[0]),
                                  C.c_float(float(radius)), C.c_int(ncls), L.ptr(target.contiguous()), L.ptr(cls), L.ptr(ws),
                                  C.c_size_t(ws.numel()), L.stream())
    L.check(rc, "cb_label_vote_radius")
    return cls


# ------------------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------------------
def _xavier(linear):
    nn.init.xavier_uniform_(linear.weight)              # tf.glorot_uniform_initializer (basic_operators.py:117)
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


class Conv1x1(nn.Module):
    """conv1d_1x1 (basic_operators.py:195-241): matmul (+ bias) -> batch norm -> relu"""

    def __init__(self, cfg, in_fdim, out_fdim, with_bias=False, bn=True, act=True):
        super().__init__()
        self.weights = Linear(in_fdim, out_fdim, bias=with_bias)
        _xavier(self.weights)
        self.bn = BatchNorm1d(out_fdim, eps=cfg.bn_eps, momentum=1.0 - cfg.bn_momentum) if bn else None
        self.act = act

    def forward(self, x):
        y = self.weights(x)
        if self.bn is not None:
            return bn_act(self.bn, y, relu=self.act)
        return F.relu(y) if self.act else y


class LocalAggregation(nn.Module):
    """AdaptiveWeight (local_aggregation_operators.py:316-500) with adapt.yaml: conv weight = fc_1(dp), element-wise
    product with the neighbour features, mean over the valid neighbours, pool_bn, relu (no output conv: fdim == out_fdim)"""

    def __init__(self, cfg, fdim, out_fdim=None):
        super().__init__()
        self.fc_1 = nn.Linear(3, fdim)                                         # batch_conv1d_1x1 'fc_1', init 'fan_in', wd 0
        nn.init.trunc_normal_(self.fc_1.weight, std=(1.0 / 3.0) ** 0.5, a=-2 * (1.0 / 3.0) ** 0.5, b=2 * (1.0 / 3.0) ** 0.5)
        nn.init.zeros_(self.fc_1.bias)
        self.pool_bn = BatchNorm1d(fdim, eps=cfg.bn_eps, momentum=1.0 - cfg.bn_momentum)
        self.output_conv = Conv1x1(cfg, fdim, out_fdim, with_bias=False) if out_fdim not in (None, fdim) else None

    def forward(self, query_points, support_points, neighbors, features, radius):
        agg = adaptive_weight(query_points, support_points, neighbors, features, self.fc_1.weight, self.fc_1.bias, radius)
        out = bn_act(self.pool_bn, agg, relu=True)
        return self.output_conv(out) if self.output_conv is not None else out


class Bottleneck(nn.Module):
    """bottleneck / strided_bottleneck (resnet.py:94-300)"""

    def __init__(self, cfg, in_fdim, out_fdim, strided):
        super().__init__()
        mid = out_fdim // cfg.bottleneck_ratio
        self.strided = strided
        self.conv1 = Conv1x1(cfg, in_fdim, mid)
        self.conv2 = LocalAggregation(cfg, mid)
        self.conv3 = Conv1x1(cfg, mid, out_fdim, act=False)
        self.shortcut = Conv1x1(cfg, in_fdim, out_fdim, act=False) if in_fdim != out_fdim else None

    def forward(self, inputs, layer_ind, features, radius):
        pts = inputs["points"]
        x = self.conv1(features)
        if self.strided:       # queries = the next level, neighbours = the pooling rows (resnet.py:240-255)
            x = self.conv2(pts[layer_ind + 1], pts[layer_ind], inputs["pools"][layer_ind], x, radius)
            shortcut = ind_max_pool(features, inputs["pools"][layer_ind])
        else:
            x = self.conv2(pts[layer_ind], pts[layer_ind], inputs["neighbors"][layer_ind], x, radius)
            shortcut = features
        x = self.conv3(x)
        if self.shortcut is not None:
            shortcut = self.shortcut(shortcut)
        return F.relu(x + shortcut)


class ResnetBackbone(nn.Module):
    """resnet_backbone (resnet.py:303-480), num_layers = 5"""

    def __init__(self, cfg):
        super().__init__()
        f = cfg.first_features_dim
        self.res1_input_conv = Conv1x1(cfg, cfg.in_features_dim, f)
        self.res1_simple_block = LocalAggregation(cfg, f)
        stages = []
        in_f = f
        for l in range(cfg.num_layers):
            out_f = f * 2 ** (l + 1)
            blocks = nn.ModuleDict()
            if l > 0:
                blocks["strided_bottleneck"] = Bottleneck(cfg, in_f, out_f, True)
                in_f = out_f
            for i in range(cfg.depth):
                blocks[f"bottleneck{i}"] = Bottleneck(cfg, in_f, out_f, False)
                in_f = out_f
            stages.append(blocks)
        self.res = nn.ModuleList(stages)
        self.cfg = cfg

    def forward(self, inputs, features):
        cfg = self.cfg
        r = cfg.first_subsampling_dl * cfg.density_parameter                   # build_models.py:186
        x = self.res1_input_conv(features)
        pts = inputs["points"]
        x = self.res1_simple_block(pts[0], pts[0], inputs["neighbors"][0], x, r)
        out = []
        for l, blocks in enumerate(self.res):
            if l > 0:          # strided block: radius of the level it pools FROM (resnet.py:383,399,415,431)
                x = blocks["strided_bottleneck"](inputs, l - 1, x, r * 2 ** (l - 1))
            for i in range(cfg.depth):
                x = blocks[f"bottleneck{i}"](inputs, l, x, r * 2 ** l)
            out.append(x)
        return out


class SegHead(nn.Module):
    """resnet_scene_segmentation_head with sep_head (seg_head.py:31-95): F_up only"""

    def __init__(self, cfg):
        super().__init__()
        f = cfg.first_features_dim
        dims = [f * 2 ** (l + 1) for l in range(5)]                            # backbone stage widths
        self.up_conv0 = Conv1x1(cfg, dims[4] + dims[3], 8 * f)
        self.up_conv1 = Conv1x1(cfg, 8 * f + dims[2], 4 * f)
        self.up_conv2 = Conv1x1(cfg, 4 * f + dims[1], 2 * f)
        self.up_conv3 = Conv1x1(cfg, 2 * f + dims[0], f)

    def forward(self, inputs, feats):
        ups = inputs["upsamples"]
        x = feats[4]
        f_up = []
        for j, conv in enumerate((self.up_conv0, self.up_conv1, self.up_conv2, self.up_conv3)):
            lvl = 4 - j                                                        # upsample level lvl -> lvl - 1
            x = closest_pool(x, ups[lvl][:, 0])
            x = conv(torch.cat((x, feats[lvl - 1]), 1))
            f_up.append(x)
        return list(reversed(f_up))                                            # [level 0 .. level 3]


class MultiscaleHead(nn.Module):
    """multiscale_head '||Ua-concat-latent' (head.py:338-425): latent_i = relu(bn(f_out_i W_i)) (mlps_by_ops '1mlp',
    head.py:268-273), nearest upsample of every latent to U0, concat, linear classifier"""

    def __init__(self, cfg, fdims):
        super().__init__()
        d = cfg.first_features_dim
        self.mlp = nn.ModuleList([Conv1x1(cfg, f, d) for f in fdims])
        self.linear = Linear(d * len(fdims), cfg.num_classes)                  # mlp_head.get_branch_head 'logits' (head.py:279-282)
        _xavier(self.linear)

    def forward(self, f_out, up_idx0):
        latents = [m(f) for m, f in zip(self.mlp, f_out)]
        cols = [latents[0]] + [closest_pool(lat, up_idx0[i]) for i, lat in enumerate(latents) if i > 0]
        return self.linear(torch.cat(cols, 1)), latents


def head_geometry(inputs, cfg: ConvNetConfig, with_labels=True):
    """everything the heads derive from coordinates (and labels) only:
    up_idx0[i]  nearest level-i point of every level-0 point within r_sample[i-1], else the shadow index (get_sample_idx kr=1,
                head.py:133-176; one stage apart = column 0 of the upsample rows)
    cls[i]      hard label of every level-i point (get_scene_label_infer reduction 'max', head.py:25-49)"""
    pts, lens = inputs["points"], inputs["batches_len"]
    nl = cfg.num_layers
    up_idx0, cls = [None] * nl, [None] * nl
    labels = inputs.get("point_labels")
    if labels is not None and with_labels:
        cls[0] = labels.to(torch.int32)
    for i in range(1, nl):
        if i == 1:
            up_idx0[i] = inputs["upsamples"][1][:, 0]
        else:
            up_idx0[i] = tf_ops.tf_batch_neighbors(pts[0], pts[i], lens[0], lens[i], cfg.r_sample[i - 1], limit=1)[:, 0]
        if labels is not None and with_labels:
            if i == 1:
                cls[i] = label_vote_idx(inputs["pools"][0], labels, cfg.num_classes, pts[0].shape[0])
            else:
                cls[i] = label_vote_radius(pts[i], pts[0], lens[i], lens[0], cfg.r_sample[i - 1], labels, cfg.num_classes)
    return {"up_idx0": up_idx0, "cls": cls}


class ConvNetSeg(nn.Module):
    """SceneSegModel (build_models.py:160-212) with arch_out = [multiscale, contrast]"""

    def __init__(self, cfg: Optional[ConvNetConfig] = None):
        super().__init__()
        self.cfg = cfg = cfg or ConvNetConfig()
        f = cfg.first_features_dim
        self.resnet_backbone = ResnetBackbone(cfg)
        self.resnet_scene_segmentation_head = SegHead(cfg)
        self.multiscale = MultiscaleHead(cfg, [f, 2 * f, 4 * f, 8 * f, 32 * f])     # up_list[4] = down_list[4] (build_models.py:203)

    def forward(self, inputs, geometry=None):
        """inputs: the pyramid dict of tf_pyramid.segmentation_inputs_radius.  -> logits (n0, classes), stage_list"""
        cfg = self.cfg
        geometry = geometry or head_geometry(inputs, cfg, self.training)
        feats = self.resnet_backbone(inputs, inputs["features"])
        f_up = self.resnet_scene_segmentation_head(inputs, feats)
        f_out = f_up + [feats[4]]
        logits, latents = self.multiscale(f_out, geometry["up_idx0"])
        pts = inputs["points"]
        stage_list = {"down": [{"p_out": pts[i], "f_out": feats[i]} for i in range(cfg.num_layers)],
                      "up": [{"p_out": pts[i], "f_out": f_out[i], "latent": latents[i]} for i in range(cfg.num_layers)],
                      "geometry": geometry, "inputs": inputs}
        flush_bn_counters()
        return logits, stage_list

    def decay_parameters(self):
        """the kernels the reference puts an L2 loss on (weight_decay > 0: every conv1d_1x1 / dense kernel; not fc_1, not biases,
        not batch norm: basic_operators.py:126-129,207-210,371-379, local_aggregation_operators.py:424-430)"""
        decay, rest = [], []
        for name, p in self.named_parameters():
            (decay if (name.endswith("weights.weight") or name.endswith("linear.weight")) else rest).append(p)
        return decay, rest


class ConvNetLoss(nn.Module):
    """loss_dict of build_loss (build_models.py:133-158) without the l2 term (that one lives in the optimiser's weight decay):
    stacked [cross entropy (multiscale 'seg'), softnn-up0 .. softnn-up4]"""

    def __init__(self, cfg: ConvNetConfig):
        super().__init__()
        self.cfg = cfg

    def forward(self, logits, labels, stage_list):
        cfg = self.cfg
        from .model import cross_entropy
        losses = [cross_entropy(logits, labels)]                                # calc_loss 'xen', mean over points
        if cfg.contrast:
            inputs, geo = stage_list["inputs"], stage_list["geometry"]
            t = cfg.contrast_temperature if cfg.contrast_temperature is not None else 1.0
            for i in range(cfg.num_layers):
                losses.append(tf_contrast_loss(stage_list["up"][i]["latent"], inputs["neighbors"][i], geo["cls"][i], t, cfg.contrast_weight))
        return torch.stack(losses)


# ------------------------------------------------------------------------------------------------------
# train step
# ------------------------------------------------------------------------------------------------------
def input_features(points, colors, in_features_dim=5):
    """'1-rgb-Z' (config/s3dis.py:69-71, datasets: ones, colours, height)"""
    ones = torch.ones_like(points[:, :1])
    if in_features_dim == 1:
        return ones
    if in_features_dim == 4:
        return torch.cat((ones, colors), 1)
    if in_features_dim == 5:
        return torch.cat((ones, colors, points[:, 2:3]), 1)
    raise NotImplementedError(in_features_dim)


# ------------------------------------------------------------------------------------------------
# TensorFlow checkpoint names -> this module's state_dict
# ------------------------------------------------------------------------------------------------
def tf_variable_to_state_dict(name):
    """('model/resnet_backbone/res2_bottleneck0/conv2/local_aggregation/fc_1/weights') ->
    ('resnet_backbone.res.1.bottleneck0.conv2.fc_1.weight', transpose?)  or None for variables this module does not hold
    (optimizer slots, the global step).  TF kernels are stored (in, out); nn.Linear weights (out, in): transpose = True."""
    import re
    bn_leaf = {"gamma": "weight", "beta": "bias", "moving_mean": "running_mean", "moving_variance": "running_var"}
    # the multi-scale head (heads/head.py:338-425): .../main/up{i}/mlp_0/{weights, batch_normalization/*}, .../main/linear/{weights, bias}
    m = re.search(r"(?:^|/)main/up(\d+)/mlp_0/(weights|batch_normalization/(\w+))$", name.split(":")[0])
    if m:
        if m.group(2) == "weights":
            return "multiscale.mlp.%s.weights.weight" % m.group(1), True
        return ("multiscale.mlp.%s.bn.%s" % (m.group(1), bn_leaf[m.group(3)]), False) if m.group(3) in bn_leaf else None
    m = re.search(r"(?:^|/)main/linear/(weights|bias)$", name.split(":")[0])
    if m:
        return ("multiscale.linear.weight", True) if m.group(1) == "weights" else ("multiscale.linear.bias", False)
    parts = [p for p in name.split(":")[0].split("/") if p not in ("model", "conv1d_1x1", "local_aggregation", "local_aggreagtion")]
    if not parts or parts[0] not in ("resnet_backbone", "resnet_scene_segmentation_head"):
        return None
    out = [parts[0]]
    for p_ in parts[1:-1]:
        m = re.fullmatch(r"res(\d+)_(strided_bottleneck|bottleneck\d+)", p_)
        out += ["res", str(int(m.group(1)) - 1), m.group(2)] if m else [p_]
    leaf, parent = parts[-1], (parts[-2] if len(parts) > 1 else "")
    is_bn = parent in ("bn", "pool_bn")
    if is_bn:
        key = {"gamma": "weight", "beta": "bias", "moving_mean": "running_mean", "moving_variance": "running_var"}.get(leaf)
        return None if key is None else (".".join(out + [key]), False)
    if parent == "fc_1":
        return {"weights": (".".join(out + ["weight"]), True), "biases": (".".join(out + ["bias"]), False)}.get(leaf)
    return {"weights": (".".join(out + ["weights", "weight"]), True), "biases": (".".join(out + ["weights", "bias"]), False)}.get(leaf)


def load_tf_variables(model, variables, strict=False):
    """variables: {TF variable name: array} (e.g. read from a reference checkpoint).  Copies every variable that maps onto
    `model` (ConvNetSeg); returns the list of TF names that were not used."""
    sd = model.state_dict()
    unused = []
    with torch.no_grad():
        for name, value in variables.items():
            m = tf_variable_to_state_dict(name)
            if m is None or m[0] not in sd:
                unused.append(name)
                continue
            t = torch.as_tensor(value, dtype=sd[m[0]].dtype)
            sd[m[0]].copy_(t.t() if m[1] else t)
    if strict and unused:
        raise KeyError("TF variables without a counterpart: %s" % unused[:8])
    return unused


class ConvNetTrainStep:
    """one training iteration of the TF trainer (utils/trainer.py): pyramid (the reference builds it on tf.data CPU workers) ->
    forward -> loss -> backward -> global-norm clip -> Momentum SGD."""

    def __init__(self, cfg: Optional[ConvNetConfig] = None, device="cuda", seed=0):
        from .tf_pyramid import PyramidConfig
        self.cfg = cfg or ConvNetConfig()
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.model = ConvNetSeg(self.cfg).to(self.device)
        self.criterion = ConvNetLoss(self.cfg)
        decay, rest = self.model.decay_parameters()
        self.opt = torch.optim.SGD([{"params": decay, "weight_decay": self.cfg.weight_decay}, {"params": rest, "weight_decay": 0.0}],
                                   lr=self.cfg.base_learning_rate, momentum=self.cfg.momentum, fused=self.device.type == "cuda")
        self.pcfg = PyramidConfig(self.cfg.num_layers, self.cfg.first_subsampling_dl, self.cfg.density_parameter,
                                  list(self.cfg.neighborhood_limits))
        self.model.train()
        self._side = None
        self._pending = None

    def build_inputs(self, batch):
        """batch: dict(points (n,3), colors (n,3), point_labels (n) int64, lens (b) int32) of device tensors"""
        from .tf_pyramid import segmentation_inputs_radius
        feats = input_features(batch["points"], batch["colors"], self.cfg.in_features_dim)
        return segmentation_inputs_radius(batch["points"], feats, batch["point_labels"], batch["lens"], self.pcfg)

    # ------------------------------------------------------------------------------------------
    # Pyramid look-ahead.  The reference builds the neighbour pyramid of the NEXT batches on tf.data CPU workers while the
    # GPU trains on the current one (tensorflow/datasets/base.py:75-118,767-842).  Here the pyramid of batch t+1 is built by
    # a worker thread on a side stream during step t: its host round trips (data-dependent sizes, the unordered_map order
    # replay) release the GIL while they wait, so they overlap with the main thread issuing the network's kernels.  Every
    # batch's pyramid is still built exactly once.
    # ------------------------------------------------------------------------------------------
    def prefetch_inputs(self, batch):
        import threading
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        ev_in = torch.cuda.Event()
        ev_in.record(main)                                   # the batch's H2D copies were enqueued on `main`
        box = {}

        def work():
            try:
                torch.cuda.set_device(self.device)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(ev_in)
                    inputs = self.build_inputs(batch)
                    ev = torch.cuda.Event()
                    ev.record(self._side)
                box["inputs"], box["ev"] = inputs, ev
            except BaseException as e:                       # re-raised by the consumer
                box["error"] = e
        th = threading.Thread(target=work, name="cb-pyramid-prefetch", daemon=True)
        th.start()
        self._pending = (batch, th, box)

    def drain_prefetch(self):
        """wait for (and drop) a pyramid that was prefetched but never consumed"""
        pend, self._pending = self._pending, None
        if pend is not None:
            pend[1].join()

    def _take_inputs(self, batch):
        pend, self._pending = self._pending, None
        if pend is None or pend[0] is not batch:
            if pend is not None:
                pend[1].join()                               # a prefetch nobody asked for: let it finish, drop it
            return self.build_inputs(batch)
        _, th, box = pend
        th.join()
        if "error" in box:
            raise box["error"]
        main = torch.cuda.current_stream(self.device)
        main.wait_event(box["ev"])
        for v in box["inputs"].values():                     # allocated on the side stream, consumed on `main`
            for t in (v if isinstance(v, (list, tuple)) else [v]):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(main)
        return box["inputs"]

    def step(self, batch, inputs=None, update=True, next_batch=None):
        """next_batch: optional batch whose pyramid is built on a worker thread / side stream during this step"""
        if inputs is None:
            inputs = self._take_inputs(batch)
        if next_batch is not None:
            self.prefetch_inputs(next_batch)
        self.opt.zero_grad(set_to_none=True)
        logits, stage_list = self.model(inputs)
        loss = self.criterion(logits, inputs["point_labels"], stage_list)
        loss.sum().backward()
        if update:
            if self.cfg.grad_norm:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.cfg.grad_norm, foreach=True)
            self.opt.step()
        return loss.detach()
