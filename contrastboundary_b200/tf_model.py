"""Host-side mirror of the TF tree's hot consumers of the radius-neighbour pyramid, on torch + libcbops:

  adaptive_weight(...)     the 'ConvNet' local aggregation  (tensorflow/models/local_aggregation_operators.py:316-500,
                           adapt.yaml config: input dp, one FC, shared_channels 1, mean, no softmax)
  tf_contrast_loss(...)    contrast_head with softnn / l2 / label sampling (tensorflow/models/heads/head.py:462-807)

TensorFlow itself is not part of this stack; BatchNorm / ReLU / 1x1 convs around these ops are plain torch."""
import torch
from torch.autograd import Function

from . import _lib as L


class AdaptiveWeightFn(Function):
    @staticmethod
    def forward(ctx, query_points, support_points, neighbors, features, fc_weight, fc_bias, radius):
        L.require_cuda(query_points, support_points, neighbors, features, fc_weight, fc_bias)
        q, s, f = query_points.contiguous(), support_points.contiguous(), features.contiguous()
        idx = neighbors.contiguous().int()
        n, k = idx.shape
        n0, c = f.shape
        out = torch.empty((n, c), dtype=torch.float32, device=f.device)
        pad = torch.zeros(1, dtype=torch.int32, device=f.device)
        w, b = fc_weight.contiguous(), fc_bias.contiguous()
        L.call("cb_adaptive_weight_forward", n, k, c, n0, q, s, idx, f, w, b, float(radius), pad, out, L.stream())
        ctx.save_for_backward(q, s, idx, f, w, b, pad)
        ctx.radius = float(radius)
        return out

    @staticmethod
    def backward(ctx, g):
        q, s, idx, f, w, b, pad = ctx.saved_tensors
        n, k = idx.shape
        n0, c = f.shape
        gf, gw, gb = torch.zeros_like(f), torch.zeros_like(w), torch.zeros_like(b)
        L.call("cb_adaptive_weight_backward", n, k, c, n0, q, s, idx, f, w, b, ctx.radius, pad, g.contiguous(), gf, gw, gb,
               L.stream())
        return None, None, None, gf, gw, gb, None


def adaptive_weight(query_points, support_points, neighbors, features, fc_weight, fc_bias, radius):
    """aggregated features [n_points, fdim] (before pool_bn / activation / output_conv of the reference op).
    fc_weight (fdim, 3), fc_bias (fdim): the single FC on dp ('fc_1', local_aggregation_operators.py:426-430)."""
    return AdaptiveWeightFn.apply(query_points, support_points, neighbors, features, fc_weight, fc_bias, radius)


class _TfCblFn(Function):
    @staticmethod
    def forward(ctx, feat, idx, cls, n_valid, temperature, weight):
        feat = feat.contiguous()
        m, d = feat.shape
        sums = torch.zeros(2, dtype=torch.float32, device=feat.device)
        L.call("cb_cbl_forward_ex", m, idx.shape[1], d, feat, idx, cls, float(temperature), sums, int(n_valid), 1, L.stream())
        ctx.save_for_backward(feat, idx, cls, sums)
        ctx.cfg = (int(n_valid), float(temperature), float(weight))
        return sums[0] / sums[1].clamp(min=1.0) * weight

    @staticmethod
    def backward(ctx, g):
        feat, idx, cls, sums = ctx.saved_tensors
        n_valid, t, w = ctx.cfg
        scale = (g * w / sums[1].clamp(min=1.0)).reshape(1).float().contiguous()
        gfeat = torch.zeros_like(feat)
        L.call("cb_cbl_backward_ex", feat.shape[0], idx.shape[1], feat.shape[1], feat, idx, cls, t, scale, gfeat, n_valid, 1,
               L.stream())
        return gfeat, None, None, None, None, None


def tf_contrast_loss(features, neighbors, labels, temperature=1.0, weight=0.1):
    """CBL at one stage, TF flavour.  features (n,d) d in {32,64,72}; neighbors (n,K) int32 radius rows whose column 0 is
    the point itself (dropped, head.py:561-562) and whose shadow entries equal n; labels (n) integer hard labels."""
    n = features.shape[0]
    return _TfCblFn.apply(features, neighbors.contiguous().int(), labels.contiguous().int(), n, temperature, weight)
