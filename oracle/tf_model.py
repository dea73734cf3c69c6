"""oracle/tf_model.py — TEST INFRASTRUCTURE.  Plain-torch restatement of the TF tree's AdaptiveWeight
aggregation (tensorflow/models/local_aggregation_operators.py:351-471, adapt.yaml config) and contrast_head
soft-NN loss (tensorflow/models/heads/head.py:180-195,641-662,725-807).  TensorFlow cannot be imported in the
build container.  Both functions are pinned under a stand-in runtime: tests/golden/tf_ops_ref.npz holds vectors from the
reference's own AdaptiveWeight and contrast_head source executed on a NumPy stand-in of the TF-1 API
(tests/golden/make_golden_tf_ops.py, tests/test_convnet_cpu.py); both also must agree with the second, independently
written restatement in oracle/tf_convnet_np.py."""
import torch

_EPS = 1e-12


def adaptive_weight(query_points, support_points, neighbors, features, fc_weight, fc_bias, radius):
    n0 = support_points.shape[0]
    idx = neighbors.long()
    shadow_f = torch.cat([features, torch.zeros_like(features[:1])], 0)                # :370
    nf = shadow_f[idx]                                                                   # :372
    shadow_p = torch.cat([support_points, torch.zeros_like(support_points[:1])], 0)     # :378
    rel = (shadow_p[idx] - query_points.unsqueeze(1)) / radius                          # :379-382
    w = rel @ fc_weight.t() + fc_bias                                                   # :426-430 (fc_num = 1)
    agg = (w * nf).sum(1)                                                               # :456-464 (shared_channels = 1)
    pad = idx.max()                                                                     # :466
    cnt = (idx < pad).float().sum(-1, keepdim=True) + 1e-5                              # :467-470
    return agg / cnt


def contrast_loss(features, neighbors, labels, temperature=1.0, weight=0.1):
    n = features.shape[0]
    idx = neighbors[:, 1:].long()                                                       # head.py:561-562
    valid = idx < n
    shadow_f = torch.cat([features, torch.zeros_like(features[:1])], 0)
    shadow_l = torch.cat([labels, labels.new_full((1,), -1)], 0)
    same = labels.unsqueeze(1) == shadow_l[idx]
    pos = same & valid                                                                  # :641-662
    neg = (~same) & valid
    pm = pos.any(1) & neg.any(1)
    if not pm.any():
        return features.sum() * 0.0
    f, nf, pos, neg = features[pm], shadow_f[idx][pm], pos[pm], neg[pm]
    dist = torch.sqrt(torch.clamp(((f.unsqueeze(1) - nf) ** 2).sum(-1), min=_EPS))      # :183-185
    d = -dist / temperature
    # numerical-stability shift (:750): any per-row constant cancels in pos / (pos + neg)
    d = d - torch.where(pos | neg, d, torch.full_like(d, -1e30)).max(-1, keepdim=True)[0]
    e = torch.exp(d)
    p = (e * pos).sum(-1)
    q = (e * neg).sum(-1)
    loss = -torch.log(p / (p + q) + _EPS)                                               # :760-771
    return loss.mean() * weight                                                          # :806
