"""Fused contrastive-boundary loss of one stage (reference: ContrastHead.point_contrast,
pytorch/model/heads.py:185-246) over libcbops' cb_cbl_classes / cb_cbl_forward / cb_cbl_backward."""
import torch
from torch.autograd import Function

from . import _lib as L


class CblStageFn(Function):
    @staticmethod
    def forward(ctx, feat, idx, cls, temperature, weight):
        feat = feat.contiguous()
        m, d = feat.shape
        k = idx.shape[1]
        sums = torch.zeros(2, dtype=torch.float32, device=feat.device)
        L.call("cb_cbl_forward", m, k, d, feat, idx, cls, float(temperature), sums, L.stream())
        ctx.save_for_backward(feat, idx, cls, sums)
        ctx.temperature, ctx.weight = float(temperature), float(weight)
        # mean over boundary points; 0 when the stage has none (heads.py:222-233, kept on the device)
        return sums[0] / sums[1].clamp(min=1.0) * weight

    @staticmethod
    def backward(ctx, g):
        feat, idx, cls, sums = ctx.saved_tensors
        m, d = feat.shape
        scale = (g * ctx.weight / sums[1].clamp(min=1.0)).reshape(1).float().contiguous()
        gfeat = torch.zeros_like(feat)
        L.call("cb_cbl_backward", m, idx.shape[1], d, feat, idx, cls, ctx.temperature, scale, gfeat, L.stream())
        return gfeat, None, None, None, None


def point_classes(l, levels, target, ncls):
    lv = levels[l]
    m = lv.n
    cls = torch.empty(m, dtype=torch.int32, device=target.device)
    if l == 0:
        L.call("cb_cbl_classes", m, 0, ncls, None, target, cls, L.stream())
    else:
        L.call("cb_cbl_classes", m, lv.label_idx.shape[1], ncls, lv.label_idx, target, cls, L.stream())
    return cls


def cbl_stage_loss(head, l, levels, latent, target):
    cc = head.cfg.contrast
    cls = point_classes(l, levels, target.contiguous(), head.cfg.classes)
    t = cc.temperature if cc.temperature is not None else 1.0
    return CblStageFn.apply(latent, levels[l].cbl_idx, cls, t, cc.weight)
