import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from test_ptlayer_gpu import make_level, rel_err
from contrastboundary_b200 import model
for (c, k, n_list) in [(32, 8, [3000, 2000]), (64, 16, [1500, 900]), (128, 16, [700, 500]), (256, 16, [300, 200]), (512, 16, [90, 70])]:
    lv = make_level(n_list, k, 100 + c)
    layer = model.PointTransformerLayer(c, c, 8, k).cuda()
    cases.deterministic_init(layer, 3)
    layer.train(True)
    torch.manual_seed(1)
    x = torch.randn(lv.n, c, device="cuda"); gout = torch.randn(lv.n, c, device="cuda")
    res = {}
    for fused in (False, True):
        layer.fused = fused
        layer.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        # grads wrt q,k,v separately
        q, kk, v = layer.linear_q(xi), layer.linear_k(xi), layer.linear_v(xi)
        q.retain_grad(); kk.retain_grad(); v.retain_grad()
        if fused:
            from contrastboundary_b200 import ptlayer
            out = ptlayer.pt_attention(layer, lv, q, kk, v)
        else:
            import torch.nn.functional as F
            from contrastboundary_b200 import pointops
            from contrastboundary_b200.model import _bn_rows
            idx = lv.knn; n = lv.n; s = 8
            p_r = pointops.grouping(lv.p, idx) - lv.p.unsqueeze(1)
            x_kg, x_vg = pointops.grouping(kk, idx), pointops.grouping(v, idx)
            p_r = layer.linear_p[0](p_r); p_r = F.relu(_bn_rows(layer.linear_p[1], p_r)); p_r = layer.linear_p[3](p_r)
            w = x_kg - q.unsqueeze(1) + p_r
            w = F.relu(_bn_rows(layer.linear_w[0], w)); w = layer.linear_w[2](w)
            w = F.relu(_bn_rows(layer.linear_w[3], w)); w = layer.linear_w[5](w)
            w = F.softmax(w, dim=1)
            out = ((x_vg + p_r).view(n, k, s, c // s) * w.unsqueeze(2)).sum(1).view(n, c)
        out.backward(gout)
        torch.cuda.synchronize()
        d = {"out": out.detach(), "gq": q.grad, "gk": kk.grad, "gv": v.grad}
        for nme, p in layer.named_parameters():
            if nme.startswith("linear_p") or nme.startswith("linear_w"):
                d["g:" + nme] = p.grad.detach().clone()
        res[fused] = d
    print(f"--- c={c} k={k} n={lv.n}")
    for key in res[False]:
        a, b = res[True][key], res[False][key]
        print(f"   {key:28s} rel {rel_err(a, b):.2e}   |ref|max {float(b.abs().max()):.3e}")
