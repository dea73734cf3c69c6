import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from test_ptlayer_gpu import make_level, rel_err
from contrastboundary_b200 import model, pointops, ptlayer
from contrastboundary_b200.model import _bn_rows
for (c, k, n_list, seed) in [(64, 16, [1500, 900], 1), (512, 16, [90, 70], 1), (128, 16, [700, 500], 144)]:
    lv = make_level(n_list, k, 100 + c)
    layer = model.PointTransformerLayer(c, c, 8, k).cuda()
    cases.deterministic_init(layer, 3)
    layer.train(True)
    torch.manual_seed(seed)
    x = torch.randn(lv.n, c, device="cuda"); gout = torch.randn(lv.n, c, device="cuda")
    cs = c // 8; n = lv.n; idx = lv.knn
    # torch path with retained intermediates
    xi = x.clone().requires_grad_(True)
    q, kk, v = layer.linear_q(xi), layer.linear_k(xi), layer.linear_v(xi)
    p_r = pointops.grouping(lv.p, idx) - lv.p.unsqueeze(1)
    x_kg, x_vg = pointops.grouping(kk, idx), pointops.grouping(v, idx)
    h1 = layer.linear_p[0](p_r); y1 = _bn_rows(layer.linear_p[1], h1); g1 = F.relu(y1); pr = layer.linear_p[3](g1)
    w0 = x_kg - q.unsqueeze(1) + pr; w0.retain_grad()
    y2 = _bn_rows(layer.linear_w[0], w0); y2.retain_grad()
    u = F.relu(y2); w2 = layer.linear_w[2](u); w2.retain_grad()
    y3 = _bn_rows(layer.linear_w[3], w2); y3.retain_grad()
    vv = F.relu(y3); w4 = layer.linear_w[5](vv); w4.retain_grad()
    a = F.softmax(w4, dim=1)
    out = ((x_vg + pr).view(n, k, 8, cs) * a.unsqueeze(2)).sum(1).view(n, c)
    out.backward(gout)
    T = dict(dw0=w0.grad, dy2=y2.grad, dw2=w2.grad, dy3=y3.grad, dw4=w4.grad)
    # fused path
    layer.zero_grad(set_to_none=True)
    xi2 = x.clone().requires_grad_(True)
    q2, k2, v2 = layer.linear_q(xi2), layer.linear_k(xi2), layer.linear_v(xi2)
    q2.retain_grad()
    out2 = ptlayer.pt_attention(layer, lv, q2, k2, v2)
    out2.backward(gout)
    torch.cuda.synchronize()
    scratch, gbuf, bnbuf = ptlayer.PtAttentionFn.debug_last
    nd = 2 * cs + 2 * c + 8
    dbl = scratch[:2 * nd].view(torch.float64) if False else scratch[:2 * nd].contiguous().view(torch.float64)
    sums3, sums2, sums1 = dbl[:2 * cs], dbl[2 * cs:2 * cs + 2 * c], dbl[2 * cs + 2 * c:2 * cs + 2 * c + 6]
    o = 2 * nd
    coef3 = scratch[o:o + 3 * cs]; o += 3 * cs
    coef2 = scratch[o:o + 3 * c]; o += 3 * c
    coef1 = scratch[o:o + 9]; o += 16
    D = scratch[o:o + n * k * cs].view(n, k, cs); o += n * k * cs
    dy1 = scratch[o:o + n * k * 3].view(n, k, 3)
    print(f"--- c={c} n={n}  out rel {rel_err(out2.detach(), out.detach()):.1e}  gq rel {rel_err(q2.grad, q.grad if q.grad is not None else torch.autograd.grad(out, q, gout, retain_graph=True)[0]) if False else 0}")
    print("   dy3 (D)   rel", f"{rel_err(D, T['dy3']):.2e}")
    print("   S3a rel", f"{rel_err(sums3[:cs].float(), T['dy3'].sum((0,1))):.2e}", " S2a rel", f"{rel_err(sums2[:c].float(), T['dy2'].sum((0,1))):.2e}",
          " |S2a|max", float(T['dy2'].sum((0,1)).abs().max()), " sum|dy2| max", float(T['dy2'].abs().sum((0,1)).max()))
    bn3 = bnbuf[24 + 4 * c:24 + 4 * c + 4 * cs]
    xh3 = (w2.detach() - bn3[2 * cs:3 * cs]) * bn3[3 * cs:4 * cs]
    dw2_m = coef3[:cs] * (D - coef3[cs:2 * cs] - xh3 * coef3[2 * cs:])
    print("   dw2 rel", f"{rel_err(dw2_m, T['dw2']):.2e}")
    # expected dw0 from torch dy2 with my coef2
    bn2 = bnbuf[24:24 + 4 * c]
    xh2 = (w0.detach() - bn2[2 * c:3 * c]) * bn2[3 * c:4 * c]
    dw0_m = coef2[:c] * (T['dy2'] - coef2[c:2 * c] - xh2 * coef2[2 * c:])
    print("   dw0 (torch dy2 + my coef2) rel", f"{rel_err(dw0_m, T['dw0']):.2e}", "  gq(mine) vs -sum_k dw0_t rel", f"{rel_err(q2.grad, -T['dw0'].sum(1)):.2e}")
    du_t = T['dw2'] @ layer.linear_w[2].weight   # (n,k,c)
    dy2_chk = du_t * (y2.detach() > 0)
    print("   torch consistency dy2 rel", f"{rel_err(dy2_chk, T['dy2']):.2e}")
    bad = ((q2.grad + T['dw0'].sum(1)).abs() > 1e-3 * T['dw0'].sum(1).abs().max())
    print("   bad gq entries:", int(bad.sum()), "of", bad.numel(), " bad rows:", int(bad.any(1).sum()), " bad cols:", bad.any(0).nonzero().flatten()[:20].tolist())
