"""Small fixed workload for ncu captures of the tensor-core kernels: one fused PointTransformer layer (forward + backward,
level-1 shape n = 40960, k = 16, c = 64) and one tall-skinny linear layer (163840 x 32 -> 96)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import linear_ops, model, pointops, ptlayer, synthetic
torch.manual_seed(0)
b = synthetic.make_batch(1, [40960], 77)
lv = model.Level()
lv.p = torch.from_numpy(b["points"]).cuda(); lv.o = torch.from_numpy(b["offset"]).cuda(); lv.n = lv.p.shape[0]
lv.knn, _ = pointops.knn_raw(16, lv.p, lv.p, lv.o, lv.o, True)
lv.rel, lv.rel_mom = ptlayer.pt_rel(lv.p, lv.knn)
layer = model.PointTransformerLayer(64, 64, 8, 16).cuda().train()
x = torch.randn(lv.n, 64, device="cuda", requires_grad=True)
g = torch.randn(lv.n, 64, device="cuda")
xl = torch.randn(163840, 32, device="cuda", requires_grad=True)
w = (torch.randn(96, 32, device="cuda") / 6).requires_grad_(True)
gl = torch.randn(163840, 96, device="cuda")
for _ in range(3):
    layer(lv, x).backward(g)
    linear_ops.fast_linear(xl, w, None).backward(gl)
torch.cuda.synchronize()
