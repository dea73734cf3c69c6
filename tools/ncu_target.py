"""Small fixed workload for ncu captures: fused KNN+gather (north-star shape) and a level-0 KNN query."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import fused, pointops, synthetic
n, k, c = 40960, 16, 256
xyz = torch.from_numpy(synthetic.make_scene(n, 4242)[0]).cuda()
off = torch.tensor([n], dtype=torch.int32, device="cuda")
feat = torch.randn(n, c, device="cuda")
grid = fused.grid_build(xyz, off, k)
out = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
for _ in range(3):
    fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out)
torch.cuda.synchronize()
pointops.set_knn_cache(0)
for _ in range(2):
    pointops.knn_raw(16, xyz, xyz, off, off, False)
torch.cuda.synchronize()
