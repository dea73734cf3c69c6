"""Small fixed workload for ncu captures of the tcgen05 / TMEM linear kernel (k_umma_linear2)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import _lib as L
for n, ci, co in ((655360, 64, 64), (40960, 64, 192)):
    x = torch.randn(n, ci, device="cuda"); w = torch.randn(co, ci, device="cuda"); b = torch.randn(co, device="cuda")
    y = torch.empty(n, co, device="cuda")
    for _ in range(3):
        L.call("cb_linear_forward", n, ci, co, x, w, b, y, L.stream())
    torch.cuda.synchronize()
