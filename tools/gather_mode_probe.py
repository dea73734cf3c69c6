"""cb_knn_gather_grid kernel time by copy-path variant (cb_knn_gather_set_mode) and row width."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from contrastboundary_b200 import _lib as L, fused, synthetic
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


for n in (40960, 262144):
    xyz = torch.from_numpy(synthetic.make_scene(n, 4242)[0]).to(dev)
    off = torch.tensor([n], dtype=torch.int32, device=dev)
    for k in (16, 32, 64):
        grid = fused.grid_build(xyz, off, k)
        for c in (32, 64, 128, 256):
            if n * k * c * 4 > (6 << 30):
                continue
            feat = torch.randn(n, c, device=dev)
            out = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
            line = f"N={n:7d} K={k:2d} C={c:3d}:"
            for mode in (3, 5, 6, 7):
                L.lib().cb_knn_gather_set_mode(mode)
                us = timed(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out))
                line += f"  mode{mode} {us:8.1f}"
            print(line, flush=True)
            del feat, out
L.lib().cb_knn_gather_set_mode(-1)
