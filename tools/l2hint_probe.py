import torch, numpy as np, sys
sys.path.insert(0, '.')
from contrastboundary_b200 import _lib as L, fused, synthetic
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush_r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
def timed(fn, iters=10, cold=True):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        if cold: flush.zero_(); flush_r.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
for n, k, c in [(40960, 16, 256), (65536, 16, 256), (65536, 32, 256), (131072, 16, 256), (131072, 16, 64), (262144, 16, 256), (1 << 20, 16, 256)]:
    xyz = torch.from_numpy(synthetic.make_scene(n, 4242)[0]).to(dev); off = torch.tensor([n], dtype=torch.int32, device=dev)
    feat = torch.randn(n, c, device=dev)
    grid = fused.grid_build(xyz, off, k); out = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
    nrep = int(grid[:32].view(torch.int32)[1])
    by = 12 * n + 4 * n * c + 8 * n * k + 4 * n * k * c
    ref = None
    line = f"N={n} K={k} C={c} replays={nrep}:"
    for hint in (0, 1, 2, 3):
        assert L.lib().cb_knn_gather_set_l2_hint(hint) == hint
        tc = timed(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out))
        tw = timed(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out), cold=False)
        line += f"  hint{hint}: cold {tc:.1f} us ({by/tc/1e3/6569.6:.2f}) warm {tw:.1f}"
        if ref is None: ref = [o.clone() for o in out]
        else: assert all(torch.equal(a, b) for a, b in zip(ref, out))
    print(line, flush=True)
    del grid, out, feat, ref; torch.cuda.empty_cache()
L.lib().cb_knn_gather_set_l2_hint(3)
