"""Write-only / read-only / copy bandwidth of this GPU with library kernels (context for the roofline of cb_knn_gather,
whose HBM traffic is ~94% writes: 4NKC out vs 4NC in)."""
import numpy as np
import torch
dev = torch.device('cuda', 0)
st = torch.cuda.current_stream()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


for mb in (671, 2048, 8192):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, dtype=torch.float32, device=dev)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    by = n * 4
    t, tm = timed(lambda: x.zero_())
    print(f"{mb:5d} MiB  fill (write only): {by / t / 1e9:7.0f} GB/s (best {by / tm / 1e9:.0f})")
    t, tm = timed(lambda: torch.cuda.memset(x.data_ptr(), 0, by) if hasattr(torch.cuda, 'memset') else x.fill_(1.0))
    print(f"{mb:5d} MiB  fill_(1.0)       : {by / t / 1e9:7.0f} GB/s (best {by / tm / 1e9:.0f})")
    t, tm = timed(lambda: y.copy_(x))
    print(f"{mb:5d} MiB  copy (r+w bytes) : {2 * by / t / 1e9:7.0f} GB/s (best {2 * by / tm / 1e9:.0f});  written bytes only: {by / t / 1e9:.0f} GB/s")
    t, tm = timed(lambda: x.sum())
    print(f"{mb:5d} MiB  sum (read only)  : {by / t / 1e9:7.0f} GB/s (best {by / tm / 1e9:.0f})")
    del x, y
