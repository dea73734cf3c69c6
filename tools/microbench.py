"""Quick device-timed microbench of the hot operators (CUDA events, L2 flushed between iterations).
    python tools/microbench.py [--ref]      (--ref also times the reference's stock kernels from oracle/_ref)
Not the contract bench (that is bench.py); this is the developer loop."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import fused, pointops, synthetic  # noqa: E402

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = "cuda"
    pointops.set_knn_cache(0)
    b = synthetic.make_batch(4, 40960, 2000)
    p0 = torch.from_numpy(b["points"]).to(dev)
    o0 = torch.from_numpy(b["offset"]).to(dev)
    ref = None
    if args.ref:
        import oracle
        ref = oracle.ref_pointops_cuda()
    print("== level-0 self KNN, n=163840 (4 x 40960)")
    import ctypes as C
    from contrastboundary_b200 import _lib as L
    L.lib().cb_knn_set_occupancy.restype = C.c_float
    for k in ((8, 16, 36) if args.only in ("", "knn") else ()):
        for f in (0.2, 0.3, 0.6, 0.45):
            L.lib().cb_knn_set_occupancy(C.c_float(f))
            med, mn = timeit(lambda: pointops.knn_raw(k, p0, p0, o0, o0, True))
            print(f"     occupancy factor {f}: {med:.1f} us")
        line = f"  ours K={k:3d}: {med:9.1f} us (min {mn:.1f})"
        if ref is not None and k in (16,):
            idx = torch.zeros((p0.shape[0], k), dtype=torch.int32, device=dev)
            d2 = torch.zeros((p0.shape[0], k), dtype=torch.float32, device=dev)
            rm, _ = timeit(lambda: ref.knnquery_cuda(p0.shape[0], k, p0, p0, o0, o0, idx, d2), iters=3, warmup=1)
            line += f"   reference stock kernel: {rm:9.1f} us  ({rm / med:.1f}x)"
        print(line)
    print("== FPS chain 163840 -> 40960 -> 10240 -> 2560 -> 640")
    p, o = p0, o0
    lens = [40960] * 4
    for lvl in (range(4) if args.only in ("", "fps") else ()):
        nl = [x // 4 for x in lens]
        no = torch.tensor(np.cumsum(nl), dtype=torch.int32, device=dev)
        med, mn = timeit(lambda: pointops.furthestsampling_known(p, o, no, max(lens), sum(nl)), iters=5, warmup=2)
        line = f"  ours n/scene={lens[0]:6d}: {med:9.1f} us"
        if ref is not None:
            idx = torch.zeros(sum(nl), dtype=torch.int32, device=dev)

            def run_ref():
                tmp = torch.full((p.shape[0],), 1e10, dtype=torch.float32, device=dev)
                ref.furthestsampling_cuda(4, max(lens), p, o, no, tmp, idx)
            rm, _ = timeit(run_ref, iters=2, warmup=1)
            line += f"   reference: {rm:9.1f} us ({rm / med:.1f}x)"
        print(line)
        idx = pointops.furthestsampling_known(p, o, no, max(lens), sum(nl))
        p, o, lens = p[idx.long()].contiguous(), no, nl
    print("== fused KNN+gather (north star): N=40960 K=16 C=256, single scene")
    for (n, k, c) in ([(40960, 16, 256), (40960, 16, 64), (1 << 16, 16, 256), (1 << 18, 16, 64)] if args.only in ("", "gather") else []):
        xyz = torch.from_numpy(synthetic.make_scene(n, 4242)[0]).to(dev)
        off = torch.tensor([n], dtype=torch.int32, device=dev)
        feat = torch.randn(n, c, device=dev)
        med, mn = timeit(lambda: fused.knn_gather(k, xyz, xyz, feat, off, off))
        by = 12 * n + 4 * n * c + 8 * n * k + 4 * n * k * c
        print(f"  N={n:7d} K={k} C={c:3d}: {med:9.1f} us (min {mn:.1f})  alg {by / 1e6:.1f} MB -> {by / med / 1e3:.0f} GB/s"
              f" ({by / med / 1e3 / 6569.6 * 100:.0f}% of measured HBM)")
        grid = fused.grid_build(xyz, off, k)
        outb = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
        from contrastboundary_b200 import _lib as L
        for ns in (0, 8, 16, 32, 64):
            L.lib().cb_knn_gather_set_spin_ns(ns); L.lib().cb_knn_gather_set_mode(3)
            medk, mnk = timeit(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, outb), iters=20)
            print(f"      kernel only, ws=3, queue-full back-off {ns:3d} ns: {medk:8.1f} us (min {mnk:.1f})")
        L.lib().cb_knn_gather_set_spin_ns(0)
        for cb, ws in ((8192, 3), (16384, 3), (8192, 5), (8192, 1), (8192, 6), (8192, 7)):     # measured: 150-158 / 159 / 155 / 160 us
            L.lib().cb_knn_gather_set_chunk_bytes(cb)
            L.lib().cb_knn_gather_set_mode(ws)
            medk, mnk = timeit(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, outb))
            print(f"      kernel only (grid prebuilt), ws={ws} chunk {cb:5d} B: {medk:8.1f} us (min {mnk:.1f}) -> {by / medk / 1e3:.0f} GB/s ({by / medk / 1e3 / 6569.6 * 100:.0f}%)")
        L.lib().cb_knn_gather_set_chunk_bytes(8192); L.lib().cb_knn_gather_set_mode(-1)
        xyz2 = xyz.clone()          # a distinct query tensor: queries are processed in ORIGINAL order (linear output addresses)
        medq, mnq = timeit(lambda: fused.knn_gather_grid(grid, k, xyz, xyz2, feat, off, off, outb))
        print(f"      kernel only, queries in original order (streamed output):  {medq:8.1f} us (min {mnq:.1f}) -> {by / medq / 1e3:.0f} GB/s ({by / medq / 1e3 / 6569.6 * 100:.0f}%)")
        med2, _ = timeit(lambda: pointops.knn_raw(k, xyz, xyz, off, off, False))
        idx, _ = pointops.knn_raw(k, xyz, xyz, off, off, False)
        med3, _ = timeit(lambda: pointops.grouping(feat, idx))
        print(f"      unfused: knn {med2:.1f} us + grouping {med3:.1f} us")


if __name__ == "__main__":
    main()
