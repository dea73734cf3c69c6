"""developer tool (GPU): the tall-skinny linear layers of the network, tcgen05/TMEM kernel (umma_linear.cu) vs the mma.sync
kernel (tc_gemm.cu), L2 flushed, CUDA events.   python tools/bench_linear.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import _lib as L  # noqa: E402

SHAPES = [(163840, 32, 32), (163840, 32, 96), (40960, 64, 64), (40960, 64, 192), (40960, 32, 64), (10240, 128, 128), (10240, 128, 384),
          (10240, 64, 128), (163840, 32, 64), (1310720, 32, 8 * 4), (655360, 64, 8 * 8),
          # the deep levels (n = 2560 / 640 rows): cuBLAS territory so far (linear_ops.MIN_ROWS)
          (2560, 256, 256), (2560, 256, 768), (2560, 128, 256), (640, 512, 512), (640, 512, 1536), (640, 256, 512), (2560, 512, 256)]


def timed(fn, flush, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.mean(ts))


def main():
    lib = L.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print("      n    ci    co |  fwd: umma v3 us (frac)  umma v2  mma.sync | dgrad: umma v3 us (frac)  umma v2  mma.sync")
    for n, ci, co in SHAPES:
        x = torch.randn(n, ci, device="cuda")
        w = torch.randn(co, ci, device="cuda")
        b = torch.randn(co, device="cuda")
        y = torch.empty(n, co, device="cuda")
        g = torch.randn(n, co, device="cuda")
        dx = torch.empty(n, ci, device="cuda")
        by = 4 * n * (ci + co)
        res = []
        for call, args in (("cb_linear_forward", (n, ci, co, x, w, b, y)), ("cb_linear_dgrad", (n, ci, co, g, w, dx))):
            for umma, ver in ((1, 3), (1, 2), (0, 3)):
                lib.cb_linear_set_umma(C.c_int(umma))
                lib.cb_linear_set_umma_version(C.c_int(ver))
                res.append(timed(lambda: L.call(call, *args, L.stream()), flush))
        lib.cb_linear_set_umma(C.c_int(1))
        lib.cb_linear_set_umma_version(C.c_int(3))
        t_fwd = timed(lambda: torch.nn.functional.linear(x, w, b), flush)
        t_dg = timed(lambda: g.mm(w), flush)
        fr = [by / t / 1e3 / peak for t in res]
        print("%7d %5d %5d | %9.1f (%.2f) %8.1f (%.2f) %8.1f (%.2f) | %9.1f (%.2f) %8.1f (%.2f) %8.1f (%.2f)"
              % (n, ci, co, res[0], fr[0], res[1], fr[1], res[2], fr[2], res[3], fr[3], res[4], fr[4], res[5], fr[5])
              + "   | cuBLAS fwd %6.1f dgrad %6.1f" % (t_fwd, t_dg))


if __name__ == "__main__":
    main()
