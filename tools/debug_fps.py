import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import pointops, synthetic, _lib as L
b = synthetic.make_batch(4, 40960, 2000)
p = torch.from_numpy(b["points"]).cuda(); o = torch.from_numpy(b["offset"]).cuda()
no = torch.tensor(np.cumsum([10240] * 4), dtype=torch.int32, device="cuda")
out = (C.c_ulonglong * 8)()
for mode in (1, 0, 2, 3, 4):
    L.lib().cb_fps_set_mode(mode, 8192)
    pointops.furthestsampling_known(p, o, no, 40960, 40960)
    L.lib().cb_debug_fps(C.c_int(1), out)
    pointops.furthestsampling_known(p, o, no, 40960, 40960); torch.cuda.synchronize()
    L.lib().cb_debug_fps(C.c_int(0), out)
    it = out[1]
    print("mode", mode, "(1 = single CTA; cluster CTAs x warps: 0 = 8x8, 2 = 4x4, 3 = 8x4, 4 = 4x8): iterations", it, "touched buckets total (4 scenes)", out[0],
          "per iter per scene", out[0] / 4 / max(it, 1), "max per warp-iteration", out[5])
    print("   cycles/iter (warp0)", out[2] / max(it, 1), " refresh cycles/iter (warp0)", out[3] / max(it, 1),
          " exchange (send + wait) cycles/iter", out[4] / max(it, 1))
L.lib().cb_fps_set_mode(0, 8192)
