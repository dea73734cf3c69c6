import os, sys, time, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic, pointops, ptlayer
dev = torch.device("cuda", 0)
cfg = model.CBLConfig()
hb = engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000))
b = engine.to_device(hb, dev)
for _ in range(3):
    model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
torch.cuda.synchronize()
acc = collections.defaultdict(float); cnt = collections.Counter()
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[name] += time.perf_counter() - t0; cnt[name] += 1
        return r
    setattr(mod, name, g)
wrap(pointops, "knn_raw"); wrap(pointops, "furthestsampling_known"); wrap(ptlayer, "pt_rel")
t0 = time.perf_counter()
model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host {1e3*(t1-t0):.2f} ms  total {1e3*(t2-t0):.2f} ms")
for k, v in acc.items():
    print(f"  {k:28s} calls {cnt[k]:3d} host {1e3*v:.2f} ms")
# per knn call detail
import ctypes as C
from contrastboundary_b200 import _lib as L
p, o = b["points"], b["offset"]
for k in (8, 36):
    torch.cuda.synchronize(); t0 = time.perf_counter(); pointops.knn_raw(k, p, p, o, o, True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"knn K={k}: host {1e3*(t1-t0):.3f} ms total {1e3*(t2-t0):.3f} ms")
