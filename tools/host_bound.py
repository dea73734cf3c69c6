"""Is the step host-launch-bound?  Host enqueue time vs GPU time per step, with/without geometry look-ahead."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic
dev = torch.device("cuda", 0)
ts = engine.TrainStep(model.CBLConfig(), dev)
hb = [engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000 + i)) for i in range(3)]
db = [engine.to_device(h, dev) for h in hb]
for i in range(4):
    ts.step(db[i % 3])
torch.cuda.synchronize()
for mode in ("plain", "lookahead"):
    if mode == "lookahead":
        ts.prefetch_geometry(db[0])
    torch.cuda.synchronize()
    host, tot = [], []
    for i in range(9):
        t0 = time.perf_counter()
        if mode == "plain":
            ts.step(db[i % 3])
        else:
            ts.step(db[i % 3], next_batch=db[(i + 1) % 3])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    print(f"{mode:10s} host enqueue {sorted(host)[4]:.1f} ms   step (enqueue + drain) {sorted(tot)[4]:.1f} ms")
# line-level host time of one forward+backward (to find host-blocking ops)
import collections, linecache
acc = collections.defaultdict(float); state = {"t": None, "k": None}
files = (model.__file__, engine.__file__)
def tracer(frame, event, arg):
    if frame.f_code.co_filename not in files:
        return None
    def local(frame, event, arg):
        now = time.perf_counter()
        if state["k"] is not None:
            acc[state["k"]] += now - state["t"]
        state["t"], state["k"] = time.perf_counter(), (frame.f_code.co_filename, frame.f_lineno)
        return local
    return local
torch.cuda.synchronize()
sys.settrace(tracer); ts.step(db[0]); sys.settrace(None); torch.cuda.synchronize()
for (fn, ln), v in sorted(acc.items(), key=lambda kv: -kv[1])[:10]:
    print(f"{1e3*v:8.2f} ms  {os.path.basename(fn)}:{ln}: {linecache.getline(fn, ln).strip()[:100]}")
# geometry alone
t0 = time.perf_counter()
for i in range(5):
    model.build_geometry(db[0]["points"], db[0]["offset"], db[0]["offset_host"], ts.cfg, True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"geometry: host {(t1 - t0) / 5 * 1e3:.1f} ms, total {(t2 - t0) / 5 * 1e3:.1f} ms")
