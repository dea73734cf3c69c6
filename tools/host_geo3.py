import os, sys, time, collections, linecache
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic
dev = torch.device("cuda", 0)
cfg = model.CBLConfig()
b = engine.to_device(engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000)), dev)
for _ in range(3):
    model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
torch.cuda.synchronize()
acc = collections.defaultdict(float); state = {"t": None, "ln": None}
code = model.build_geometry.__code__
def tracer(frame, event, arg):
    if frame.f_code is not code:
        return None
    def local(frame, event, arg):
        now = time.perf_counter()
        if state["ln"] is not None:
            acc[state["ln"]] += now - state["t"]
        state["t"], state["ln"] = time.perf_counter(), frame.f_lineno
        return local
    return local
sys.settrace(tracer)
model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
sys.settrace(None)
torch.cuda.synchronize()
for ln, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]:
    print(f"{1e3*v:8.2f} ms  line {ln}: {linecache.getline(code.co_filename, ln).strip()[:110]}")
