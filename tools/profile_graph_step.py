"""Per-kernel device time of the benchmark step AS THE BENCH RUNS IT: engine.GraphTrainStep replays (network graph on the main
stream, geometry graph of the next batch on the side stream), warm, via kineto (CUPTI sees the kernels inside graph launches).
    python tools/profile_graph_step.py [--top 40] [--steps 4]"""
import argparse
import json
import os
import sys
import tempfile
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ts = engine.GraphTrainStep(model.CBLConfig(fused=True), dev, seed=0)
    pool = [engine.to_device(engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000 + i)), dev) for i in range(2)]
    for w in range(8):
        ts.step(pool[w % 2], next_batch=pool[(w + 1) % 2])
    torch.cuda.synchronize()
    assert not ts.graph_error, ts.graph_error
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(a.steps):
        ts.step(pool[s % 2], next_batch=pool[(s + 1) % 2])
    t1.record()
    torch.cuda.synchronize()
    print(f"un-profiled: {t0.elapsed_time(t1) / a.steps:.2f} ms/step")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for s in range(a.steps):
            ts.step(pool[s % 2], next_batch=pool[(s + 1) % 2])
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "cb_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    per_stream = defaultdict(lambda: defaultdict(lambda: [0.0, 0]))
    span = defaultdict(lambda: [1e30, 0.0])
    for e in ev:
        if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e:
            st = e.get("args", {}).get("stream", -1)
            nm = e["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0].replace("void ", "").replace("at::native::", "")[:90]
            per_stream[st][nm][0] += e["dur"]
            per_stream[st][nm][1] += 1
            span[st][0] = min(span[st][0], e["ts"])
            span[st][1] = max(span[st][1], e["ts"] + e["dur"])
    for st, ks in sorted(per_stream.items(), key=lambda kv: -sum(v[0] for v in kv[1].values())):
        tot = sum(v[0] for v in ks.values())
        n = sum(v[1] for v in ks.values())
        print(f"\n== stream {st}: busy {tot / a.steps / 1e3:.2f} ms/step in {n // a.steps} launches/step "
              f"(span {(span[st][1] - span[st][0]) / a.steps / 1e3:.2f} ms/step)")
        for nm, (us, cnt) in sorted(ks.items(), key=lambda kv: -kv[1][0])[:a.top]:
            print(f"  {us / a.steps:9.1f} us/step  {cnt / a.steps:6.1f} launches  {us / cnt:8.1f} us avg   {nm}")


if __name__ == "__main__":
    main()
