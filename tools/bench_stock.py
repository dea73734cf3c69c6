"""Baseline B1: the reference network restated op by op (oracle/ref_model.py) on top of the
reference's OWN pointops CUDA kernels compiled unmodified for sm_100a (oracle/_ref/pointops_cuda.so), on the
same B200, same synthetic batch, same step (fwd + Loss + backward + SGD).  None of this repo's kernels runs here.
    python tools/bench_stock.py [--steps 5 --warmup 2 --scenes 4 --points 40960]
Prints one JSON line (points/s).  A developer / documentation tool, not the contract bench."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--points", type=int, default=40960)
    a = ap.parse_args()
    from oracle import gpu_pointops, ref_model
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = ref_model.RefSeg(gpu_pointops).to(dev)
    crit = ref_model.RefLoss(gpu_pointops).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.5, momentum=0.9, weight_decay=1e-4)
    model.train()
    batches = []
    for i in range(3):
        b = synthetic.make_batch(a.scenes, a.points, 5000 + i)
        batches.append(({k: torch.from_numpy(b[k]).to(dev) for k in ("points", "features", "offset")},
                        torch.from_numpy(b["point_labels"]).to(dev)))

    def step(i):
        inputs, target = batches[i % len(batches)]
        opt.zero_grad(set_to_none=True)
        out, up = model(inputs)
        loss = crit(out, target, up)
        loss.sum().backward()
        opt.step()
        return loss

    for w in range(a.warmup):
        step(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        loss = step(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"baseline": "B1 stock pointops CUDA build (reference kernels, sm_100a) + restated reference network",
                      "points_per_s": a.scenes * a.points / (ms * 1e-3), "ms_per_step": ms, "scenes": a.scenes,
                      "points_per_scene": a.points, "steps": a.steps, "warmup": a.warmup,
                      "loss": [round(float(x), 5) for x in loss]}))


if __name__ == "__main__":
    main()
