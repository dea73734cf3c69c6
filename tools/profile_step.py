"""Kernel-time breakdown of one training step (torch.profiler / kineto; no nsys in this image).
    python tools/profile_step.py [--fused] [--top 45]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ts = engine.TrainStep(model.CBLConfig(fused=a.fused), dev)
    batch = engine.to_device(engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000)), dev)
    for _ in range(3):
        ts.step(batch)
    torch.cuda.synchronize()
    # coarse phases with CUDA events
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ts.opt.zero_grad(set_to_none=True)
    ev[0].record()
    levels = model.build_geometry(batch["points"], batch["offset"], batch["offset_host"], ts.cfg, True)
    ev[1].record()
    out, stages = ts.model(batch, levels)
    ev[2].record()
    loss = ts.criterion(out, batch["point_labels"], stages)
    ev[3].record()
    loss.sum().backward()
    ts.opt.step()
    ev[4].record()
    torch.cuda.synchronize()
    names = ["geometry (FPS + all KNN)", "forward (network)", "loss (CE + CBL)", "backward + SGD"]
    for i, nm in enumerate(names):
        print(f"{nm:28s} {ev[i].elapsed_time(ev[i + 1]):8.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.step(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.top, max_name_column_width=70))


if __name__ == "__main__":
    main()
