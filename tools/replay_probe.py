"""Cost of the exact tie replay (knn.cu k_knn_replay): per-kernel device times of one cb_knn_query call (CUPTI via
torch.profiler) on a scene with a few exact ties (one duplicated point -> the queries around it are replayed)."""
import sys
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, '.')
from contrastboundary_b200 import pointops, synthetic

dev = torch.device('cuda', 0)
pointops.set_knn_cache(0)
for n in (40960, 65536, 1 << 18, 1 << 20):
    arr = synthetic.make_scene(n, 4242)[0].copy()
    arr[n // 2 + 1] = arr[n // 2]
    off = torch.tensor([n], dtype=torch.int32, device=dev)
    xyz = torch.from_numpy(arr).to(dev)
    for k in (16, 32, 64):
        for _ in range(3):
            pointops.knn_raw(k, xyz, xyz, off, off, False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(5):
                pointops.knn_raw(k, xyz, xyz, off, off, False)
            torch.cuda.synchronize()
        rows = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                rows.setdefault(e.name.split('(')[0][:48], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
        line = f"N={n:8d} K={k:3d}: " + "  ".join(f"{nm}={np.median(v):.1f}us" for nm, v in rows.items() if np.median(v) > 3)
        print(line, flush=True)
