import os, sys, cProfile, pstats, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic
dev = torch.device("cuda", 0)
cfg = model.CBLConfig()
b = engine.to_device(engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000)), dev)
for _ in range(3):
    model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
model.build_geometry(b["points"], b["offset"], b["offset_host"], cfg, True)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
