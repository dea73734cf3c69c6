"""cProfile of the stream-mode training step (host side): which Python functions the ~27 ms of enqueue time go to."""
import cProfile
import os
import pstats
import sys

import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrastboundary_b200 import engine, model, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
ts = engine.TrainStep(model.CBLConfig(), dev)
db = [engine.to_device(engine.host_batch_from_numpy(synthetic.make_batch(4, 40960, 5000 + i)), dev) for i in range(2)]
for i in range(4):
    ts.step(db[i % 2], next_batch=db[(i + 1) % 2])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    ts.step(db[i % 2], next_batch=db[(i + 1) % 2])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
