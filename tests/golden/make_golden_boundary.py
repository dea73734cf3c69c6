"""Golden vectors for get_boundary_mask from the REAL reference function (pytorch/model/basic_operators.py:69-97),
imported in the build container (torch CPU).    python tests/golden/make_golden_boundary.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pytorch"
sys.path.insert(0, REF)
# the module imports the compiled pointops at import time; the function under test does not use it
stub = types.ModuleType("lib.pointops.functions.pointops")
for name in ("lib", "lib.pointops", "lib.pointops.functions"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["lib.pointops.functions.pointops"] = stub
sys.modules["lib.pointops.functions"].pointops = stub
from model.basic_operators import get_boundary_mask  # noqa: E402

rng = np.random.default_rng(11)
out = {}
for case, (n, kr) in enumerate([(500, 16), (2000, 32), (64, 5)]):
    labels = rng.integers(0, 6, n).astype(np.int64)
    labels[rng.random(n) < 0.1] = -1                       # invalid labels
    region = (np.arange(n) * 6 // n)
    labels = np.where(labels >= 0, region, labels)         # spatially coherent labels -> few boundaries
    idx = np.clip(np.arange(n)[:, None] + rng.integers(-40, 41, (n, kr)), 0, n - 1).astype(np.int32)
    valid = labels >= 0
    with contextlib.redirect_stdout(io.StringIO()):        # the reference prints shapes
        b, p = get_boundary_mask(torch.from_numpy(labels), neighbor_idx=torch.from_numpy(idx), get_plain=True)
        bv, pv = get_boundary_mask(torch.from_numpy(labels), neighbor_idx=torch.from_numpy(idx),
                                   valid_mask=torch.from_numpy(valid), get_plain=True)
        c = get_boundary_mask(torch.from_numpy(labels), neighbor_idx=torch.from_numpy(idx),
                              valid_mask=torch.from_numpy(valid), get_cnt=True)
    out.update({f"{case}/labels": labels, f"{case}/idx": idx, f"{case}/bound": b.numpy(), f"{case}/plain": p.numpy(),
                f"{case}/bound_valid": bv.numpy(), f"{case}/plain_valid": pv.numpy(), f"{case}/cnt": c.numpy()})
np.savez_compressed(os.path.join(HERE, "boundary_ref.npz"), **out)
print("wrote", os.path.join(HERE, "boundary_ref.npz"), {k: v.shape for k, v in out.items() if k.startswith("0/")})
