"""CPU suite: the restated reference model (oracle/ref_model.py on oracle/cpu_pointops.py) against
the golden vectors produced by the REAL reference model code (tests/golden/make_golden_model.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from oracle import cpu_pointops, ref_model  # noqa: E402


def run_oracle_model():
    torch.manual_seed(0)
    model = ref_model.RefSeg(cpu_pointops)
    crit = ref_model.RefLoss(cpu_pointops)
    cases.deterministic_init(model, 0)
    model.train()
    b = cases.model_batch()
    inputs = {k: torch.from_numpy(b[k]) for k in ("points", "features", "offset")}
    target = torch.from_numpy(b["point_labels"])
    out, up = model(inputs)
    loss = crit(out, target, up)
    loss.sum().backward()
    return model, out, loss, up


def test_oracle_model_matches_real_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_ref.npz"))
    model, out, loss, up = run_oracle_model()
    assert np.allclose(out.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-5)
    assert np.allclose(loss.detach().numpy(), g["loss"], rtol=1e-6, atol=1e-7)
    norms = json.loads(bytes(g["grad_norms_json"]).decode())
    params = dict(model.named_parameters())
    assert set(norms) == {n for n, p in params.items() if p.grad is not None}
    for name, ref in norms.items():
        if cases.grad_is_analytically_zero(name):
            assert float(params[name].grad.norm()) < 1e-3 and ref < 1e-3
            continue
        assert abs(float(params[name].grad.norm()) - ref) <= 1e-4 * max(ref, 1e-6) + 1e-7, name
    for name in cases.GOLDEN_GRADS:
        # atol: torch-CPU reductions change their summation order with the thread count of the process (flaky at 1e-6)
        assert np.allclose(params[name].grad.numpy(), g["grad/" + name], rtol=1e-4, atol=2e-5), name
    for i in range(5):
        assert np.allclose(up[i]["latent"].detach().numpy()[:64], g[f"latent/{i}"], rtol=1e-5, atol=1e-5)
