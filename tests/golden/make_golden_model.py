"""Generate tests/golden/model_ref.npz by running the REAL reference model code
(/root/reference/pytorch/model/*.py, imported unmodified) on CPU in the build container, with
`lib.pointops.functions.pointops` provided by oracle/cpu_pointops.py (the C restatement of the
reference kernels, itself pinned bit-for-bit by tests/golden/pointops_ref_gpu.npz).

    python tests/golden/make_golden_model.py

Stores logits, the loss vector [CE, cbl_0..cbl_4], selected parameter gradients and the norm of
every parameter gradient for cases.model_batch() with cases.deterministic_init(seed=0) weights.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from oracle import cpu_pointops  # noqa: E402

REF = "/root/reference/pytorch"


def import_reference():
    for name in ("lib", "lib.pointops", "lib.pointops.functions"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["lib.pointops.functions.pointops"] = cpu_pointops
    sys.modules["lib.pointops.functions"].pointops = cpu_pointops
    sys.path.insert(0, REF)
    torch.cuda.IntTensor = lambda x: torch.tensor(x, dtype=torch.int32)     # blocks.py:68 on a CPU-only box
    from model import pointtransformer_seg as pts
    from util.config import CfgNode
    return pts, CfgNode


def main():
    pts, CfgNode = import_reference()
    cfg = CfgNode({
        "base_fdim": 32, "nsample": [36, 24, 24, 24, 24], "nstride": [4, 4, 4, 4], "ignore_label": 255,
        "voxel_size": 0.04,
        "contrast": {"stage": "Ua", "contrast": "softnn", "ftype": "latent", "sample": "label", "pos": "cnt",
                     "dist": "l2", "temperature": 1, "weight": "w.1"},
        "multi": {"stage": "Ua", "ftype": "latent", "combine": "concat"},
    }, default="")
    torch.manual_seed(0)
    model = pts.pointtransformer_seg_repro(c=6, k=13, config=cfg)
    crit = pts.Loss(cfg)
    cases.deterministic_init(model, 0)
    model.train()
    b = cases.model_batch()
    inputs = {"points": torch.from_numpy(b["points"]), "features": torch.from_numpy(b["features"]),
              "offset": torch.from_numpy(b["offset"])}
    target = torch.from_numpy(b["point_labels"])
    out, stage_list = model(inputs)
    loss = crit(out, target, stage_list)
    loss.sum().backward()
    res = {"logits": out.detach().numpy(), "loss": loss.detach().numpy()}
    norms = {}
    for name, p in model.named_parameters():
        if p.grad is not None:
            norms[name] = float(p.grad.norm())
            if name in cases.GOLDEN_GRADS:
                res["grad/" + name] = p.grad.numpy()
    res["grad_norms_json"] = np.frombuffer(json.dumps(norms).encode(), dtype=np.uint8)
    for i in range(5):
        res[f"latent/{i}"] = stage_list["up"][i]["latent"].detach().numpy()[:64]
    res["bn_running_mean/enc1.0.bn"] = model.enc1[0].bn.running_mean.numpy()
    out_path = os.path.join(ROOT, "tests", "golden", "model_ref.npz")
    np.savez_compressed(out_path, **res)
    print("wrote", out_path, "loss", res["loss"], "params with grad", len(norms))


if __name__ == "__main__":
    main()
