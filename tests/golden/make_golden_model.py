"""Generate tests/golden/model_ref.npz and tests/golden/model_ref_cfg2.npz by running the REAL reference
model code (/root/reference/pytorch/model/*.py, imported unmodified) on CPU in the build container,
with `lib.pointops.functions.pointops` provided by oracle/cpu_pointops.py (the C restatement of the
reference kernels, itself pinned bit-for-bit by tests/golden/pointops_ref_gpu.npz).

    python tests/golden/make_golden_model.py [small] [cfg2]

Each case runs the reference twice on identical inputs and weights (cases.deterministic_init(seed=0)):
  * float32 — what the reference computes;
  * float64 — the same network in double precision (the searches stay fp32, so the geometry is identical):
    the "truth" against which a gradient ERROR can be defined.
Why both: the gradient of this network is badly conditioned in fp32 (40 layers of train-mode BatchNorm, ReLU
flips near 0, a 3-channel BatchNorm inside every attention layer): the reference's own fp32 run differs from
its own fp64 run by up to several 1e-2 (relative, per parameter tensor) although losses agree to 1e-6.  The
golden therefore stores, per parameter, the float64 gradient norm and the reference's own fp32 errors
`ref32_err = |g32 - g64| / |g64|` and `ref32_norm_err = | |g32| - |g64| | / |g64|`; the GPU tests require the product's error against float64 to stay within
a small multiple of the reference's own.

Stored: `logits` / `loss` / `grad/<name>` / `grad_norms_json` / `latent/<i>` from the float32 run (small case:
every logit row; cfg2: the rows in `rows`), `f64/...` twins from the float64 run, `ref32_err_json`.
  small = cases.model_batch()                      (4096 + 3000 points: padding paths, tiny deep levels)
  cfg2  = synthetic.make_batch(4, 40960, 5000)     (BASELINE configs[1], the first batch bench.py times)
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

REF_CFG = {
    "base_fdim": 32, "nsample": [36, 24, 24, 24, 24], "nstride": [4, 4, 4, 4], "ignore_label": 255,
    "voxel_size": 0.04,
    "contrast": {"stage": "Ua", "contrast": "softnn", "ftype": "latent", "sample": "label", "pos": "cnt",
                 "dist": "l2", "temperature": 1, "weight": "w.1"},
    "multi": {"stage": "Ua", "ftype": "latent", "combine": "concat"},
}


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


def run_reference(pts, CfgNode, b, dtype):
    cfg = CfgNode(json.loads(json.dumps(REF_CFG)), default="")
    torch.manual_seed(0)
    model = pts.pointtransformer_seg_repro(c=6, k=13, config=cfg)
    crit = pts.Loss(cfg)
    cases.deterministic_init(model, 0)
    model = model.to(dtype)
    model.train()
    inputs = {"points": torch.from_numpy(b["points"]).to(dtype), "features": torch.from_numpy(b["features"]).to(dtype),
              "offset": torch.from_numpy(b["offset"])}
    target = torch.from_numpy(b["point_labels"])
    out, stage_list = model(inputs)
    loss = crit(out, target, stage_list)
    loss.sum().backward()
    return model, out.detach(), loss.detach(), stage_list


def generate(pts, CfgNode, b, out_path, rows=None):
    res = {}
    grads = {}
    for tag, dtype in (("", torch.float32), ("f64/", torch.float64)):
        model, out, loss, stage_list = run_reference(pts, CfgNode, b, dtype)
        o = out.numpy()
        res[tag + "logits"] = (o if rows is None else o[rows]).astype(np.float32)
        res[tag + "loss"] = loss.numpy()
        norms = {}
        grads[tag] = {}
        for name, p in model.named_parameters():
            if p.grad is not None:
                norms[name] = float(p.grad.double().norm())
                grads[tag][name] = p.grad.double()
                if name in cases.GOLDEN_GRADS:
                    res[tag + "grad/" + name] = p.grad.numpy().astype(np.float32)
        res[tag + "grad_norms_json"] = np.frombuffer(json.dumps(norms).encode(), dtype=np.uint8)
        for i in range(5):
            res[tag + f"latent/{i}"] = stage_list["up"][i]["latent"].detach().numpy()[:64].astype(np.float32)
        res[tag + "bn_running_mean/enc1.0.bn"] = model.enc1[0].bn.running_mean.numpy().astype(np.float32)
    err = {n: float((grads[""][n] - g).norm() / g.norm().clamp(min=1e-300)) for n, g in grads["f64/"].items()}
    res["ref32_err_json"] = np.frombuffer(json.dumps(err).encode(), dtype=np.uint8)
    nerr = {n: float(abs(grads[""][n].norm() - g.norm()) / g.norm().clamp(min=1e-300)) for n, g in grads["f64/"].items()}
    res["ref32_norm_err_json"] = np.frombuffer(json.dumps(nerr).encode(), dtype=np.uint8)
    if rows is not None:
        res["rows"] = rows
    np.savez_compressed(out_path, **res)
    live = [v for n, v in err.items() if not cases.grad_is_analytically_zero(n)]
    print("wrote", out_path, "loss", res["loss"], "f64", res["f64/loss"], "params with grad", len(err),
          "ref fp32-vs-fp64 gradient error: median %.2e max %.2e" % (float(np.median(live)), float(np.max(live))))


def main():
    which = set(sys.argv[1:]) or {"small", "cfg2"}
    pts, CfgNode = import_reference()
    gdir = os.path.join(ROOT, "tests", "golden")
    if "small" in which:
        generate(pts, CfgNode, cases.model_batch(), os.path.join(gdir, "model_ref.npz"))
    if "cfg2" in which:
        b = cases.model_batch_cfg2()
        n = b["points"].shape[0]
        generate(pts, CfgNode, b, os.path.join(gdir, "model_ref_cfg2.npz"), rows=np.arange(0, n, 41, dtype=np.int64))


if __name__ == "__main__":
    main()
