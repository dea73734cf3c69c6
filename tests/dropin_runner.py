"""Runs the REFERENCE's own Python model code (baseline/_ref/pytorch/model/*.py — copied, unmodified, from
/root/reference at build time by oracle/build_ref.sh; git-ignored) on the GPU on top of this repo's operators and
prints one JSON line of errors against tests/golden/model_ref.npz.  Launched by tests/test_dropin_gpu.py in a
subprocess (the reference's top-level package names `model`, `lib`, `util` stay out of the test process).

    python tests/dropin_runner.py python_api    reference model + `lib.pointops.functions.pointops` := contrastboundary_b200.pointops
    python tests/dropin_runner.py native_abi    reference model + the reference's OWN pointops.py + its OWN pybind glue,
                                                linked against libcbops.so (oracle/_ref/dropin/pointops_cuda.so)
    python tests/dropin_runner.py loss_adapter  reference model (python_api) + THIS repo's Loss(config) / ContrastHead(head_cfg,
                                                config) called with the reference's forward(output, target, stage_list)
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REFPY = os.path.join(ROOT, "baseline", "_ref", "pytorch")


def main(mode):
    import cases
    from make_golden_model import REF_CFG
    if mode == "native_abi":
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "dropin"))      # `import pointops_cuda` -> the drop-in build
        sys.path.insert(0, REFPY)
        import pointops_cuda
        assert "dropin" in pointops_cuda.__file__, pointops_cuda.__file__
    else:
        from contrastboundary_b200 import pointops as ours
        for name in ("lib", "lib.pointops", "lib.pointops.functions"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["lib.pointops.functions.pointops"] = ours
        sys.modules["lib.pointops.functions"].pointops = ours
        sys.path.insert(0, REFPY)
    from model import pointtransformer_seg as pts
    from util.config import CfgNode
    assert pts.__file__.startswith(REFPY), pts.__file__
    cfg = CfgNode(json.loads(json.dumps(REF_CFG)), default="")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = pts.pointtransformer_seg_repro(c=6, k=13, config=cfg)
    cases.deterministic_init(model, 0)
    model = model.to(dev).train()
    if mode == "loss_adapter":
        from contrastboundary_b200 import model as M
        crit = M.Loss(cfg).to(dev)                                   # the reference's constructor argument, unchanged
        head = M.ContrastHead(cfg.contrast, cfg)                     # heads.py:66 signature
        assert len(head.stages) == 5
    else:
        crit = pts.Loss(cfg).to(dev)
    b = cases.model_batch()
    inputs = {"points": torch.from_numpy(b["points"]).to(dev), "features": torch.from_numpy(b["features"]).to(dev),
              "offset": torch.from_numpy(b["offset"]).to(dev)}
    target = torch.from_numpy(b["point_labels"]).to(dev)
    out, stage_list = model(inputs)
    loss = crit(out, target, stage_list)
    loss.sum().backward()
    torch.cuda.synchronize()
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_ref.npz"))
    res = {"mode": mode}
    o = out.detach().cpu().numpy()
    res["logits_err"] = float(np.abs(o - g["logits"]).max() / np.abs(g["logits"]).max())
    res["loss"] = [float(x) for x in loss.detach().cpu()]
    res["loss_ref"] = [float(x) for x in g["loss"]]
    res["latent_err"] = [float(np.abs(stage_list["up"][i]["latent"].detach().cpu().numpy()[:64] - g[f"latent/{i}"]).max()
                               / max(np.abs(g[f"latent/{i}"]).max(), 1e-6)) for i in range(5)]
    rep = cases.grad_report_vs_f64({n: p.grad for n, p in model.named_parameters() if p.grad is not None}, g)
    cases.print_grad_report(rep, mode)
    res["grad_failures"] = cases.grad_failures(rep, cases.GRAD_FACTOR_TORCH_CUDA)
    res["grad_quantiles"] = {str(q): v for q, v in rep["quantiles"].items()}
    loaded = [l.split()[-1] for l in open("/proc/self/maps") if l.rstrip().endswith(".so") and ("cbops" in l or "pointops_cuda" in l)]
    res["loaded"] = sorted(set(os.path.relpath(x, ROOT) for x in loaded))
    print("DROPIN " + json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1])
