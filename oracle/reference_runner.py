"""oracle/reference_runner.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Builds the REFERENCE network for the baseline legs of bench.py (`--impl reference`, `cpu_baseline`, `gpu_baseline`):
the reference's own, unmodified Python model code (pytorch/model/*.py — read from /root/reference in the build
container, from its build-time copy baseline/_ref/pytorch on the GPU box) on top of
  backend "cpu"        oracle.cpu_pointops: the C/OpenMP restatement of the reference's CUDA kernels (pinned bit-for-bit
                       against them by tests/golden/pointops_ref_gpu.npz)  -> the reference's CPU path
  backend "stock-gpu"  the reference's own pointops.py + its own CUDA extension compiled unmodified for sm_100a
                       (oracle/_ref/pointops_cuda.so)                      -> baseline B1 of BASELINE.md §3
When the copy of the model code is absent, the op-by-op restatement oracle/ref_model.py takes its place.
None of this repo's kernels is on either path."""
import json
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

REF_CFG = {
    "base_fdim": 32, "nsample": [36, 24, 24, 24, 24], "nstride": [4, 4, 4, 4], "ignore_label": 255, "voxel_size": 0.04,
    "contrast": {"stage": "Ua", "contrast": "softnn", "ftype": "latent", "sample": "label", "pos": "cnt",
                 "dist": "l2", "temperature": 1, "weight": "w.1"},
    "multi": {"stage": "Ua", "ftype": "latent", "combine": "concat"},
}


def _refpy():
    for p in ("/root/reference/pytorch", os.path.join(ROOT, "baseline", "_ref", "pytorch")):
        if os.path.exists(os.path.join(p, "model", "blocks.py")) and os.path.exists(os.path.join(p, "util", "config.py")):
            return p
    return None


def build(backend):
    """-> (model, criterion, description).  forward: out, stage_list = model(inputs); loss = criterion(out, target, stage_list)"""
    import warnings

    import torch
    refpy = _refpy()
    if backend == "cpu":
        from oracle import cpu_pointops as ops
    else:
        ops = None
    if refpy is None:
        from oracle import ref_model
        if ops is None:
            from oracle import gpu_pointops as ops
        return ref_model.RefSeg(ops), ref_model.RefLoss(ops), "restated reference network (oracle/ref_model.py)"
    if backend == "cpu":
        for name in ("lib", "lib.pointops", "lib.pointops.functions"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["lib.pointops.functions.pointops"] = ops
        sys.modules["lib.pointops.functions"].pointops = ops
        torch.cuda.IntTensor = lambda x: torch.tensor(x, dtype=torch.int32)     # blocks.py:68 on a CPU run
    else:
        lib_py = refpy if os.path.exists(os.path.join(refpy, "lib", "pointops", "functions", "pointops.py")) else None
        assert lib_py, "the reference's lib/pointops/functions/pointops.py is missing"
        sys.path.insert(0, os.path.join(_HERE, "_ref"))                         # `import pointops_cuda` -> the stock build
    sys.path.insert(0, refpy)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from model import pointtransformer_seg as pts
        from util.config import CfgNode
    cfg = CfgNode(json.loads(json.dumps(REF_CFG)), default="")
    model = pts.pointtransformer_seg_repro(c=6, k=13, config=cfg)
    crit = pts.Loss(cfg)
    return model, crit, "the reference's own model code (pytorch/model/*.py, unmodified)"
