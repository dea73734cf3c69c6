"""developer tool (GPU): attribute a gradient discrepancy of the ConvNet path to one operator by swapping the libcbops
operators for plain-torch autograd one at a time.   python tools/debug_convnet_grads.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from contrastboundary_b200 import convnet, linear_ops, synthetic  # noqa: E402
from oracle import tf_model as T  # noqa: E402


def batch(sizes, seed, dev):
    scenes = [synthetic.make_scene(n, seed + i) for i, n in enumerate(sizes)]
    return {"points": torch.from_numpy(np.concatenate([s[0] for s in scenes])).to(dev),
            "colors": torch.from_numpy(np.concatenate([s[1] for s in scenes])).to(dev),
            "point_labels": torch.from_numpy(np.concatenate([s[2] for s in scenes])).to(dev),
            "lens": torch.tensor(sizes, dtype=torch.int32, device=dev)}


ORIG = {"aw": convnet.adaptive_weight, "pool": convnet.ind_max_pool, "cbl": convnet.tf_contrast_loss, "min_rows": linear_ops.MIN_ROWS,
        "min_rows_w": linear_ops.MIN_ROWS_WGRAD, "bn": linear_ops.FUSED_BN}


def torch_pool(x, inds):
    xs = torch.cat([x, x.min(0, keepdim=True)[0].detach()], 0)
    return xs[inds.long()].max(1)[0]


def configure(fused):
    """fused: set of operator names that stay on libcbops"""
    convnet.adaptive_weight = ORIG["aw"] if "aw" in fused else (lambda q, s, nb, f, w, b, r: T.adaptive_weight(q, s, nb, f, w, b, r))
    convnet.ind_max_pool = ORIG["pool"] if "pool" in fused else torch_pool
    convnet.tf_contrast_loss = ORIG["cbl"] if "cbl" in fused else (lambda f, nb, c, t, w: T.contrast_loss(f, nb, c.long(), t, w))
    linear_ops.MIN_ROWS = ORIG["min_rows"] if "linear" in fused else 1 << 60
    linear_ops.MIN_ROWS_WGRAD = ORIG["min_rows_w"] if "linear" in fused else 1 << 60
    linear_ops.FUSED_BN = ORIG["bn"] if "bn" in fused else False


def grads(ts, inputs):
    ts.model.zero_grad(set_to_none=True)
    logits, sl = ts.model(inputs)
    loss = ts.criterion(logits, inputs["point_labels"], sl)
    loss.sum().backward()
    return {n: p.grad.detach().double().clone() for n, p in ts.model.named_parameters() if p.grad is not None}, loss.detach().cpu().numpy()


def main():
    dev = torch.device("cuda", 0)
    ts = convnet.ConvNetTrainStep(convnet.ConvNetConfig(), dev, seed=1)
    inputs = ts.build_inputs(batch([4000, 3500], 300, dev))
    allops = {"aw", "pool", "cbl", "linear", "bn"}
    configure(set())
    ref, lref = grads(ts, inputs)
    print("torch-only loss", lref)
    for name, fused in [("ALL", allops)] + [(o, {o}) for o in sorted(allops)]:
        configure(fused)
        g, l = grads(ts, inputs)
        rows = sorted(((float((g[n] - ref[n]).norm() / ref[n].norm().clamp(min=1e-30)), n) for n in ref), reverse=True)
        tot = float(torch.cat([(g[n] - ref[n]).reshape(-1) for n in ref]).norm() / torch.cat([ref[n].reshape(-1) for n in ref]).norm())
        print(f"== libcbops for {name}: loss {l}, whole-gradient rel diff {tot:.3e}; worst tensors:")
        for r in rows[:6]:
            print("     %.3e  %s" % r)
    configure(allops)


if __name__ == "__main__":
    main()
