"""SURVEY §8(f) row 4 — the pieces of the reference trainer that sit directly around the step
(pytorch/tool/train.py): optimiser and learning-rate schedule (:154-165), checkpoint format and resume (:198-224,
:289-296; test-time loader pytorch/tool/test.py:100-112).  Dataset I/O, logging and TensorBoard stay out of scope.

A checkpoint written here loads in the reference and vice versa: the dictionary keys, the `module.` prefix that
DistributedDataParallel puts on every state_dict key there, and the optimizer / scheduler state dicts are the same."""
import collections
import os
import shutil

import torch
from torch.optim import lr_scheduler


def build_optimizer(model, base_lr=0.5, momentum=0.9, weight_decay=1e-4, fused=None):
    """SGD exactly as train.py:154 (config/s3dis/*.yaml: base_lr 0.5, momentum 0.9, weight_decay 1e-4)"""
    params = list(model.parameters())
    if fused is None:
        fused = bool(params) and params[0].is_cuda
    return torch.optim.SGD(params, lr=base_lr, momentum=momentum, weight_decay=weight_decay, fused=fused)


def build_scheduler(optimizer, epochs, name="multistep", milestones=(0.6, 0.8), gamma=0.1, **step_kwargs):
    """train.py:156-165: MultiStepLR at int(epochs * s) for s in milestones (default 60 % / 80 %, gamma 0.1), or StepLR"""
    if name == "multistep":
        assert all(0 < s < 1 for s in milestones), f"invalid milestones ( <0 or >1 ) - {milestones}"
        return lr_scheduler.MultiStepLR(optimizer, milestones=[int(epochs * s) for s in milestones], gamma=gamma)
    if name == "step":
        return lr_scheduler.StepLR(optimizer, **step_kwargs)
    raise ValueError(f"not support scheduler = {name}")


def _with_prefix(state_dict, prefix):
    return collections.OrderedDict((prefix + k, v) for k, v in state_dict.items())


def _strip_prefix(state_dict, prefix="module."):
    if state_dict and all(k.startswith(prefix) for k in state_dict):
        return collections.OrderedDict((k[len(prefix):], v) for k, v in state_dict.items())      # test.py:106-109
    return state_dict


def save_checkpoint(path, epoch, model, optimizer, scheduler, best_iou, is_best=False, ddp_prefix=True):
    """train.py:289-296.  ddp_prefix: write `module.`-prefixed keys, as the reference's DDP-wrapped model does, so that
    the reference's test.py (which strips 7 characters unconditionally, :107) can read the file."""
    sd = model.state_dict()
    if ddp_prefix:
        sd = _with_prefix(sd, "module.")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    osd = optimizer.state_dict()
    # `fused` is an implementation switch of THIS process' optimiser, not training state: a reference resume
    # (train.py:214) would load it into its own SGD's param_groups
    osd = {"state": osd["state"], "param_groups": [{k: v for k, v in g.items() if k != "fused"} for g in osd["param_groups"]]}
    torch.save({"epoch": epoch, "state_dict": sd, "optimizer": osd, "scheduler": scheduler.state_dict(),
                "best_iou": best_iou, "is_best": is_best}, path)
    if is_best:
        shutil.copyfile(path, os.path.join(os.path.dirname(os.path.abspath(path)), "model_best.pth"))


def load_weights(path, model, map_location=None):
    """`weight:` of the reference config (train.py:198-207) and the test-time loader (test.py:100-112)"""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(_strip_prefix(ckpt["state_dict"]), strict=True)
    return ckpt.get("epoch", 0)


def resume(path, model, optimizer, scheduler, map_location=None):
    """train.py:209-224 -> (start_epoch, best_iou)"""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(_strip_prefix(ckpt["state_dict"]), strict=True)
    osd = ckpt["optimizer"]
    cur = optimizer.state_dict()["param_groups"]
    groups = [dict(g, **({"fused": c["fused"]} if "fused" in c and "fused" not in g else {})) for g, c in zip(osd["param_groups"], cur)]
    optimizer.load_state_dict({"state": osd["state"], "param_groups": groups})
    scheduler.load_state_dict(ckpt["scheduler"])
    return ckpt["epoch"], ckpt["best_iou"]
