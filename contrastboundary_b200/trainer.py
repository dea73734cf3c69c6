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


# ------------------------------------------------------------------------------------------------------
# the rest of the trainer's per-step / per-epoch bookkeeping (train.py:148-149,265-284,315-345)
# ------------------------------------------------------------------------------------------------------
def build_model(cfg=None, sync_bn=False, device=None):
    """`Model(c=, k=, config=)` + the `sync_bn` switch of train.py:147-149.
    sync_bn=False (the reference's shipped setting): the fused libcbops layers with per-rank BatchNorm statistics.
    sync_bn=True: `nn.SyncBatchNorm.convert_sync_batchnorm(model)` exactly as the reference does it; the statistics of a
    synchronised BatchNorm need a collective in the middle of what the fused layer kernels compute in one pass, so this
    option runs the network op by op (model.set_fused(False)) on the stand-alone operators."""
    import torch.nn as nn
    from .model import CBLConfig, PointTransformerSeg
    cfg = cfg or CBLConfig()
    model = PointTransformerSeg(cfg)
    if sync_bn:
        model.set_fused(False)
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    return model.to(device) if device is not None else model


def pack_step_metrics(loss, output, target, classes, ignore_label=255):
    """everything the reference all-reduces per step (train.py:328-338: loss * n, count, intersection, union, target —
    five collectives and an `.item()`), packed into ONE float64 device vector of 2 + len(loss) + 3 * classes entries:
    [n, len(loss), loss * n ..., intersection ..., union ..., target ...].  No host synchronisation."""
    from .boundary_eval import intersection_and_union
    n = target.shape[0]
    pred = output.max(1)[1]
    i, u, t = intersection_and_union(pred, target, classes, ignore_label)
    head = torch.tensor([float(n), float(loss.numel())], dtype=torch.float64, device=loss.device)
    return torch.cat([head, loss.detach().double() * n, i.double(), u.double(), t.double()])


class MetricsAccumulator:
    """AverageMeter bookkeeping of train()/validate() (train.py:300-306,339-356) without per-step host reads: packed
    step metrics are summed on the device, all-reduced once per `reduce_every` steps (ONE collective of 46 doubles for the
    shipped 6-loss / 13-class configuration, off the critical path) and read back once per epoch."""

    def __init__(self, classes, reduce_every=1):
        self.classes, self.reduce_every = classes, max(1, int(reduce_every))
        self.total = None          # reduced sums
        self.pending = None
        self.steps = 0

    def update(self, packed):
        self.pending = packed.clone() if self.pending is None else self.pending + packed
        self.steps += 1
        if self.steps % self.reduce_every == 0:
            self._flush()

    def _flush(self):
        if self.pending is None:
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            nl = self.pending[1].clone()
            dist.all_reduce(self.pending)                                      # one collective for everything
            self.pending[1] = nl                                               # (the loss count is not a sum over ranks)
        self.total = self.pending if self.total is None else self.total + self.pending
        self.pending = None

    def summary(self):
        """-> (loss vector averaged over points, mIoU, mAcc, allAcc)   (train.py:358-364); one device->host read"""
        self._flush()
        t = self.total.cpu().numpy()
        nl = int(round(t[1] / max(self.steps, 1)))                             # entry 1 = (#losses) summed over the steps
        n = t[0]
        loss = t[2:2 + nl] / max(n, 1.0)
        k = self.classes
        inter, union, target = t[2 + nl:2 + nl + k], t[2 + nl + k:2 + nl + 2 * k], t[2 + nl + 2 * k:2 + nl + 3 * k]
        iou = inter / (union + 1e-10)
        acc = inter / (target + 1e-10)
        return loss, float(iou.mean()), float(acc.mean()), float(inter.sum() / (target.sum() + 1e-10))


class ScalarLog:
    """`writer.add_scalar(tag, value, epoch)` of train.py:265-284 (loss_train, loss_train_<i>, mIoU_train, mAcc_train,
    allAcc_train and the *_val twins).  Uses TensorBoard's SummaryWriter when the package is importable, and always
    appends the same records to <save_path>/scalars.jsonl."""

    def __init__(self, save_path):
        import json
        self._json = json
        os.makedirs(save_path, exist_ok=True)
        self.path = os.path.join(save_path, "scalars.jsonl")
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.tb = SummaryWriter(save_path)
        except Exception:
            self.tb = None

    def add_scalar(self, tag, value, step):
        value = float(value)
        with open(self.path, "a") as f:
            f.write(self._json.dumps({"tag": tag, "value": value, "step": int(step)}) + "\n")
        if self.tb is not None:
            self.tb.add_scalar(tag, value, step)

    def log_epoch(self, split, loss, miou, macc, allacc, epoch):
        self.add_scalar(f"loss_{split}", float(sum(loss)), epoch)
        for i, v in enumerate(loss):
            self.add_scalar(f"loss_{split}_{i}", v, epoch)
        self.add_scalar(f"mIoU_{split}", miou, epoch)
        self.add_scalar(f"mAcc_{split}", macc, epoch)
        self.add_scalar(f"allAcc_{split}", allacc, epoch)

    def close(self):
        if self.tb is not None:
            self.tb.close()
