"""Training-step harness for the Point-Transformer + CBL network (the bench / smoke driver).

Mirrors what the reference trainer does per iteration (pytorch/tool/train.py:315-326): inputs to
the GPU, forward, Loss = CE + CBL list, loss.sum().backward(), SGD step — and nothing else
(dataset I/O, logging, checkpointing are out of scope)."""
import torch

from .model import CBLConfig, Loss, PointTransformerSeg, build_geometry


class TrainStep:
    def __init__(self, cfg: CBLConfig = None, device="cuda", ddp=False, lr=0.5, momentum=0.9, weight_decay=1e-4, seed=0):
        self.cfg = cfg or CBLConfig()
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.model = PointTransformerSeg(self.cfg).to(self.device)
        self.criterion = Loss(self.cfg).to(self.device)
        self.net = self.model
        if ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            self.net = DDP(self.model, device_ids=[self.device.index], gradient_as_bucket_view=True)
        # reference optimiser: SGD(lr=base_lr, momentum, weight_decay) (train.py:154)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay,
                                   fused=self.device.type == "cuda")
        self.model.train()

    def step(self, batch, update=True):
        """batch: dict of DEVICE tensors points/features/point_labels/offset + python list offset_host.
        returns the stacked loss vector [CE, cbl_0..cbl_4] (device tensor)."""
        self.opt.zero_grad(set_to_none=True)
        out, stages = self.net(batch)
        loss = self.criterion(out, batch["point_labels"], stages)
        loss.sum().backward()
        if update:
            self.opt.step()
        return loss.detach()


def to_device(host_batch, device, non_blocking=True):
    """host_batch: dict of (pinned) CPU tensors + offset_host list"""
    out = {}
    for k, v in host_batch.items():
        out[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
    return out


def host_batch_from_numpy(b, pin=True):
    t = {
        "points": torch.from_numpy(b["points"]),
        "features": torch.from_numpy(b["features"]),
        "point_labels": torch.from_numpy(b["point_labels"]),
        "offset": torch.from_numpy(b["offset"]),
    }
    if pin and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    t["offset_host"] = [int(x) for x in b["offset"]]
    return t
