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
        self.side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None     # sampling chain
        self.side2 = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None    # neighbour searches
        self._geo = {}

    # ------------------------------------------------------------------------------------------
    # Geometry (FPS chain + every neighbour search) depends on coordinates only.  Like the reference's
    # TF input pipeline, which builds the neighbour pyramid of the NEXT batch on host workers while the
    # GPU trains on the current one (tensorflow/datasets/base.py:75-118,767-842), the geometry of batch
    # t+1 can be computed on a side stream while batch t is in forward/backward.  Every batch's geometry
    # is still computed exactly once.
    # ------------------------------------------------------------------------------------------
    def prefetch_geometry(self, batch):
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)                      # the batch's H2D copies were enqueued on `main`
        with torch.cuda.stream(self.side):
            levels = build_geometry(batch["points"], batch["offset"], batch["offset_host"], self.cfg,
                                    self.cfg.contrast is not None, knn_stream=self.side2)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self._geo[id(batch)] = (levels, ev)

    def _take_geometry(self, batch):
        geo = self._geo.pop(id(batch), None)
        if geo is None:
            return None
        levels, ev = geo
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        for lv in levels:                                 # allocated on the side stream, consumed on `main`
            for name in lv.__slots__:
                t = getattr(lv, name)
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(main)
        return levels

    def step(self, batch, update=True, next_batch=None):
        """batch: dict of DEVICE tensors points/features/point_labels/offset + python list offset_host.
        next_batch: optional batch whose geometry is computed on the side stream during this step.
        returns the stacked loss vector [CE, cbl_0..cbl_4] (device tensor)."""
        self.opt.zero_grad(set_to_none=True)
        levels = self._take_geometry(batch)
        if next_batch is not None:
            self.prefetch_geometry(next_batch)
        out, stages = self.net(batch, levels)
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
