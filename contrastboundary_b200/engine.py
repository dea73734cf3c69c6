"""Training-step harness for the Point-Transformer + CBL network (the bench / smoke driver).

Mirrors what the reference trainer does per iteration (pytorch/tool/train.py:315-326): inputs to
the GPU, forward, Loss = CE + CBL list, loss.sum().backward(), SGD step — and nothing else
(dataset I/O, logging, checkpointing are out of scope)."""
import torch

from .model import CBLConfig, Loss, PointTransformerSeg, build_geometry


import os as _os
WGRAD_FORK = _os.environ.get("CB_WGRAD_FORK", "1") != "0"   # weight gradients of the linear layers on a side stream during the
                                                            # backward (linear_ops.wgrad_fork); CB_WGRAD_FORK=0 keeps them in line


def _backward(loss, fork=True):
    """loss.sum().backward() as the reference's iteration does it (pytorch/tool/train.py:322-324), with the weight-gradient
    kernels of the linear layers forked to a side stream and joined before anyone reads the gradients.  fork=False under
    torch's DistributedDataParallel: its gradient hooks read every gradient the moment autograd hands it over."""
    if WGRAD_FORK and fork and loss.is_cuda:
        from .linear_ops import wgrad_fork
        with wgrad_fork():
            loss.sum().backward()
    else:
        loss.sum().backward()


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
        self._geo[id(batch)] = (batch, levels, ev)        # the batch object itself is kept: id() alone can be recycled

    def _take_geometry(self, batch):
        geo = self._geo.pop(id(batch), None)
        self._geo.clear()                                 # geometry prefetched for a batch that never came: drop it
        if geo is None or geo[0] is not batch:
            return None
        _, levels, ev = geo
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
        _backward(loss, fork=self.net is self.model)
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


# ------------------------------------------------------------------------------------------------------
# Whole-step CUDA graphs.
#
# A training step of this network is ~1400 kernel launches, most of them a few microseconds long: issued
# one by one from Python the HOST is the bottleneck (tools/host_bound.py).  Shapes are static for a given
# tuple of scene sizes, and no kernel of libcbops needs a host decision (tie replays, boundary masks and
# flagged-query lists are all device-side), so the step is captured ONCE per batch signature into two CUDA
# graphs per slot and replayed:
#     geo[s]  coordinates of slot s -> every level's sampling / neighbour search / relative positions
#     net[s]  inputs + geometry of slot s -> forward, Loss, backward, gradients packed into one flat buffer
# Two slots alternate: while net[s] trains on batch t (main stream), geo[1-s] builds the geometry of batch
# t+1 (side stream) — the same look-ahead as TrainStep.prefetch_geometry, now without any launch overhead.
# The optimiser (and, data-parallel, ONE NCCL all-reduce of the flat gradient) stays outside the graphs, so a
# learning-rate schedule needs no re-capture.
# ------------------------------------------------------------------------------------------------------
class _Slot:
    def __init__(self):
        self.inputs = None      # static device tensors: points, features, point_labels, offset (+ offset_host list)
        self.o_flat = None
        self.levels = None
        self.geo = self.net = None
        self.loss = None
        self.ev_geo = None
        self.holds = None       # the batch (dict object) whose data / geometry the slot currently holds


class GraphTrainStep(TrainStep):
    """TrainStep whose step() replays captured CUDA graphs.  Same numerics (the captured work IS TrainStep's
    work); falls back to TrainStep's stream mode for batch signatures it could not capture."""

    def __init__(self, cfg: CBLConfig = None, device="cuda", ddp=False, eager_warmup=2, max_signatures=4, net_priority=False, **kw):
        # data parallel here = one all-reduce of the packed gradient after the graph; no DDP wrapper
        super().__init__(cfg, device, ddp=False, **kw)
        import torch.distributed as dist
        self.world = dist.get_world_size() if (ddp and dist.is_initialized()) else 1
        if self.world > 1:                       # replicas must start from identical parameters (DDP does this itself)
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, 0)
        self.eager_warmup = max(1, eager_warmup)
        self.max_signatures = max_signatures
        # signature (tuple of scene sizes) -> [slot0, slot1] | None (capture failed); least-recently-used first.  A captured
        # signature pins ~5 GB of activations for 4 x 40960 points, so the cache is small and evicts (the reference's loader
        # crops every cloud larger than voxel_max to exactly voxel_max points, data_util.py:64-67: batches of equal-size scenes
        # recur; anything else runs in stream mode)
        import collections
        self._sigs = collections.OrderedDict()
        self._seen = {}
        self._eager_done = 0
        self._net_pool = None      # net[0] / net[1] replay back to back on one stream: they may share temporaries
        self.flat = None
        self.pflat = None                   # device table of the one-launch optimiser step (see _pack_params)
        self._opt_refs = None
        self._gparams = None
        self.launches_per_step = None
        self.graph_error = None
        self._packed = False       # True while every p.grad is a view of self.flat
        # optional (measured: no effect on this workload): capture / replay the network graph on a HIGH-priority stream,
        # so that the concurrently replaying geometry graph of the next batch only fills the SMs the network leaves idle
        self.hp = torch.cuda.Stream(device=self.device, priority=-1) if net_priority else None

    # -- eager (stream-mode) step that also averages gradients across ranks -----------------------------
    def _set_packed(self, flag):
        if flag == self._packed:
            return
        for p in self.model.parameters():
            p.grad = None
        if flag:
            o = 0
            for p in self._gparams:
                p.grad = self.flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        self._packed = flag

    def _eager_step(self, batch, update=True):
        batch = to_device(batch, self.device)
        self._set_packed(False)
        self.opt.zero_grad(set_to_none=True)
        self._geo.clear()
        out, stages = self.model(batch, None)
        loss = self.criterion(out, batch["point_labels"], stages)
        _backward(loss)
        if self._gparams is None:
            # parameters that receive a gradient (the rest keep grad None, as in stream mode)
            self._gparams = [p for p in self.model.parameters() if p.grad is not None]
        if self.world > 1:
            import torch.distributed as dist
            gs = [p.grad for p in self.model.parameters() if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in gs])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            o = 0
            for g in gs:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        if update:
            self.opt.step()
        return loss.detach()

    def stream_loss_and_grad(self, batch):
        """test hook: stream-mode forward + backward on the CURRENT parameters, no update.
        Returns (loss vector, packed gradient in the order of the graph's flat buffer)."""
        batch = to_device(batch, self.device)
        was = self._packed
        self._set_packed(False)
        self.opt.zero_grad(set_to_none=True)
        out, stages = self.model(batch, None)
        loss = self.criterion(out, batch["point_labels"], stages)
        _backward(loss)
        ps = self._gparams if self._gparams is not None else [p for p in self.model.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        for p in self.model.parameters():
            p.grad = None
        self._set_packed(was)
        return loss.detach(), flat

    # -- one-launch optimiser ---------------------------------------------------------------------------
    def _pack_params(self):
        """Once, before the first capture: a device table (offsets into self.flat, parameter pointers, momentum-buffer
        pointers) so that the optimiser step after a graph replay is ONE kernel (cb_sgd_momentum_step) instead of torch's 39
        multi-tensor launches.  Nothing is re-allocated: parameters and torch.optim.SGD's momentum buffers stay where they
        are (stream-mode steps, state_dict / checkpoints keep working on the same tensors)."""
        if self.pflat is not None or self.device.type != "cuda":
            return
        g = self.opt.param_groups
        ok = (len(g) == 1 and g[0].get("momentum", 0) > 0 and g[0].get("dampening", 0) == 0 and not g[0].get("nesterov", False)
              and not g[0].get("maximize", False)
              and all(p.dtype == torch.float32 and p.is_contiguous() for p in self._gparams)
              and all(self.opt.state.get(p, {}).get("momentum_buffer") is not None for p in self._gparams))
        if not ok:
            self.pflat = False          # unusual optimiser configuration (or no step taken yet): keep torch's step
            return
        self._bind_opt_table()
        if hasattr(self.opt, "register_load_state_dict_post_hook"):
            self.opt.register_load_state_dict_post_hook(lambda opt: setattr(self, "_opt_refs", None))

    def _bind_opt_table(self):
        offs, o = [0], 0
        for p in self._gparams:
            o += p.numel()
            offs.append(o)
        mbs = [self.opt.state[p]["momentum_buffer"] for p in self._gparams]
        assert all(m.is_contiguous() and m.dtype == torch.float32 for m in mbs)
        self._opt_refs = (list(self._gparams), mbs, [p.data_ptr() for p in self._gparams])
        self.pflat = (torch.tensor(offs, dtype=torch.int64, device=self.device),
                      torch.tensor([p.data_ptr() for p in self._gparams], dtype=torch.int64, device=self.device),
                      torch.tensor([m.data_ptr() for m in mbs], dtype=torch.int64, device=self.device), o)

    def _opt_step(self):
        """optimiser step on the packed gradient (self.flat holds this step's gradient, every p.grad is a view of it)"""
        if not isinstance(self.pflat, tuple):
            self.opt.step()
            return
        # someone may have re-bound the storage since (optimizer.load_state_dict replaces the momentum tensors, module.to() the
        # parameters): rebuild the table instead of updating stale memory
        refs = self._opt_refs
        if (refs is None or refs[0][0].data_ptr() != refs[2][0] or refs[0][-1].data_ptr() != refs[2][-1]
                or self.opt.state[refs[0][0]]["momentum_buffer"] is not refs[1][0]
                or self.opt.state[refs[0][-1]]["momentum_buffer"] is not refs[1][-1]):
            self._bind_opt_table()
        import ctypes as C
        from . import _lib as L
        g = self.opt.param_groups[0]
        off, pp, mp, total = self.pflat
        rc = L.lib().cb_sgd_momentum_step(C.c_longlong(total), C.c_int(len(self._gparams)), L.ptr(off), L.ptr(pp), L.ptr(mp),
                                          L.ptr(self.flat), C.c_float(float(g["lr"])), C.c_float(float(g["momentum"])),
                                          C.c_float(float(g["weight_decay"])), C.c_int(0), L.stream())
        L.check(rc, "cb_sgd_momentum_step")
        self.opt._opt_called = True         # what lr_scheduler.step() looks at to tell that the optimiser has stepped

    # -- capture ------------------------------------------------------------------------------------------
    def _static_inputs(self, batch, sig):
        n, b = sig[-1], len(sig)
        dev = self.device
        return {"points": torch.empty((n, 3), dtype=torch.float32, device=dev),
                "features": torch.empty((n, batch["features"].shape[1]), dtype=torch.float32, device=dev),
                "point_labels": torch.empty((n,), dtype=batch["point_labels"].dtype, device=dev),
                "offset": torch.tensor(list(sig), dtype=torch.int32, device=dev),
                "offset_host": list(sig)}

    def _capture(self, batch, sig):
        from . import _lib as L
        from .model import level_offsets_host
        dev = self.device
        params = [p for p in self.model.parameters()] + [p for p in self.criterion.parameters()]
        if self.flat is None:
            assert self._gparams, "GraphTrainStep needs at least one stream-mode step before the capture"
            self.flat = torch.zeros(sum(p.numel() for p in self._gparams), dtype=torch.float32, device=dev)
        self._pack_params()
        ohs = level_offsets_host(list(sig), self.cfg)
        slots = [_Slot(), _Slot()]
        for sl in slots:
            sl.inputs = self._static_inputs(batch, sig)
            for k in ("points", "features", "point_labels"):
                sl.inputs[k].copy_(batch[k], non_blocking=True)       # valid data for the capture-time warm run
            sl.o_flat = torch.tensor([v for oh_ in ohs[1:] for v in oh_], dtype=torch.int32, device=dev)
            sl.ev_geo = torch.cuda.Event()
        torch.cuda.synchronize(dev)
        lc0 = L.launch_count()
        L.WS_NO_CACHE = True
        try:
            for sl in slots:
                # NO pool sharing between the two geometry graphs: geo[1-s] replays while net[s] still reads the
                # outputs of geo[s]; in a shared pool the temporaries of one graph may alias the outputs of the other
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sl.levels = build_geometry(sl.inputs["points"], sl.inputs["offset"], sl.inputs["offset_host"], self.cfg,
                                               self.cfg.contrast is not None, knn_stream=self.side2, o_flat=sl.o_flat)
                sl.geo = g
            lc1 = L.launch_count()
            for sl in slots:
                sl.geo.replay()                                        # real geometry for the net capture's shapes
            torch.cuda.synchronize(dev)
            for sl in slots:
                for p in params:
                    p.grad = None
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self._net_pool, **({"stream": self.hp} if self.hp is not None else {})):
                    out, stages = self.model(sl.inputs, sl.levels)
                    loss = self.criterion(out, sl.inputs["point_labels"], stages)
                    _backward(loss)
                    torch.cat([p.grad.reshape(-1) for p in self._gparams], out=self.flat)
                    sl.loss = loss.detach()
                sl.net = g
                if self._net_pool is None:
                    self._net_pool = g.pool()
            lc2 = L.launch_count()
        finally:
            L.WS_NO_CACHE = False
        for p in params:
            p.grad = None
        self._packed = False
        self.launches_per_step = (lc1 - lc0) // 2 + (lc2 - lc1) // 2 + (1 if isinstance(self.pflat, tuple) else 0)
        return slots

    # -- replay -------------------------------------------------------------------------------------------
    def _load(self, sl, batch):
        # a geometry replay of this slot that was never consumed may still be reading the inputs
        torch.cuda.current_stream(self.device).wait_event(sl.ev_geo)
        for k in ("points", "features", "point_labels"):
            sl.inputs[k].copy_(batch[k], non_blocking=True)            # H2D from pinned memory, or D2D
        sl.holds = batch

    def _launch_geo(self, sl):
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)                                    # inputs copied; previous reader of sl.levels done
        with torch.cuda.stream(self.side):
            sl.geo.replay()
            sl.ev_geo.record(self.side)

    def step(self, batch, update=True, next_batch=None):
        """batch / next_batch: dicts of DEVICE or pinned-HOST tensors (+ offset_host).  Returns the loss vector."""
        sig = tuple(int(v) for v in batch["offset_host"])
        slots = self._sigs.get(sig, False)
        if slots is False:
            if self._eager_done < self.eager_warmup or not update:
                self._eager_done += 1
                return self._eager_step(batch, update)
            # a capture costs ~1 s: only signatures that come back are worth it (first sighting runs in stream mode)
            self._seen[sig] = self._seen.get(sig, 0) + 1
            if len(self._seen) > 4096:
                self._seen.clear()
            if self._seen[sig] < 2 and any(v is not None for v in self._sigs.values()):
                return self._eager_step(batch, update)
            slots = None
            while sum(v is not None for v in self._sigs.values()) >= self.max_signatures:
                old = next(k for k, v in self._sigs.items() if v is not None)       # evict the least recently used capture
                torch.cuda.synchronize(self.device)                                 # its graphs may still be replaying
                del self._sigs[old]
                torch.cuda.empty_cache()
            if True:
                try:
                    slots = self._capture(batch, sig)
                except Exception as e:                                  # keep training in stream mode
                    self.graph_error = repr(e)
                    try:
                        torch.cuda.synchronize(self.device)
                    except Exception:
                        pass
                    for p in self.model.parameters():
                        p.grad = None
                    self._packed = False
            self._sigs[sig] = slots
        if slots is None or not update:
            return self._eager_step(batch, update)
        self._sigs.move_to_end(sig)
        self._set_packed(True)                                          # the optimiser reads the packed gradient
        cur = next((s for s in slots if s.holds is batch), None)
        if cur is None:
            cur = slots[0]
            self._load(cur, batch)
            self._launch_geo(cur)
        if next_batch is not None and tuple(int(v) for v in next_batch["offset_host"]) == sig:
            other = slots[1] if cur is slots[0] else slots[0]
            self._load(other, next_batch)
            self._launch_geo(other)
        main = torch.cuda.current_stream(self.device)
        main.wait_event(cur.ev_geo)
        if self.hp is not None:
            self.hp.wait_stream(main)
            with torch.cuda.stream(self.hp):
                cur.net.replay()
            main.wait_stream(self.hp)
        else:
            cur.net.replay()
        cur.holds = None
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        self._opt_step()
        return cur.loss.clone()
