"""Tall-skinny FP32 linear layers (cb_linear_forward / dgrad / wgrad) for the per-point MLPs.
`Linear` is a drop-in nn.Linear (same parameters / state_dict) that routes large-n, narrow
problems to libcbops and everything else to torch."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib as L

MIN_ROWS = 8192        # forward / dgrad: below this cuBLAS wins (measured: 2048 and 512 are slower; the deep levels have
                       # K, N up to 512-768, where re-staging the split weight tile per k-chunk dominates the custom kernel)
MIN_ROWS_WGRAD = 8192  # wgrad: with few rows the split-K atomics of the custom kernels cost more than cuBLAS' single pass
MAX_CO = 512
MAX_WGRAD = 16384      # ci*co handled by the SIMT wgrad kernel (the tensor-core kernel has no limit)
TENSOR_CORES = True    # 3xTF32 tensor-core kernels (tc_gemm.cu); False = FP32 SIMT kernels (linear_ops.cu)


def set_tensor_cores(flag):
    global TENSOR_CORES
    TENSOR_CORES = bool(flag)
    import ctypes as C
    L.lib().cb_linear_set_tensor_cores(C.c_int(1 if flag else 0))


# ------------------------------------------------------------------------------------------------
# Weight gradients on a side stream.  In the backward of y = x W^T + b, dW = g^T x / db = sum(g) feed nothing but the
# optimizer, while dx = g W is on the critical path of everything upstream.  engine.TrainStep wraps its backward in
# `wgrad_fork()`: every linear layer then launches its weight-gradient kernels on ONE side stream (forked after g is ready),
# and the engine joins that stream once, after the backward and before the gradients are packed.  Inside a CUDA-graph capture
# the fork / join become graph edges: ~70 small kernels per step leave the serial chain of the network graph and overlap with
# it.  Off by default: a caller who runs loss.backward() without the join must not see gradients that are still in flight.
# ------------------------------------------------------------------------------------------------
_fork = {"on": False, "streams": {}, "used": set()}


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _fork["streams"].get(key)
    if st is None:
        st = _fork["streams"][key] = torch.cuda.Stream(device=key)
    return st


class wgrad_fork:
    """with wgrad_fork(): loss.backward()   -- weight gradients of the linear layers run on a side stream; the exit joins it."""

    def __enter__(self):
        self.prev = _fork["on"]
        _fork["on"] = True
        return self

    def __exit__(self, *exc):
        _fork["on"] = self.prev
        join_wgrad()
        return False


def join_wgrad():
    """The current stream waits for every weight-gradient kernel forked since the last join."""
    for key in list(_fork["used"]):
        torch.cuda.current_stream(key).wait_stream(_fork["streams"][key])
    _fork["used"].clear()


class _SkinnyLinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        n, ci = x.shape
        co = weight.shape[0]
        ctx.custom = n >= MIN_ROWS and co <= MAX_CO
        if ctx.custom:
            y = torch.empty((n, co), dtype=torch.float32, device=x.device)
            L.call("cb_linear_forward", n, ci, co, x, weight, bias, y, L.stream())
        else:
            y = F.linear(x, weight, bias)              # few rows: cuBLAS (see MIN_ROWS)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        # the fork is only safe when dW / db go straight to AccumulateGrad (which adopts the tensor without launching a kernel
        # while p.grad is None); a weight that is itself computed (e.g. a slice of a parameter) has more backward ahead of it
        ctx.leaf_params = weight.is_leaf and (bias is None or bias.is_leaf)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        n, ci = x.shape
        co = weight.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if ctx.custom:
                gx = torch.empty_like(x)
                L.call("cb_linear_dgrad", n, ci, co, g, weight, gx, L.stream())
            else:
                gx = g.mm(weight)
        want_w, want_b = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if want_w or want_b:
            side = None
            if _fork["on"] and g.is_cuda and ctx.leaf_params and weight.grad is None:
                side = _side_stream(g.device)
                side.wait_stream(torch.cuda.current_stream(g.device))          # g is ready; x, weight long since
            import contextlib
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                if want_w and n >= MIN_ROWS_WGRAD and co <= MAX_CO and (TENSOR_CORES or (ci * co <= MAX_WGRAD and co <= 256)):
                    gw = torch.empty_like(weight)
                    gb = torch.empty(co, dtype=torch.float32, device=x.device) if want_b else None
                    L.call("cb_linear_wgrad", n, ci, co, x, g, gw, gb, L.stream())
                else:
                    if want_w:
                        gw = g.t().mm(x)
                    if want_b:
                        gb = g.sum(0)
            if side is not None:
                x.record_stream(side)
                g.record_stream(side)
                _fork["used"].add(side.device.index)
        return gx, gw, gb


def fast_linear(x, weight, bias=None):
    if x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_contiguous() and x.dim() >= 2:
        shp = x.shape
        y = _SkinnyLinearFn.apply(x.reshape(-1, shp[-1]).contiguous(), weight, bias)
        return y.view(*shp[:-1], weight.shape[0])
    return F.linear(x, weight, bias)


class Linear(nn.Linear):
    def forward(self, x):
        return fast_linear(x, self.weight, self.bias)


# ------------------------------------------------------------------------------------------------
# BatchNorm1d with a deferred batch counter.  nn.BatchNorm1d increments `num_batches_tracked` with one tiny
# kernel per layer and forward (123 of them per step of this network).  This subclass (same parameters, buffers
# and state_dict) queues the counters instead; `flush_bn_counters()` bumps all of them with ONE multi-tensor op.
# The counter only matters for momentum=None (cumulative average), which the reference never uses
# (pytorch/model/blocks.py: nn.BatchNorm1d defaults, momentum 0.1).
# ------------------------------------------------------------------------------------------------
_bn_pending = []


def bump_bn_counter(bn):
    if bn.num_batches_tracked is not None:
        _bn_pending.append(bn.num_batches_tracked)
        if len(_bn_pending) >= 1024:          # stand-alone use of the layers (no network forward to flush them)
            flush_bn_counters()


def flush_bn_counters():
    if _bn_pending:
        torch._foreach_add_(list(_bn_pending), 1)
        _bn_pending.clear()


class BatchNorm1d(nn.BatchNorm1d):
    def forward(self, x):
        if self.momentum is None or not self.track_running_stats or not x.is_cuda:
            return super().forward(x)
        if self.training:
            bump_bn_counter(self)
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training,
                            self.momentum, self.eps)


# ------------------------------------------------------------------------------------------------
# fused BatchNorm (+ residual) (+ ReLU): libcbops' cb_bn_act_forward / cb_bn_act_backward (bn_ops.cu)
# ------------------------------------------------------------------------------------------------
FUSED_BN = True
PERSISTENT_BN_ACC = True      # per-module self-cleaning statistics accumulator instead of a memset per BatchNorm call (bn_ops.cu)


class _BnActFn(Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, momentum, eps, training, relu, acc=None):
        import ctypes as C
        x = x.contiguous()
        n, c = x.shape
        y = torch.empty_like(x)
        bnbuf = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        # acc: the module's persistent, self-cleaning accumulator (2c doubles + ticket; zero between uses) -> no memset
        persistent = acc is not None
        stats = acc if persistent else torch.empty(2 * c, dtype=torch.float64, device=x.device)
        res = residual.contiguous() if residual is not None else None
        rc = L.lib().cb_bn_act_forward(C.c_longlong(n), C.c_int(c), L.ptr(x), L.ptr(res), L.ptr(gamma), L.ptr(beta),
                                       L.ptr(running_mean), L.ptr(running_var), C.c_float(momentum), C.c_float(eps),
                                       C.c_int((1 if training else 0) | (2 if persistent else 0)), C.c_int(1 if relu else 0),
                                       L.ptr(y), L.ptr(bnbuf), L.ptr(stats), L.stream())
        L.check(rc, "cb_bn_act_forward")
        ctx.save_for_backward(x, y, gamma, bnbuf)
        ctx.cfg = (bool(training), bool(relu), residual is not None, stats, persistent)
        return y

    @staticmethod
    def backward(ctx, gy):
        import ctypes as C
        x, y, gamma, bnbuf = ctx.saved_tensors
        training, relu, has_res, stats, persistent = ctx.cfg
        n, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        gres = torch.empty_like(x) if has_res else None
        gg = torch.empty(c, dtype=torch.float32, device=x.device)
        gb = torch.empty(c, dtype=torch.float32, device=x.device)
        rc = L.lib().cb_bn_act_backward(C.c_longlong(n), C.c_int(c), L.ptr(x), L.ptr(y), L.ptr(gamma), L.ptr(bnbuf),
                                        C.c_int((1 if training else 0) | (2 if persistent else 0)), C.c_int(1 if relu else 0),
                                        L.ptr(gy), L.ptr(gx), L.ptr(gres), L.ptr(gg), L.ptr(gb), L.ptr(stats), L.stream())
        L.check(rc, "cb_bn_act_backward")
        return gx, gres, gg, gb, None, None, None, None, None, None, None


def bn_act(bn, x, residual=None, relu=True):
    """act(bn(x) [+ residual]) for a BatchNorm1d module `bn` on an (n, c) tensor: the reference's
    relu(bn(.)) / relu(bn3(.) + identity) patterns (blocks.py:76,127-133) in two kernels per direction."""
    if (FUSED_BN and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[1] % 4 == 0 and x.shape[1] <= 1024
            and bn.track_running_stats and bn.momentum is not None and bn.affine and x.shape[0] > 0
            and not isinstance(bn, nn.SyncBatchNorm)):
        if bn.training:
            bump_bn_counter(bn)
        acc = None
        if PERSISTENT_BN_ACC:
            acc = bn.__dict__.get("_cb_acc")           # plain attribute: not a registered buffer, not in the state_dict
            if acc is None or acc.device != x.device:
                acc = bn.__dict__["_cb_acc"] = torch.zeros(2 * x.shape[1] + 2, dtype=torch.float64, device=x.device)
        return _BnActFn.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum),
                              float(bn.eps), bn.training, relu, acc)
    y = bn(x)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y
