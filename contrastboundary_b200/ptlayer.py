"""Fused PointTransformer local aggregation (reference: PointTransformerLayer.forward,
pytorch/model/blocks.py:31-44) as one autograd Function over libcbops' cb_pt_layer_forward /
cb_pt_layer_backward.  Parameters stay ordinary nn.Parameters of the layer module (state_dict
compatible with the reference); BatchNorm running statistics are updated in-kernel."""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L
from .linear_ops import bump_bn_counter

READY = True
_FP = C.c_void_p
TENSOR_CORES = True     # the (n*k) x c x c/8 contraction on the tensor cores (3xTF32), forward and backward


def set_tensor_cores(flag):
    """True (default): mma.sync 3xTF32 kernels; the forward keeps w0 (n,k,c) for the backward GEMMs.
    False: FP32 SIMT kernels that recompute every (n,k,c) quantity from gathers."""
    global TENSOR_CORES
    TENSOR_CORES = bool(flag)
    L.lib().cb_pt_set_tensor_cores(C.c_int(1 if flag else 0))


class CbPtLayer(C.Structure):
    _fields_ = [(n, _FP) for n in (
        "w1", "b1", "bn1_weight", "bn1_bias", "bn1_running_mean", "bn1_running_var",
        "w2", "b2", "bn2_weight", "bn2_bias", "bn2_running_mean", "bn2_running_var",
        "w3", "b3", "bn3_weight", "bn3_bias", "bn3_running_mean", "bn3_running_var",
        "w4", "b4")] + [("momentum", C.c_float), ("eps", C.c_float), ("training", C.c_int)]


_size_cache = {}


def _sizes(lib, c):
    """(bnbuf floats, stats doubles) of a layer with c channels — constants of the library, asked once per width"""
    v = _size_cache.get(c)
    if v is None:
        v = _size_cache[c] = (int(lib.cb_pt_bnbuf_floats(C.c_int(c))), int(lib.cb_pt_stats_doubles(C.c_int(c))))
    return v


_struct_cache = {}


def _param_struct(params, buffers, momentum, eps, training):
    # the struct holds 20 raw pointers; it only changes when a tensor is re-allocated, so it is cached per layer (identity of
    # its first parameter, kept alive by the cache entry) and validated against the current addresses of all its tensors
    key = (id(params[0]), id(buffers[0]), float(momentum), float(eps), int(training))
    addrs = tuple(t.data_ptr() for t in params) + tuple(t.data_ptr() for t in buffers)
    hit = _struct_cache.get(key)
    if hit is not None and hit[1] == addrs and hit[2] is params[0]:
        return hit[0]
    s = _param_struct_build(params, buffers, momentum, eps, training)
    if len(_struct_cache) > 4096:
        _struct_cache.clear()
    _struct_cache[key] = (s, addrs, params[0])
    return s


def _param_struct_build(params, buffers, momentum, eps, training):
    (w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, g3, be3, w4, b4) = params
    (rm1, rv1, rm2, rv2, rm3, rv3) = buffers
    s = CbPtLayer()
    for name, t in (("w1", w1), ("b1", b1), ("bn1_weight", g1), ("bn1_bias", be1), ("bn1_running_mean", rm1),
                    ("bn1_running_var", rv1), ("w2", w2), ("b2", b2), ("bn2_weight", g2), ("bn2_bias", be2),
                    ("bn2_running_mean", rm2), ("bn2_running_var", rv2), ("w3", w3), ("b3", b3), ("bn3_weight", g3),
                    ("bn3_bias", be3), ("bn3_running_mean", rm3), ("bn3_running_var", rv3), ("w4", w4), ("b4", b4)):
        assert t.is_contiguous() and t.dtype == torch.float32
        setattr(s, name, t.data_ptr())
    s.momentum, s.eps, s.training = float(momentum), float(eps), int(training)
    return s


class PtAttentionFn(Function):
    @staticmethod
    def forward(ctx, rel, moments, idx, qkv, buffers, momentum, eps, training, *params):
        # qkv (n, 3c): the fused q/k/v projection; x_q, x_k, x_v are its column blocks (row stride ld = 3c)
        n, k = idx.shape
        c = qkv.shape[1] // 3
        cs = c // 8
        dev = qkv.device
        qkv = qkv.contiguous()
        ld = 3 * c
        xq, xk, xv = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
        w2buf = torch.empty((n, k, cs), dtype=torch.float32, device=dev)
        abuf = torch.empty((n, k, cs), dtype=torch.float32, device=dev)
        lib = L.lib()
        bnbuf = torch.empty(_sizes(lib, c)[0], dtype=torch.float32, device=dev)
        stats = torch.empty(_sizes(lib, c)[1], dtype=torch.float64, device=dev)
        ps = _param_struct(params, buffers, momentum, eps, training)
        # tensor-core backward: keep the pre-BatchNorm activation w0 (n,k,c) instead of re-gathering it three times
        keep_w0 = bool(training) and TENSOR_CORES and any(ctx.needs_input_grad)
        w0buf = torch.empty((n, k, c), dtype=torch.float32, device=dev) if keep_w0 else None
        rc = lib.cb_pt_layer_forward(C.c_int(n), C.c_int(k), C.c_int(c), C.c_int(ld), C.byref(ps), L.ptr(rel), L.ptr(moments), L.ptr(idx),
                                     L.ptr(xq), L.ptr(xk), L.ptr(xv), L.ptr(out), L.ptr(w2buf), L.ptr(abuf), L.ptr(bnbuf),
                                     L.ptr(stats), L.ptr(w0buf), L.stream())
        L.check(rc, "cb_pt_layer_forward")
        ctx.has_w0 = w0buf is not None
        ctx.save_for_backward(rel, idx, qkv, w2buf, abuf, bnbuf, *params, *((w0buf,) if w0buf is not None else ()))
        ctx.training = training
        return out

    @staticmethod
    def backward(ctx, gout):
        rel, idx, qkv, w2buf, abuf, bnbuf = ctx.saved_tensors[:6]
        params = ctx.saved_tensors[6:20]
        w0buf = ctx.saved_tensors[20] if ctx.has_w0 else None
        n, k = idx.shape
        c = qkv.shape[1] // 3
        cs = c // 8
        dev = qkv.device
        ld = 3 * c
        xq, xk, xv = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
        gout = gout.contiguous()
        gqkv = torch.zeros_like(qkv)                  # x_k / x_v blocks are scatter-add targets, x_q block is overwritten
        gxq, gxk, gxv = gqkv[:, :c], gqkv[:, c:2 * c], gqkv[:, 2 * c:]
        lib = L.lib()
        scratch = torch.empty(lib.cb_pt_bwd_scratch_floats(C.c_int(n), C.c_int(k), C.c_int(c)), dtype=torch.float32, device=dev)
        # parameter gradients in one zeroed buffer (layout of cb_pt_layer_backward)
        sizes = [9, 3, 3, 3, c * 3, c, c, c, cs * c, cs, cs, cs, cs * cs, cs]
        gbuf = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        ps = _param_struct(params, (params[0],) * 6, 0.0, 0.0, ctx.training)   # running stats unused in backward
        rc = lib.cb_pt_layer_backward(C.c_int(n), C.c_int(k), C.c_int(c), C.c_int(ld), C.byref(ps), L.ptr(rel), L.ptr(idx), L.ptr(xq),
                                      L.ptr(xk), L.ptr(xv), L.ptr(w2buf), L.ptr(abuf), L.ptr(bnbuf), L.ptr(gout),
                                      L.ptr(gxq), L.ptr(gxk), L.ptr(gxv), L.ptr(gbuf), L.ptr(scratch), L.ptr(w0buf), L.stream())
        L.check(rc, "cb_pt_layer_backward")
        grads, o = [], 0
        for sz, p in zip(sizes, params):
            grads.append(gbuf[o:o + sz].view_as(p))
            o += sz
        return (None, None, None, gqkv, None, None, None, None, *grads)


def pt_attention(layer, lv, qkv):
    lp, lw = layer.linear_p, layer.linear_w
    params = (lp[0].weight, lp[0].bias, lp[1].weight, lp[1].bias, lp[3].weight, lp[3].bias,
              lw[0].weight, lw[0].bias, lw[2].weight, lw[2].bias, lw[3].weight, lw[3].bias, lw[5].weight, lw[5].bias)
    buffers = (lp[1].running_mean, lp[1].running_var, lw[0].running_mean, lw[0].running_var,
               lw[3].running_mean, lw[3].running_var)
    training = layer.training
    if training:
        for bn in (lp[1], lw[0], lw[3]):
            bump_bn_counter(bn)
    return PtAttentionFn.apply(lv.rel, lv.rel_mom, lv.knn, qkv, buffers, lp[1].momentum, lp[1].eps,
                               training, *params)


def pt_rel(p, idx):
    """rel (n,k,3) = p[idx] - p[n], moments (9) float64 — geometry only, once per level"""
    n, k = idx.shape
    rel = torch.empty((n, k, 3), dtype=torch.float32, device=p.device)
    mom = torch.empty(9, dtype=torch.float64, device=p.device)
    L.call("cb_pt_rel", n, k, p, idx, rel, mom, L.stream())
    return rel, mom


# ------------------------------------------------------------------------------------------------
# fused TransitionDown (blocks.py:69-73): out = max_k relu(bn(Wxyz rel + z[idx])),  z = x Wf^T
# ------------------------------------------------------------------------------------------------
class TransitionDownFn(Function):
    @staticmethod
    def forward(ctx, rel, idx, z, wxyz, gamma, beta, rm, rv, momentum, eps, training):
        m, k = idx.shape
        c = z.shape[1]
        dev = z.device
        z, wxyz = z.contiguous(), wxyz.contiguous()
        out = torch.empty((m, c), dtype=torch.float32, device=dev)
        argk = torch.empty((m, c), dtype=torch.uint8, device=dev)
        bnbuf = torch.empty(4 * c, dtype=torch.float32, device=dev)
        stats = torch.empty(2 * c, dtype=torch.float64, device=dev)
        L.call("cb_td_forward", m, k, c, rel, idx, z, wxyz, gamma, beta, rm, rv, float(momentum), float(eps), int(training), out,
               argk, bnbuf, stats, L.stream())
        ctx.save_for_backward(rel, idx, z, wxyz, gamma, bnbuf, out, argk)
        ctx.training = int(training)
        return out

    @staticmethod
    def backward(ctx, g):
        rel, idx, z, wxyz, gamma, bnbuf, out, argk = ctx.saved_tensors
        m, k = idx.shape
        c = z.shape[1]
        dev = z.device
        gz = torch.zeros_like(z)
        gw = torch.zeros_like(wxyz)
        gg = torch.empty(c, dtype=torch.float32, device=dev)
        gb = torch.empty(c, dtype=torch.float32, device=dev)
        scratch = torch.empty(7 * c + 16, dtype=torch.float32, device=dev)
        L.call("cb_td_backward", m, k, c, rel, idx, z, wxyz, gamma, ctx.training, bnbuf, out, argk, g.contiguous(), gz, gw, gg, gb,
               scratch, L.stream())
        return None, None, gz, gw, gg, gb, None, None, None, None, None


def transition_down(td, x, prev_level, level):
    """fused body of TransitionDown.forward for stride > 1"""
    from .linear_ops import fast_linear
    w = td.linear.weight
    z = fast_linear(x, w[:, 3:].contiguous())
    bn = td.bn
    if td.training:
        bump_bn_counter(bn)
    return TransitionDownFn.apply(level.rel_down, level.down_idx, z, w[:, :3], bn.weight, bn.bias, bn.running_mean,
                                  bn.running_var, bn.momentum, bn.eps, td.training)


def td_rel(p_support, p_query, idx):
    m, k = idx.shape
    rel = torch.empty((m, k, 3), dtype=torch.float32, device=p_support.device)
    L.call("cb_td_rel", m, k, p_support, p_query, idx, rel, L.stream())
    return rel
