"""GPU parity of the fused PointTransformer layer kernels (cb_pt_layer_forward/backward) against the
op-by-op torch path of the same module (which mirrors blocks.py:31-44), forward and every gradient."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


def make_level(n_list, k, seed):
    from contrastboundary_b200 import model, pointops, ptlayer, synthetic
    b = synthetic.make_batch(len(n_list), n_list, seed)
    lv = model.Level()
    lv.p = torch.from_numpy(b["points"]).cuda()
    lv.o = torch.from_numpy(b["offset"]).cuda()
    lv.n = lv.p.shape[0]
    lv.knn, _ = pointops.knn_raw(k, lv.p, lv.p, lv.o, lv.o, True)
    lv.rel, lv.rel_mom = ptlayer.pt_rel(lv.p, lv.knn)
    return lv


def rel_err(a, b):
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp(min=1e-12))


def outlier_fraction(a, b, tol):
    """fraction of entries off by more than tol * max|b|.  The fused backward RE-COMPUTES the pre-ReLU
    activations; an activation within one ulp of zero can land on the other side of the ReLU than in
    torch's stored copy, which flips one element's subgradient (both are valid).  Such flips touch a
    handful of entries, so gradients are compared by the fraction of entries that disagree."""
    d = (a.detach() - b.detach()).abs() > tol * b.detach().abs().max().clamp(min=1e-12)
    if d.dim() > 1:
        d = d.reshape(d.shape[0], -1).any(1)      # count ROWS: one flipped (row, channel) spreads over the row through W
    return float(d.float().mean())


@pytest.mark.parametrize("c,k,n_list", [(32, 8, [3000, 2000]), (64, 16, [1500, 900]), (128, 16, [700, 500]),
                                        (256, 16, [300, 200]), (512, 16, [90, 70]), (32, 16, [12, 700])])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("tensor_cores", [True, False])
def test_fused_layer_matches_unfused(c, k, n_list, training, tensor_cores):
    """tensor_cores: the (n*k) x c x c/8 contraction on mma.sync with 3xTF32 (ptlayer_mma.cu) or FP32 SIMT"""
    from contrastboundary_b200 import ptlayer
    ptlayer.set_tensor_cores(tensor_cores)
    try:
        _fused_layer_case(c, k, n_list, training)
    finally:
        ptlayer.set_tensor_cores(True)


def _fused_layer_case(c, k, n_list, training):
    from contrastboundary_b200 import model
    lv = make_level(n_list, k, 100 + c)
    torch.manual_seed(c + k)
    layer = model.PointTransformerLayer(c, c, 8, k).cuda()
    cases.deterministic_init(layer, 3)
    layer.train(training)
    x = torch.randn(lv.n, c, device="cuda")
    gout = torch.randn(lv.n, c, device="cuda")
    res = {}
    for fused in (False, True):
        layer.fused = fused
        for bn in (layer.linear_p[1], layer.linear_w[0], layer.linear_w[3]):   # same running stats for both runs
            bn.running_mean.zero_().add_(0.05); bn.running_var.fill_(1.3)
        layer.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        out = layer(lv, xi)
        out.backward(gout)
        torch.cuda.synchronize()
        res[fused] = (out.detach(), xi.grad.detach(), {n: p.grad.detach().clone() for n, p in layer.named_parameters()},
                      {n: b2.detach().clone() for n, b2 in layer.named_buffers() if b2.dtype.is_floating_point})
    o0, gx0, gp0, bf0 = res[False]
    o1, gx1, gp1, bf1 = res[True]
    assert rel_err(o1, o0) < 2e-5, f"out {rel_err(o1, o0)}"
    assert outlier_fraction(gx1, gx0, 2e-4) < 2e-2, f"grad x {rel_err(gx1, gx0)} {outlier_fraction(gx1, gx0, 2e-4)}"
    errs = []
    for name in gp0:
        scale = gp0[name].abs().max()
        if cases.grad_is_analytically_zero("transformer2." + name) or scale < 1e-6:
            continue
        e = rel_err(gp1[name], gp0[name])
        errs.append(e)
        assert e < 5e-2, f"grad {name} {e}"            # a single ReLU flip moves a reduced gradient by ~1/rows
    assert sorted(errs)[len(errs) // 2] < 1e-3, errs    # ... but the typical parameter gradient agrees to 1e-4
    if training:
        for name in bf0:
            assert rel_err(bf1[name], bf0[name]) < 1e-4, f"buffer {name}"


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("n,ci,co,bias", [(20000, 32, 32, True), (163840, 6, 32, False), (50000, 35, 64, False),
                                         (16384, 67, 128, False), (30000, 160, 13, True), (9000, 131, 256, False),
                                         (40960, 64, 192, True), (10240, 128, 384, True), (8200, 256, 96, False)])
def test_skinny_linear(n, ci, co, bias, tensor_cores):
    """tall-skinny linear layers vs float64: the 3xTF32 tensor-core kernels must hold the same FP32-level error
    as the exact-FP32 SIMT kernels"""
    from contrastboundary_b200 import linear_ops
    linear_ops.set_tensor_cores(tensor_cores)
    try:
        _skinny_linear_case(n, ci, co, bias)
    finally:
        linear_ops.set_tensor_cores(True)


def _skinny_linear_case(n, ci, co, bias):
    from contrastboundary_b200 import linear_ops
    torch.manual_seed(n % 97)
    x = torch.randn(n, ci, device="cuda", requires_grad=True)
    w = (torch.randn(co, ci, device="cuda") / ci ** 0.5).requires_grad_(True)
    b = torch.randn(co, device="cuda", requires_grad=True) if bias else None
    g = torch.randn(n, co, device="cuda")
    y = linear_ops.fast_linear(x, w, b)
    y.backward(g)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone() if bias else None)
    x.grad = None; w.grad = None
    if bias:
        b.grad = None
    yr = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    yr.backward(g.double())
    assert rel_err(got[0], yr.float()) < 1e-5
    assert rel_err(got[1], x.grad.float()) < 1e-5
    assert rel_err(got[2], w.grad.float()) < 1e-4
    if bias:
        assert rel_err(got[3], b.grad.float()) < 1e-4


@pytest.mark.parametrize("c_in,c_out,n_list", [(32, 64, [4000, 2500]), (64, 128, [1200, 800]), (256, 512, [300, 200])])
@pytest.mark.parametrize("training", [True, False])
def test_fused_transition_down(c_in, c_out, n_list, training):
    """fused TransitionDown (linear-before-gather + BN + ReLU + max) vs the op-by-op path (blocks.py:69-73)"""
    from contrastboundary_b200 import model, pointops, ptlayer, synthetic
    b = synthetic.make_batch(len(n_list), n_list, 77)
    prev, lv = model.Level(), model.Level()
    prev.p = torch.from_numpy(b["points"]).cuda(); prev.o = torch.from_numpy(b["offset"]).cuda()
    new_lens = [x // 4 for x in n_list]
    lv.o = torch.tensor(np.cumsum(new_lens), dtype=torch.int32, device="cuda")
    fidx = pointops.furthestsampling_known(prev.p, prev.o, lv.o, max(n_list), sum(new_lens))
    lv.p = prev.p[fidx.long()].contiguous()
    lv.down_idx, _ = pointops.knn_raw(16, prev.p, lv.p, prev.o, lv.o, True)
    lv.rel_down = ptlayer.td_rel(prev.p, lv.p, lv.down_idx)
    td = model.TransitionDown(c_in, c_out, 4, 16).cuda()
    cases.deterministic_init(td, 5)
    td.train(training)
    torch.manual_seed(3)
    x = torch.randn(prev.p.shape[0], c_in, device="cuda")
    g = torch.randn(lv.p.shape[0], c_out, device="cuda")
    res = {}
    for fused in (False, True):
        td.fused = fused
        td.bn.running_mean.zero_().add_(0.05); td.bn.running_var.fill_(1.3)
        td.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        out = td(xi, prev, lv)
        out.backward(g)
        res[fused] = (out.detach(), xi.grad.clone(), td.linear.weight.grad.clone(), td.bn.weight.grad.clone(), td.bn.bias.grad.clone(),
                      td.bn.running_mean.clone(), td.bn.running_var.clone())
    a, r = res[True], res[False]
    assert rel_err(a[0], r[0]) < 2e-5
    assert outlier_fraction(a[1], r[1], 2e-4) < 2e-2, rel_err(a[1], r[1])
    assert rel_err(a[2], r[2]) < 2e-2 and rel_err(a[3], r[3]) < 2e-2 and rel_err(a[4], r[4]) < 2e-2
    if training:
        assert rel_err(a[5], r[5]) < 1e-4 and rel_err(a[6], r[6]) < 1e-4


@pytest.mark.parametrize("n,c", [(163840, 32), (40960, 64), (2560, 256), (640, 512), (7, 32)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("residual,relu", [(False, True), (True, True), (False, False)])
def test_fused_bn_act(n, c, training, residual, relu):
    """fused BatchNorm (+ residual) (+ ReLU) kernels vs torch.nn.BatchNorm1d + add + relu (blocks.py:127-133)"""
    from contrastboundary_b200 import linear_ops
    torch.manual_seed(c + n % 13)
    x0 = torch.randn(n, c, device="cuda") * 1.7 + 0.3
    r0 = torch.randn(n, c, device="cuda") if residual else None
    g = torch.randn(n, c, device="cuda")
    res = {}
    for fused in (False, True):
        bn = linear_ops.BatchNorm1d(c).cuda()
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, c)); bn.bias.copy_(torch.linspace(-0.2, 0.2, c))
            bn.running_mean.fill_(0.1); bn.running_var.fill_(1.2)
        bn.train(training)
        x = x0.clone().requires_grad_(True)
        r = r0.clone().requires_grad_(True) if residual else None
        linear_ops.FUSED_BN = fused
        try:
            y = linear_ops.bn_act(bn, x, residual=r, relu=relu)
            y.backward(g)
        finally:
            linear_ops.FUSED_BN = True
        res[fused] = (y.detach(), x.grad.clone(), r.grad.clone() if residual else None, bn.weight.grad.clone(), bn.bias.grad.clone(),
                      bn.running_mean.clone(), bn.running_var.clone())
    a, b = res[False], res[True]
    assert rel_err(b[0], a[0]) < 2e-5
    assert rel_err(b[1], a[1]) < 2e-4
    if residual:
        assert torch.equal(b[2], a[2])
    assert rel_err(b[3], a[3]) < 2e-4 and rel_err(b[4], a[4]) < 2e-4
    assert rel_err(b[5], a[5]) < 1e-5 and rel_err(b[6], a[6]) < 1e-5


@pytest.mark.parametrize("n,c", [(40960, 64), (300, 512), (5, 32)])
def test_fused_bn_act_self_cleaning_accumulator(n, c):
    """The module's persistent statistics accumulator (bn_ops.cu: the last block of the apply kernel zeroes it) must be all-zero
    after every forward and every backward, so that repeated calls — eager or as CUDA-graph replays — give what the
    memset-per-call path gives."""
    from contrastboundary_b200 import linear_ops
    torch.manual_seed(n)
    x0 = torch.randn(n, c, device="cuda") * 2 + 1
    g = torch.randn(n, c, device="cuda")
    outs = {}
    for persistent in (False, True):
        linear_ops.PERSISTENT_BN_ACC = persistent
        try:
            bn = linear_ops.BatchNorm1d(c).cuda().train()
            runs = []
            for it in range(3):
                x = x0.clone().requires_grad_(True)
                y = linear_ops.bn_act(bn, x, relu=True)
                if persistent:
                    assert float(bn.__dict__["_cb_acc"].abs().max()) == 0.0 and int(bn.__dict__["_cb_acc"][2 * c:].view(torch.int32)[0]) == 0
                y.backward(g)
                if persistent:
                    assert float(bn.__dict__["_cb_acc"].abs().max()) == 0.0
                runs.append((y.detach().clone(), x.grad.clone(), bn.weight.grad.clone()))
                bn.weight.grad = None
                bn.bias.grad = None
            for r in runs[1:]:
                assert rel_err(r[0], runs[0][0]) < 1e-6 and rel_err(r[1], runs[0][1]) < 1e-5 and rel_err(r[2], runs[0][2]) < 1e-5
            outs[persistent] = runs[0]
        finally:
            linear_ops.PERSISTENT_BN_ACC = True
    assert rel_err(outs[True][0], outs[False][0]) < 1e-6 and rel_err(outs[True][1], outs[False][1]) < 1e-5
    # the same through a CUDA graph: capture once, replay three times
    bn = linear_ops.BatchNorm1d(c).cuda().train()
    xs = x0.clone().requires_grad_(True)
    linear_ops.bn_act(bn, xs, relu=True).backward(g)          # warm-up outside the capture (allocates the accumulator)
    xs.grad = None
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr):
            ys = linear_ops.bn_act(bn, xs, relu=True)
            ys.backward(g)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    assert rel_err(ys, outs[False][0]) < 1e-6 and rel_err(xs.grad, outs[False][1]) < 1e-5
    assert float(bn.__dict__["_cb_acc"].abs().max()) == 0.0
