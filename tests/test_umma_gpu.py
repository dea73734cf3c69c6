"""tcgen05 / TMEM linear layers (csrc/umma_linear.cu: tcgen05.mma kind::tf32 with 3xTF32 compensation, accumulator in tensor
memory) against float64: the same 1e-5 bar as the mma.sync and SIMT kernels they replace (tests/test_ptlayer_gpu.py)."""
import numpy as np
import pytest
import torch

from contrastboundary_b200 import _lib as L

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 32, 96), (150000, 32, 96), (60000, 72, 144), (40960, 64, 192), (163840, 32, 32), (5000, 128, 384), (777, 8, 16), (300, 40, 48), (129, 256, 256),
          (8192, 512, 512), (1, 16, 16)]


@pytest.fixture(params=[3, 2, 1], ids=["pipelined_cp_async", "pipelined", "simple"])
def version(request):
    import ctypes as C
    L.lib().cb_linear_set_umma_version(C.c_int(request.param))
    yield request.param
    L.lib().cb_linear_set_umma_version(C.c_int(3))


@pytest.mark.parametrize("n,ci,co", SHAPES)
def test_umma_forward_matches_float64(n, ci, co, version):
    g = torch.Generator(device="cuda").manual_seed(n + ci)
    x = torch.randn(n, ci, device="cuda", generator=g)
    w = torch.randn(co, ci, device="cuda", generator=g) / ci ** 0.5
    b = torch.randn(co, device="cuda", generator=g)
    y = torch.full((n, co), float("nan"), device="cuda")
    L.call("cb_umma_linear_forward", n, ci, co, x, w, b, y, L.stream())
    ref = x.double() @ w.double().t() + b.double()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
    y2 = torch.empty_like(y)
    L.call("cb_umma_linear_forward", n, ci, co, x, w, None, y2, L.stream())
    assert float((y2.double() - (ref - b.double())).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("n,ci,co", SHAPES)
def test_umma_dgrad_matches_float64(n, ci, co, version):
    if co % 8 or ci % 16:
        pytest.skip("dgrad needs co % 8 == 0 and ci % 16 == 0")
    g = torch.Generator(device="cuda").manual_seed(n + co)
    gy = torch.randn(n, co, device="cuda", generator=g)
    w = torch.randn(co, ci, device="cuda", generator=g) / co ** 0.5
    dx = torch.full((n, ci), float("nan"), device="cuda")
    L.call("cb_umma_linear_dgrad", n, ci, co, gy, w, dx, L.stream())
    ref = gy.double() @ w.double()
    assert float((dx.double() - ref).abs().max() / ref.abs().max()) < 1e-5
