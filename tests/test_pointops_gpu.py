"""GPU parity suite (-m gpu): the CUDA path (through the C ABI / drop-in pointops module) against
the CPU oracle on seeded inputs, and against golden vectors produced by the REFERENCE's own CUDA
kernels (tests/golden/pointops_ref_gpu.npz, made by tests/golden/make_golden_gpu.py on the B200
box).  Bit-exact for indices and squared distances; 1e-5 relative for float scatter-adds (the
reference's backward kernels use float atomics, so their own run-to-run order is not fixed)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import oracle  # noqa: E402
from contrastboundary_b200 import pointops, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def golden(golden_dir):
    p = os.path.join(golden_dir, "pointops_ref_gpu.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/pointops_ref_gpu.npz not generated yet")
    return np.load(p)


def run_knn(k, xyz, q, off, qoff):
    pointops.set_knn_cache(0)
    idx, dist = pointops.knn_raw(k, t(xyz), t(q), t(off), t(qoff), sqrt_dist=False)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), dist.cpu().numpy()


@pytest.mark.parametrize("name", list(cases.KNN_CASES))
def test_knn_matches_oracle_and_reference(name, request):
    builder, ks, cross = cases.KNN_CASES[name]
    xyz, off = builder()
    q, qoff = cases.cross_queries(xyz, off) if cross else (xyz, off)
    gp = os.path.join(ROOT, "tests", "golden", "pointops_ref_gpu.npz")
    gold = np.load(gp) if os.path.exists(gp) else None
    for k in ks:
        idx, d2 = run_knn(k, xyz, q, off, qoff)
        oi, od = oracle.knnquery(k, xyz, q, off, qoff)
        assert np.array_equal(idx, oi), f"{name} K={k}: idx differs from oracle in {(idx != oi).any(1).sum()} rows"
        assert np.array_equal(d2.view(np.uint32), od.view(np.uint32)), f"{name} K={k}: dist2 bits differ"
        if gold is not None:
            assert np.array_equal(idx, gold[f"knn/{name}/{k}/idx"]), f"{name} K={k}: idx differs from reference kernel"
            assert np.array_equal(d2.view(np.uint32), gold[f"knn/{name}/{k}/d2"].view(np.uint32))


def test_knn_drop_in_returns_sqrt_and_self_first():
    xyz, off = cases.scene_multi()
    idx, dist = pointops.knnquery(16, t(xyz), t(xyz), t(off), t(off))
    oi, od = oracle.knnquery(16, xyz, None, off, off)
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(od))           # reference pointops.py:43
    assert (idx[:, 0].cpu().numpy() == np.arange(len(xyz))).all()


@pytest.mark.parametrize("n,k", [(40960, 16), (40960, 36), (20000, 8)])
def test_knn_full_size_scene(n, k):
    """BASELINE config-2 scene size; the oracle (brute force, 8 threads) still finishes in seconds."""
    b = synthetic.make_batch(2, [n, n // 2], 77)
    xyz, off = b["points"], b["offset"]
    idx, d2 = run_knn(k, xyz, xyz, off, off)
    oi, od = oracle.knnquery(k, xyz, None, off, off)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_knn_uniform_cube_and_far_queries():
    xyz = synthetic.uniform_cube(30000, 3)
    off = cases.cumsum_i32([30000])
    rng = np.random.default_rng(1)
    q = np.concatenate([xyz[:500], (rng.random((200, 3)) * 30 - 10).astype(np.float32)], 0)   # far outside the bbox too
    qoff = cases.cumsum_i32([700])
    for k in (3, 16, 100):
        idx, d2 = run_knn(k, xyz, q, off, qoff)
        oi, od = oracle.knnquery(k, xyz, q, off, qoff)
        assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_knn_large_k_bruteforce_path():
    xyz, off = cases.scene_multi()
    idx, d2 = run_knn(300, xyz, xyz, off, off)
    oi, od = oracle.knnquery(300, xyz, None, off, off)
    assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("k", [4, 16, 300])
def test_knn_replay_adversarial_point_order(k):
    """Tie replay (knn.cu k_knn_replay) on a support set stored FARTHEST FIRST from the queries at the near end: every
    candidate beats the running K-th distance, so a doubled batch overflows the survivor list and is retried smaller.
    Every point is duplicated, so every query sees exact ties and goes through the replay."""
    n = 30000
    rng = np.random.default_rng(11)
    x = np.sort(rng.random(n // 2).astype(np.float32) * 50)[::-1]
    xyz = np.zeros((n, 3), np.float32)
    xyz[0::2, 0] = x
    xyz[1::2, 0] = x
    xyz[:, 1] = np.repeat(rng.random(n // 2).astype(np.float32) * 0.01, 2)
    off = cases.cumsum_i32([n])
    qsel = np.concatenate([np.arange(0, 64), np.arange(n // 2, n // 2 + 64), np.arange(n - 128, n)])
    q = xyz[qsel] + np.float32(0.0)
    q[1::2, 2] += np.float32(0.5)                      # off-support queries as well
    qoff = cases.cumsum_i32([len(q)])
    idx, d2 = run_knn(k, xyz, q, off, qoff)
    oi, od = oracle.knnquery(k, xyz, q, off, qoff)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_knn_properties_at_microbench_size():
    """size-independent properties at N = 2^18 (no oracle): self first, ascending, in-scene, distances consistent"""
    n, k = 1 << 18, 16
    xyz = synthetic.uniform_cube(n, 9)
    off = cases.cumsum_i32([n // 2, n])[-2:] if False else np.array([n // 2, n], np.int32)
    idx, d2 = run_knn(k, xyz, xyz, off, off)
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (idx[:, 0] == np.arange(n)).all() and (d2[:, 0] == 0).all()
    scene = (np.arange(n) >= n // 2)
    assert ((idx >= n // 2) == scene[:, None]).all()
    sel = np.random.default_rng(0).integers(0, n, 2000)
    diff = xyz[sel][:, None, :].astype(np.float64) - xyz[idx[sel]].astype(np.float64)
    assert np.allclose((diff ** 2).sum(-1), d2[sel], rtol=1e-5, atol=1e-9)
    # exact check of a sample against the oracle restricted to those queries
    qs = np.sort(sel[:300])
    qoff = np.array([(qs < n // 2).sum(), len(qs)], np.int32)
    oi, od = oracle.knnquery(k, xyz, xyz[qs], off, qoff)
    assert np.array_equal(idx[qs], oi)


@pytest.mark.parametrize("name", list(cases.FPS_CASES))
def test_fps_matches_oracle_and_reference(name):
    builder, stride = cases.FPS_CASES[name]
    xyz, off = builder()
    noff = cases.fps_new_offset(off, stride)
    idx = pointops.furthestsampling(t(xyz), t(off), t(noff)).cpu().numpy()
    oi = oracle.furthestsampling(xyz, off, noff)
    assert np.array_equal(idx, oi), f"{name}: {np.flatnonzero(idx != oi)[:5]}"
    gp = os.path.join(ROOT, "tests", "golden", "pointops_ref_gpu.npz")
    if os.path.exists(gp):
        gold = np.load(gp)
        assert np.array_equal(idx, gold[f"fps/{name}/idx"])
        # the running min-distance buffer the reference leaves behind, bit for bit
        from contrastboundary_b200 import _lib as L
        tmp = torch.full((len(xyz),), 1e10, dtype=torch.float32, device=DEV)
        out = torch.zeros(int(noff[-1]), dtype=torch.int32, device=DEV)
        lens = np.diff(np.concatenate([[0], off]))
        L.call("cb_furthest_sampling", len(off), int(lens.max()), t(xyz), t(off), t(noff), tmp, out, L.stream())
        assert np.array_equal(tmp.cpu().numpy().view(np.uint32), gold[f"fps/{name}/tmp"].view(np.uint32))


@pytest.mark.parametrize("lens", [[40960, 40960, 30000, 40960], [10240, 10240, 7000], [2560, 640], [9000, 70000]])
def test_fps_stage_sizes(lens):
    """every kernel configuration (1 CTA / 8-CTA cluster, register- and global-resident)"""
    b = synthetic.make_batch(len(lens), [min(x, 40960) for x in lens], 5)
    pts = []
    prev = 0
    for i, x in enumerate(lens):
        seg = b["points"][prev:prev + min(x, 40960)]
        prev += min(x, 40960)
        if x > 40960:
            seg = np.concatenate([seg, seg[: x - 40960] + np.float32(0.013)], 0)
        pts.append(seg)
    xyz = np.concatenate(pts, 0)
    off = cases.cumsum_i32(lens)
    noff = cases.fps_new_offset(off, 4 if max(lens) <= 40960 else 64)
    oi = oracle.furthestsampling(xyz, off, noff)
    from contrastboundary_b200 import _lib as L
    try:
        # 0 = cluster bucket kernel (distributed shared memory, 8 CTAs x 8 warps), 1 = single-CTA bucket kernel,
        # 2 / 3 / 4 = other cluster shapes
        for mode in (0, 1, 2, 3, 4):
            L.lib().cb_fps_set_mode(mode, 8192)
            if mode == 1 and max(lens) > 49152:
                continue
            idx = pointops.furthestsampling(t(xyz), t(off), t(noff)).cpu().numpy()
            assert np.array_equal(idx, oi), f"mode {mode}: first mismatch at {np.flatnonzero(idx != oi)[:5]}"
    finally:
        L.lib().cb_fps_set_mode(0, 8192)


def test_gather_ops_forward_backward(golden):
    n, k, c, wc, inp, inp2, pos, w, idx, go_nkc, go_nc, wk = cases.ops_inputs()
    close = lambda a, b: np.allclose(a, b, rtol=1e-5, atol=1e-5)  # noqa: E731
    ti = t(idx)
    # grouping
    x = t(inp).requires_grad_(True)
    out = pointops.grouping(x, ti)
    assert np.array_equal(out.detach().cpu().numpy(), oracle.grouping_forward(inp, idx))
    assert np.array_equal(out.detach().cpu().numpy(), golden["ops/grouping_fwd"])
    out.backward(t(go_nkc))
    assert close(x.grad.cpu().numpy(), oracle.grouping_backward(go_nkc, idx, n)) and close(x.grad.cpu().numpy(), golden["ops/grouping_bwd"])
    # subtraction
    a, b2 = t(inp).requires_grad_(True), t(inp2).requires_grad_(True)
    out = pointops.subtraction(a, b2, ti)
    assert np.array_equal(out.detach().cpu().numpy(), golden["ops/subtraction_fwd"])
    out.backward(t(go_nkc))
    assert close(a.grad.cpu().numpy(), golden["ops/subtraction_bwd1"]) and close(b2.grad.cpu().numpy(), golden["ops/subtraction_bwd2"])
    g1, g2 = oracle.subtraction_backward(idx, go_nkc)
    assert close(a.grad.cpu().numpy(), g1) and close(b2.grad.cpu().numpy(), g2)
    # aggregation
    x, p, ww = t(inp).requires_grad_(True), t(pos).requires_grad_(True), t(w).requires_grad_(True)
    out = pointops.aggregation(x, p, ww, ti)
    assert np.array_equal(out.detach().cpu().numpy(), golden["ops/aggregation_fwd"])          # same fma order
    assert np.array_equal(out.detach().cpu().numpy(), oracle.aggregation_forward(inp, pos, w, idx))
    out.backward(t(go_nc))
    assert close(x.grad.cpu().numpy(), golden["ops/aggregation_bwd_i"])
    assert close(p.grad.cpu().numpy(), golden["ops/aggregation_bwd_p"])
    assert close(ww.grad.cpu().numpy(), golden["ops/aggregation_bwd_w"])
    # interpolation kernels
    from contrastboundary_b200.pointops import _InterpolationFn
    idx3 = np.ascontiguousarray(idx[:, :3])
    x = t(inp).requires_grad_(True)
    out = _InterpolationFn.apply(x, t(idx3), t(wk))
    assert np.array_equal(out.detach().cpu().numpy(), golden["ops/interpolation_fwd"])
    out.backward(t(go_nc))
    assert close(x.grad.cpu().numpy(), golden["ops/interpolation_bwd"])
    assert close(x.grad.cpu().numpy(), oracle.interpolation_backward(go_nc, idx3, wk, n))


def test_queryandgroup_and_interpolation_api():
    xyz, off = cases.scene_multi()
    q, qoff = cases.cross_queries(xyz, off)
    rng = np.random.default_rng(2)
    feat = rng.standard_normal((len(xyz), 32)).astype(np.float32)
    g = pointops.queryandgroup(16, t(xyz), t(q), t(feat), None, t(off), t(qoff), use_xyz=True).cpu().numpy()
    oi, _ = oracle.knnquery(16, xyz, q, off, qoff)
    assert g.shape == (len(q), 16, 35)
    assert np.array_equal(g[:, :, 3:], feat[oi])
    assert np.array_equal(g[:, :, :3], xyz[oi] - q[:, None])
    # interpolation: coarse (q) -> fine (xyz), k=3 and k=1   (reference pointops.py:164-178)
    cf = rng.standard_normal((len(q), 8)).astype(np.float32)
    for k in (3, 1):
        out = pointops.interpolation(t(q), t(xyz), t(cf), t(qoff), t(off), k=k).cpu().numpy()
        ii, d2 = oracle.knnquery(k, q, xyz, qoff, off)
        dr = 1.0 / (np.sqrt(d2) + np.float32(1e-8))
        wgt = (dr / dr.sum(1, keepdims=True)).astype(np.float32)
        ref = oracle.interpolation_forward(cf, ii, wgt)
        assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", [-1, 0, 1, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("k,c", [(16, 256), (16, 64), (8, 32), (36, 32), (64, 64), (3, 6)])
def test_fused_knn_gather(k, c, mode):
    from contrastboundary_b200 import fused, _lib
    _lib.lib().cb_knn_gather_set_mode(mode)          # 0 one-warp TMA, 1/3 warp-specialised TMA rings, 4 direct register copy, 5 loader + storer warps, 6/7 LSU copy warps
    xyz, off = cases.scene_multi()
    rng = np.random.default_rng(4)
    feat = rng.standard_normal((len(xyz), c)).astype(np.float32)
    idx, d2, grouped = fused.knn_gather(k, t(xyz), t(xyz), t(feat), t(off), t(off))
    oi, od = oracle.knnquery(k, xyz, None, off, off)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy().view(np.uint32), od.view(np.uint32))
    assert np.array_equal(grouped.cpu().numpy(), feat[oi])
    _lib.lib().cb_knn_gather_set_mode(-1)


def test_fused_knn_gather_ties_and_cross():
    from contrastboundary_b200 import fused
    xyz, off = cases.tie_lattice()
    q, qoff = cases.cross_queries(xyz, off)
    feat = np.random.default_rng(4).standard_normal((len(xyz), 64)).astype(np.float32)
    idx, d2, grouped = fused.knn_gather(8, t(xyz), t(q), t(feat), t(off), t(qoff))
    oi, od = oracle.knnquery(8, xyz, q, off, qoff)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(grouped.cpu().numpy(), feat[oi])
