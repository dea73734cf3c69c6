"""Generate tests/golden/pointops_ref_gpu.npz from the REFERENCE's own CUDA kernels.

Runs on the B200 box only (needs a GPU and oracle/_ref/pointops_cuda.so, which
oracle/build_ref.sh compiles unmodified from /root/reference in the build container):

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/pointops_ref_gpu.npz'

then copy the file into tests/golden/.  Inputs are regenerated from tests/golden/cases.py seeds,
so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import oracle  # noqa: E402


def main(out):
    ref = oracle.ref_pointops_cuda()
    dev = torch.device("cuda")
    res = {}
    for name, (builder, ks, cross) in cases.KNN_CASES.items():
        xyz, off = builder()
        q, qoff = cases.cross_queries(xyz, off) if cross else (xyz, off)
        txyz, tq = torch.from_numpy(xyz).to(dev), torch.from_numpy(q).to(dev)
        toff, tqoff = torch.from_numpy(off).to(dev), torch.from_numpy(qoff).to(dev)
        for k in ks:
            idx = torch.zeros((q.shape[0], k), dtype=torch.int32, device=dev)
            d2 = torch.zeros((q.shape[0], k), dtype=torch.float32, device=dev)
            ref.knnquery_cuda(q.shape[0], k, txyz, tq, toff, tqoff, idx, d2)
            torch.cuda.synchronize()
            res[f"knn/{name}/{k}/idx"] = idx.cpu().numpy()
            res[f"knn/{name}/{k}/d2"] = d2.cpu().numpy()
    for name, (builder, stride) in cases.FPS_CASES.items():
        xyz, off = builder()
        noff = cases.fps_new_offset(off, stride)
        lens = np.diff(np.concatenate([[0], off]))
        txyz = torch.from_numpy(xyz).to(dev)
        idx = torch.zeros(int(noff[-1]), dtype=torch.int32, device=dev)
        tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=dev)
        ref.furthestsampling_cuda(len(off), int(lens.max()), txyz, torch.from_numpy(off).to(dev),
                                  torch.from_numpy(noff).to(dev), tmp, idx)
        torch.cuda.synchronize()
        res[f"fps/{name}/idx"] = idx.cpu().numpy()
        res[f"fps/{name}/tmp"] = tmp.cpu().numpy()
    # K3-K6 forward/backward on random idx
    n, k, c, wc, inp, inp2, pos, w, idx, go_nkc, go_nc, wk = cases.ops_inputs()
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
    o = z(n, k, c); ref.grouping_forward_cuda(n, k, c, t(inp), t(idx), o); res["ops/grouping_fwd"] = o.cpu().numpy()
    o = z(n, c); ref.grouping_backward_cuda(n, k, c, t(go_nkc), t(idx), o); res["ops/grouping_bwd"] = o.cpu().numpy()
    o = z(n, k, c); ref.subtraction_forward_cuda(n, k, c, t(inp), t(inp2), t(idx), o); res["ops/subtraction_fwd"] = o.cpu().numpy()
    g1, g2 = z(n, c), z(n, c); ref.subtraction_backward_cuda(n, k, c, t(idx), t(go_nkc), g1, g2)
    res["ops/subtraction_bwd1"], res["ops/subtraction_bwd2"] = g1.cpu().numpy(), g2.cpu().numpy()
    o = z(n, c); ref.aggregation_forward_cuda(n, k, c, wc, t(inp), t(pos), t(w), t(idx), o); res["ops/aggregation_fwd"] = o.cpu().numpy()
    gi, gp, gw = z(n, c), z(n, k, c), z(n, k, wc)
    ref.aggregation_backward_cuda(n, k, c, wc, t(inp), t(pos), t(w), t(idx), t(go_nc), gi, gp, gw)
    res["ops/aggregation_bwd_i"], res["ops/aggregation_bwd_p"], res["ops/aggregation_bwd_w"] = gi.cpu().numpy(), gp.cpu().numpy(), gw.cpu().numpy()
    idx3 = np.ascontiguousarray(idx[:, :3])
    o = z(n, c); ref.interpolation_forward_cuda(n, c, 3, t(inp), t(idx3), t(wk), o); res["ops/interpolation_fwd"] = o.cpu().numpy()
    o = z(n, c); ref.interpolation_backward_cuda(n, c, 3, t(go_nc), t(idx3), t(wk), o); res["ops/interpolation_bwd"] = o.cpu().numpy()
    torch.cuda.synchronize()
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **res)
    print("wrote", out, len(res), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pointops_ref_gpu.npz"))
