"""bench.py --config 3: BASELINE configs[2] — ConvNet (AdaptiveWeight ResNet) + CBL forward + backward (+ global-norm clip +
Momentum SGD, the reference's train iteration, tensorflow/utils/trainer.py) on 4 synthetic input spheres of 15000 points
(in_radius 2.0 on a 0.04 m grid), INCLUDING the device-side build of the 5-level radius pyramid the reference builds on
tf.data CPU workers (tensorflow/datasets/base.py:767-842).  Same JSON contract as the default bench; `parity` says what the
checker of this configuration is (TensorFlow is not installable here)."""
import json
import os
import sys
import time

import numpy as np

SPHERES_PER_GPU = 4
POINTS_PER_SPHERE = 15000
METRIC = "points/sec fwd+bwd S3DIS-shape spheres (ConvNet+CBL)"
WORKLOAD = (f"ConvNet (AdaptiveWeight ResNet, first_features_dim 72, 5 levels, limits [26,31,38,41,39]) + multi-scale head + CBL "
            f"fwd+bwd+clip+Momentum, {SPHERES_PER_GPU} x {POINTS_PER_SPHERE}-pt synthetic input spheres per GPU, 5-level radius pyramid "
            f"built on the device inside the step")


def _host_batch(torch, seed):
    from . import synthetic
    scenes = [synthetic.make_scene(POINTS_PER_SPHERE, seed * 1000 + i) for i in range(SPHERES_PER_GPU)]
    b = {"points": torch.from_numpy(np.concatenate([s[0] for s in scenes])).pin_memory(),
         "colors": torch.from_numpy(np.concatenate([s[1] for s in scenes])).pin_memory(),
         "point_labels": torch.from_numpy(np.concatenate([s[2] for s in scenes])).pin_memory(),
         "lens": torch.tensor([POINTS_PER_SPHERE] * SPHERES_PER_GPU, dtype=torch.int32).pin_memory()}
    return b


def _roofline_adaptive_weight(torch, ts, inputs, root):
    """the aggregation kernel at the widest level-0 call (n0 points, K = limits[0], c = first_features_dim): SURVEY 8(d)
    fused-aggregation bytes 12n + 4nc(in) + 4nc(out) + 4nK"""
    from .tf_model import adaptive_weight
    pts, nb = inputs["points"][0], inputs["neighbors"][0]
    n, k = nb.shape
    c = ts.cfg.first_features_dim
    la = ts.model.resnet_backbone.res1_simple_block
    feat = torch.randn(n, c, device=pts.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=pts.device)
    st = torch.cuda.current_stream()
    r = ts.cfg.first_subsampling_dl * ts.cfg.density_parameter

    def run():
        with torch.no_grad():
            return adaptive_weight(pts, pts, nb, feat, la.fc_1.weight, la.fc_1.bias, r)
    for _ in range(3):
        run()
    ts_ = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        run()
        b.record(st)
        torch.cuda.synchronize()
        ts_.append(a.elapsed_time(b) * 1e-3)
    t = float(np.mean(ts_))
    alg = 12 * n + 4 * n * c + 4 * n * c + 4 * n * k
    try:
        peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
    except Exception:
        peak, src = 6650.0, "fallback 6650 (B200_PROFILING.md)"
    return {"bound": "hbm", "kernel": f"k_aw<0> + k_aw_maxidx (cb_adaptive_weight_forward), n={n} K={k} c={c}", "achieved": alg / t / 1e9,
            "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak, "traffic": None, "alg_bytes": alg, "us_per_launch": t * 1e6,
            "peak_source": src, "l2": "a 256 MiB buffer is zeroed between timed iterations",
            "note": "the gather re-reads every feature row ~K times out of L2 (table = %.1f MB): latency/L2 bound, not HBM bound" % (4e-6 * n * c)}


def run(args, Clocks, root, cpu_baseline_fn=None):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        raise SystemExit("bench.py --config 3 is a single-GPU bench line (configs[2]); the scaling bench is the default configuration")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from . import _lib, convnet
    _lib.lib()
    ts = convnet.ConvNetTrainStep(convnet.ConvNetConfig(), dev, seed=0)
    npool = 3
    host = [_host_batch(torch, 7000 + i) for i in range(npool)]
    devb = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    loss_host = torch.empty(6, dtype=torch.float32).pin_memory()

    def timed(fn, steps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(steps):
            fn(s)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3

    pipeline = not getattr(args, "no_pipeline", False)

    def step_plain(s):
        ts.step(devb[s % npool])

    def step_resident(s):
        # the pyramid of batch s+1 is built by a worker thread on a side stream while the network of batch s is issued
        ts.step(devb[s % npool], next_batch=devb[(s + 1) % npool] if pipeline else None)

    e2e_next = {}

    def step_e2e(s):
        cur = e2e_next.pop(s, None)
        if cur is None:
            cur = {k: v.to(dev, non_blocking=True) for k, v in host[s % npool].items()}
        nxt = None
        if pipeline:
            nxt = e2e_next[s + 1] = {k: v.to(dev, non_blocking=True) for k, v in host[(s + 1) % npool].items()}   # H2D of batch s+1
        loss = ts.step(cur, next_batch=nxt)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    nwarm = max(args.warmup, 3)
    for w in range(nwarm):
        step_plain(w)
    t_plain = timed(step_plain, args.steps)                      # reference point: no look-ahead
    for w in range(npool):
        step_resident(w)                                         # primes the pipeline
    lc0 = _lib.launch_count()
    clocks = Clocks(local)
    clocks.start()
    t_val = timed(step_resident, args.steps)
    clk = clocks.stop()
    launches = (_lib.launch_count() - lc0) // max(args.steps, 1)
    e2e_next.clear()
    for w in range(2):
        step_e2e(w)
    e2e_off = 2

    def step_e2e_timed(s):
        step_e2e(s + e2e_off)
    t_e2e = timed(step_e2e_timed, args.steps)
    e2e_next.clear()
    ts.drain_prefetch()
    # breakdown: pyramid alone, network alone (on a prebuilt pyramid)
    t_pyr = timed(lambda s: ts.build_inputs(devb[s % npool]), args.steps)
    inputs = ts.build_inputs(devb[0])
    t_net = timed(lambda s: ts.step(devb[0], inputs), args.steps)
    pts = SPHERES_PER_GPU * POINTS_PER_SPHERE
    line = {
        "metric": METRIC, "value": pts * args.steps / t_val, "unit": "points/s", "n_gpus": 1, "steps": args.steps, "warmup": nwarm,
        "ms_per_step": 1e3 * t_val / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch_spheres": SPHERES_PER_GPU, "parallelism": "dp1",
                   "launch_mode": "stream mode (every kernel issued from Python; pyramid sizes are data dependent)",
                   "pyramid_pipeline": ("look-ahead 1: the 5-level pyramid of batch t+1 is built by a worker thread on a side stream during "
                                        "step t, as the reference's tf.data workers do (every batch's pyramid is built exactly once, "
                                        "inside the timed region)") if pipeline else "none",
                   "ms_per_step_without_lookahead": 1e3 * t_plain / args.steps,
                   "ms_pyramid": 1e3 * t_pyr / args.steps, "ms_network_fwd_bwd_update": 1e3 * t_net / args.steps,
                   "level_points": [int(p.shape[0]) for p in inputs["points"]],
                   "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
        "parity": "forward pinned under a stand-in runtime: TensorFlow is not installable here, so the reference's own TF source - its "
                  "operators, and its whole SceneSegModel builder with its own config object + adapt.yaml - is executed on a NumPy "
                  "stand-in of the TF-1 API (tests/golden/make_golden_tf_ops.py). Directly against those vectors: the CUDA AdaptiveWeight "
                  "kernel, label votes and soft-NN loss; transitively (CUDA network = float64 restatement on the GPU, restatement = "
                  "executed reference on the CPU): logits, every loss entry, the L2 term. Gradients: float64 finite differences of "
                  "that forward and two independent restatements (oracle/tf_model.py, oracle/tf_convnet_np.py)",
        "e2e": {"value": pts * args.steps / t_e2e, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 24,
                "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clk,
    }
    if not args.no_roofline:
        line["roofline"] = _roofline_adaptive_weight(torch, ts, inputs, root)
    if not args.no_cpu_baseline and cpu_baseline_fn is not None:
        line["cpu_baseline"] = cpu_baseline_fn(host[0]["points"].numpy(), host[0]["lens"].numpy())
    if rank == 0:
        print(json.dumps(line))
