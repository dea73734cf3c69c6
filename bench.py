#!/usr/bin/env python
"""bench.py — contract bench for the hot path of LiyaoTang/contrastBoundary on B200.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Metric (BASELINE.json): points/sec, Point-Transformer + CBL forward + backward (+ SGD step, as the
reference's train iteration, pytorch/tool/train.py:315-326) on synthetic S3DIS-shape scenes,
per-GPU batch 4 x 40960 points (configs[1]; configs[3] under torchrun = data parallel, NCCL grad
all-reduce only).  One JSON line on rank 0.  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES_PER_GPU = 4
POINTS_PER_SCENE = 40960
METRIC = "points/sec fwd+bwd S3DIS-shape scenes (PT+CBL)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "stock-gpu"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3], help="BASELINE.json configs[] (1-based): 2 = PT+CBL (the metric's "
                    "configuration, default), 3 = ConvNet(AdaptiveWeight)+CBL on the device-built radius pyramid")
    ap.add_argument("--sweep", action="store_true", help="config 5: the KNN+gather grid N x K x C -> profiles/knn_gather_sweep.json")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=360.0)
    ap.add_argument("--no-pipeline", action="store_true", help="compute each batch's geometry inside its own step (no look-ahead)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python (stream mode) instead of replaying CUDA graphs")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# --------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's network restated op-by-op (oracle/ref_model.py) on
# the box's host cores with the C restatement of its kernels (oracle/) — bounded sample.
# --------------------------------------------------------------------------------------------------
WORKLOAD = (f"Point-Transformer+CBL fwd+bwd+SGD, {SCENES_PER_GPU} x {POINTS_PER_SCENE}-pt synthetic S3DIS-shape scenes per GPU, "
            f"K=16 (stage 0: 8), C=32->512, CBL nsample [36,24,24,24,24]")


def cpu_reference_run(steps, warmup, budget_s):
    """the reference's CPU path: its own model code (oracle/reference_runner.py) on torch-CPU + the C/OpenMP restatement of
    its pointops kernels, all host threads.  Runs the FULL workload (4 x 40960 points per step, the batches bench.py's own arm
    times) whenever (steps + warmup) of them fit `budget_s`; otherwise a bounded sample (fewer scenes, then fewer points)."""
    import torch
    from contrastboundary_b200 import synthetic
    import oracle
    from oracle import reference_runner
    cores = os.cpu_count() or 1
    torch.manual_seed(0)
    model, crit, what = reference_runner.build("cpu")
    opt = torch.optim.SGD(model.parameters(), lr=0.5, momentum=0.9, weight_decay=1e-4)
    model.train()

    def one(n_scenes, n_pts, seed):
        b = synthetic.make_batch(n_scenes, n_pts, seed)
        inputs = {k: torch.from_numpy(b[k]) for k in ("points", "features", "offset")}
        target = torch.from_numpy(b["point_labels"])
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out, up = model(inputs)
        loss = crit(out, target, up)
        loss.sum().backward()
        opt.step()
        return time.perf_counter() - t0

    # "all the host threads it can use": pick the thread count that maximises throughput on a probe
    # (torch's CPU kernels on these small (n,k,c) tensors get SLOWER when oversubscribed across sockets)
    best_t, t8, t4 = cores, None, None
    for t in sorted({min(cores, x) for x in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(t)
        oracle.set_num_threads(t)
        d4 = one(1, 4096, 0)
        d8 = one(1, 8192, 1)
        if t8 is None or d8 < t8:
            best_t, t8, t4 = t, d8, d4
    cores = best_t
    torch.set_num_threads(cores)
    oracle.set_num_threads(cores)
    # cost model t(n) = a n + b n^2 per scene (dense part linear, brute-force searches quadratic), fitted on the two probes
    bq = max((t8 - 2.0 * t4) / (8192.0 ** 2 - 2.0 * 4096.0 ** 2), 0.0)
    al = max((t4 - bq * 4096.0 ** 2) / 4096.0, 1e-9)
    est = lambda sc, n: sc * (al * n + bq * n * n)
    total = max(steps + warmup, 1)
    n_scenes, n_pts = SCENES_PER_GPU, POINTS_PER_SCENE
    while est(n_scenes, n_pts) * total > budget_s and n_scenes > 1:
        n_scenes //= 2
    while est(n_scenes, n_pts) * total > budget_s and n_pts > 4096:
        n_pts -= 4096
    full = (n_scenes, n_pts) == (SCENES_PER_GPU, POINTS_PER_SCENE)
    for w in range(warmup):
        one(n_scenes, n_pts, 5000 + w)
    ts = [one(n_scenes, n_pts, 5000 + warmup + s) for s in range(steps)]
    dt = float(np.sum(ts))
    sample = (("the full workload: " if full else f"bounded sample (full config is {SCENES_PER_GPU} x {POINTS_PER_SCENE}): ")
              + f"{n_scenes} scene(s) x {n_pts} pts per step; fwd+bwd+SGD of {what}, torch CPU ({cores} threads) + C/OpenMP "
                f"restatement of the reference's pointops kernels")
    return {"value": steps * n_scenes * n_pts / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores, "n_pts": n_pts,
            "n_scenes": n_scenes, "full": full, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, args.cpu_budget_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch_scenes": r["n_scenes"], "points_per_scene": r["n_pts"],
                   "full_workload": r["full"], "parallelism": "host cores", "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# gpu_baseline = B1 of BASELINE.md §3, the north star's >= 10x denominator: the reference's own model code on the
# reference's own pointops CUDA kernels (compiled unmodified for sm_100a, oracle/_ref/pointops_cuda.so), same B200, same
# batches, same step.  Runs in a child process (`--impl stock-gpu`) AFTER this arm's timed regions.
# --------------------------------------------------------------------------------------------------
def run_stock_gpu(args):
    import torch
    from contrastboundary_b200 import synthetic
    from oracle import reference_runner
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model, crit, what = reference_runner.build("stock-gpu")
    model, crit = model.to(dev), crit.to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.5, momentum=0.9, weight_decay=1e-4)
    model.train()
    batches = []
    for i in range(3):
        b = synthetic.make_batch(SCENES_PER_GPU, POINTS_PER_SCENE, 5000 + i)
        batches.append(({k: torch.from_numpy(b[k]).to(dev) for k in ("points", "features", "offset")},
                        torch.from_numpy(b["point_labels"]).to(dev)))

    def step(i):
        inputs, target = batches[i % len(batches)]
        opt.zero_grad(set_to_none=True)
        out, up = model(inputs)
        loss = crit(out, target, up)
        loss.sum().backward()
        opt.step()
        return loss

    for w in range(max(args.warmup, 1)):
        step(w)
    torch.cuda.synchronize()
    clocks = Clocks(dev.index)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in range(args.steps):
        loss = step(s_)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / args.steps
    loaded = sorted({l.split()[-1] for l in open("/proc/self/maps") if "pointops_cuda" in l or "libcbops" in l})
    print("STOCK " + json.dumps({
        "what": f"B1: {what} + the reference's stock pointops CUDA kernels (oracle/_ref/pointops_cuda.so, sm_100a), torch CUDA fp32",
        "value": SCENES_PER_GPU * POINTS_PER_SCENE / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "steps": args.steps,
        "warmup": max(args.warmup, 1), "workload": WORKLOAD, "clocks": clk, "loss": [round(float(x), 5) for x in loss],
        "native_loaded": [os.path.relpath(x, ROOT) for x in loaded]}))


def gpu_baseline_subprocess(steps=5, warmup=2, timeout=600):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "stock-gpu", "--steps", str(steps), "--warmup", str(warmup)],
                           capture_output=True, text=True, timeout=timeout)
        for l in r.stdout.splitlines():
            if l.startswith("STOCK "):
                return json.loads(l[6:])
        return {"unavailable": (r.stderr or r.stdout)[-300:]}
    except Exception as e:                                   # the baseline must never take the bench down
        return {"unavailable": repr(e)[:300]}


# --------------------------------------------------------------------------------------------------
# roofline of the north-star kernel (fused KNN + gather), measured live
# --------------------------------------------------------------------------------------------------
def roofline_knn_gather(torch, dev):
    """the north-star kernel: fused KNN + neighbour-feature gather, N=40960, K=16, C=256, single scene.
    `achieved` times the search+gather kernel on a prebuilt support grid (cb_knn_gather_grid: k_knn_gather +
    its tie-replay / re-gather followers); the whole operator including the grid build is reported next to it."""
    from contrastboundary_b200 import fused, synthetic
    n, k, c = 40960, 16, 256
    xyz = torch.from_numpy(synthetic.make_scene(n, 4242)[0]).to(dev)
    off = torch.tensor([n], dtype=torch.int32, device=dev)
    feat = torch.randn(n, c, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    grid = fused.grid_build(xyz, off, k)
    out = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
    st = torch.cuda.current_stream()

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()                      # L2 flush between timed iterations: write 256 MiB (> 126 MB L2) ...
            flush_r.sum()                      # ... then read 256 MiB, so that the L2 holds CLEAN foreign lines: the timed
                                               # kernel is not charged for writing back the flush's dirty lines
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            fn()
            b.record(st)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return float(np.mean(ts))

    t_kernel = timed(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out))
    t_call = timed(lambda: fused.knn_gather(k, xyz, xyz, feat, off, off))
    alg = 12 * n + 4 * n * c + 8 * n * k + 4 * n * k * c
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "knn_gather_traffic.json"))).get("dram_bytes_per_call")
    except Exception:
        pass
    ach = alg / t_kernel / 1e9
    return {"bound": "hbm", "kernel": "k_knn_gather (fused grid-KNN + TMA neighbour-row gather), N=40960 K=16 C=256, grid prebuilt",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "alg_bytes": alg, "us_per_launch": t_kernel * 1e6,
            "whole_operator": {"what": "cb_knn_gather incl. the grid build (one cooperative kernel) and the tie-replay / re-gather followers", "us": t_call * 1e6,
                               "achieved": alg / t_call / 1e9, "frac": alg / t_call / 1e9 / peak},
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)" if peaks else "fallback 6650 (B200_PROFILING.md)",
            "l2": "between timed iterations a 256 MiB buffer is zeroed and another 256 MiB buffer is read (L2 left full of clean foreign lines)"}


def roofline_linear(torch, dev):
    """a kernel that IS in the timed step: the fused q/k/v projection of the level-1 PointTransformer blocks
    (40960 x 64 -> 192) and the level-0 one (163840 x 32 -> 96) through cb_linear_forward = k_umma_linear2
    (tcgen05.mma kind::tf32, TMEM accumulators).  Algorithmic bytes 4 n (ci + co)."""
    from contrastboundary_b200 import _lib as L
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    out = []
    for n, ci, co in ((40960, 64, 192), (163840, 32, 96)):
        x, w, b = torch.randn(n, ci, device=dev), torch.randn(co, ci, device=dev), torch.randn(co, device=dev)
        y = torch.empty(n, co, device=dev)
        for _ in range(3):
            L.call("cb_linear_forward", n, ci, co, x, w, b, y, L.stream())
        ts = []
        for _ in range(10):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            L.call("cb_linear_forward", n, ci, co, x, w, b, y, L.stream())
            e.record(st)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e-3)
        t = float(np.mean(ts))
        alg = 4 * n * (ci + co)
        out.append({"kernel": f"k_umma_linear2 (tcgen05 kind::tf32 3xTF32, TMEM), n={n} ci={ci} co={co}", "bound": "hbm", "alg_bytes": alg,
                    "us_per_launch": t * 1e6, "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak})
    return out


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1
    if ddp:
        dist.init_process_group("nccl", device_id=dev)
    from contrastboundary_b200 import _lib, engine, model, synthetic
    _lib.lib()   # fail loudly if libcbops.so is missing

    cfg = model.CBLConfig()
    npool = 3
    host = [engine.host_batch_from_numpy(synthetic.make_batch(SCENES_PER_GPU, POINTS_PER_SCENE, 5000 + 97 * rank + i))
            for i in range(npool)]
    dev_batches = [engine.to_device(h, dev) for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values() if isinstance(v, torch.Tensor))
    loss_host = torch.empty(6, dtype=torch.float32).pin_memory()

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(steps):
            fn(s)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) * 1e-3], device=dev, dtype=torch.float64)
        if ddp:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pipeline = not args.no_pipeline
    nwarm = max(args.warmup, 3)
    mode, graph_error, ts = "stream", None, None
    if not args.no_graph:
        # whole-step CUDA graphs (engine.GraphTrainStep): 2 eager steps, capture, then replays
        ts = engine.GraphTrainStep(cfg, dev, ddp=ddp, seed=0)
        for w in range(nwarm + 2):
            ts.step(dev_batches[w % npool])
        ok = torch.tensor([0 if ts.graph_error else 1], device=dev)
        if ddp:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            mode = "graph"
        else:
            graph_error = ts.graph_error or "capture failed on another rank"
            print("bench.py: CUDA-graph capture failed, falling back to stream mode: %s" % graph_error, file=sys.stderr)
            ts = None
            torch.cuda.empty_cache()

    if mode == "graph":
        def step_plain(s):
            ts.step(dev_batches[s % npool])

        def step_resident(s):
            # geometry graph of batch s+1 replays on the side stream while the network graph of batch s replays
            ts.step(dev_batches[s % npool], next_batch=dev_batches[(s + 1) % npool] if pipeline else None)

        def step_e2e(s):
            # pinned HOST batches: each batch is copied host->device exactly once (into the graph's static slot)
            loss = ts.step(host[s % npool], next_batch=host[(s + 1) % npool] if pipeline else None)
            loss_host.copy_(loss, non_blocking=True)             # D2H of the step's result
            torch.cuda.current_stream().synchronize()

        t_plain = timed(step_plain, args.steps)                  # reference point: no look-ahead
        for w in range(npool):
            step_resident(w)                                     # primes the pipeline: slot holds batch 0's geometry
        clocks = Clocks(local)
        if rank == 0:
            clocks.start()
        t_val = timed(step_resident, args.steps)
        clk = clocks.stop() if rank == 0 else None
        launches = int(ts.launches_per_step or 0)
        for w in range(npool):
            step_e2e(w)
        t_e2e = timed(step_e2e, args.steps)
    else:
        ts = engine.TrainStep(cfg, dev, ddp=ddp, seed=0)

        def step_resident(s):
            # geometry of batch s+1 (FPS + all neighbour searches) runs on a side stream during step s
            ts.step(dev_batches[s % npool], next_batch=dev_batches[(s + 1) % npool] if pipeline else None)

        def step_plain(s):
            ts.step(dev_batches[s % npool])

        e2e_next = {}

        def step_e2e(s):
            cur = e2e_next.pop("b", None)
            if cur is None:
                cur = engine.to_device(host[s % npool], dev)
            nxt = None
            if pipeline:
                nxt = engine.to_device(host[(s + 1) % npool], dev)   # H2D of ONE batch per step (the next one), from pinned memory
                e2e_next["b"] = nxt
            loss = ts.step(cur, next_batch=nxt)
            loss_host.copy_(loss, non_blocking=True)             # D2H of the step's result
            torch.cuda.current_stream().synchronize()

        for w in range(nwarm):
            step_plain(w)
        t_plain = timed(step_plain, args.steps)                  # reference point: no look-ahead
        if pipeline:
            ts.prefetch_geometry(dev_batches[0])                 # prologue: batch 0's geometry (outside the timed region)
            step_resident(0); step_resident(1); step_resident(2)
            ts.prefetch_geometry(dev_batches[0]) if id(dev_batches[0]) not in ts._geo else None
        lc0 = _lib.launch_count()
        clocks = Clocks(local)
        if rank == 0:
            clocks.start()
        t_val = timed(step_resident, args.steps)
        launches = (_lib.launch_count() - lc0) // max(args.steps, 1)
        clk = clocks.stop() if rank == 0 else None
        ts._geo.clear()
        for w in range(2):
            step_e2e(w)
        e2e_next.clear(); ts._geo.clear()
        if pipeline:
            first = engine.to_device(host[0], dev)
            e2e_next["b"] = first
            ts.prefetch_geometry(first)
        t_e2e = timed(step_e2e, args.steps)
    pts_per_step = world * SCENES_PER_GPU * POINTS_PER_SCENE
    line = {
        "metric": METRIC, "value": pts_per_step * args.steps / t_val, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_val / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch_scenes": world * SCENES_PER_GPU, "parallelism": f"dp{world}",
                   "fused": bool(cfg.fused),
                   "launch_mode": ("CUDA graphs: geometry graph + network(fwd+loss+bwd) graph per slot, 2 slots; optimizer"
                                   + (" and ONE NCCL all-reduce of the packed gradient" if ddp else "") + " outside the graphs")
                   if mode == "graph" else "stream mode (every kernel issued from Python)" + (", graph capture failed: " + graph_error if graph_error else ""),
                   "geometry_pipeline": ("look-ahead 1: FPS + neighbour searches of batch t+1 run on a side stream during step t "
                                         "(every batch's geometry is computed exactly once, inside the timed region)") if pipeline else "off",
                   "ms_per_step_without_lookahead": 1e3 * t_plain / args.steps,
                   "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": pts_per_step * args.steps / t_e2e, "unit": "points/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 24, "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    if rank == 0 and world == 1 and not args.no_roofline:
        line["roofline"] = roofline_knn_gather(torch, dev)
        line["roofline_step_kernels"] = roofline_linear(torch, dev)
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        # B1: the reference's stock pointops CUDA build on this same GPU, after the timed regions, in a child process
        del ts
        torch.cuda.empty_cache()
        line["gpu_baseline"] = gpu_baseline_subprocess()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 0, 45.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line))
    if ddp:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# config 3 (ConvNet + CBL, bench_convnet.py): its CPU baseline leg — the oracle side stays out of the package
# --------------------------------------------------------------------------------------------------
def convnet_cpu_pyramid_baseline(pts, lens):
    """the reference's own C++ operators (oracle/_ref/libref_cpu.so = its sources compiled unmodified) building the same
    5-level pyramid on one host thread, as the TF op runs them"""
    import oracle
    from contrastboundary_b200.tf_pyramid import PyramidConfig
    cfg = PyramidConfig()
    if oracle.have_ref_cpu():
        nb, sub, kind = oracle.ref_batch_neighbors, oracle.ref_batch_grid_subsampling, "reference"
    else:
        nb, sub, kind = oracle.batch_neighbors, oracle.batch_grid_subsampling, "port"
    t0 = time.perf_counter()
    dl, r = cfg.first_subsampling_dl, cfg.first_subsampling_dl * cfg.density_parameter / 2
    p, l = pts, lens
    for lvl in range(cfg.num_layers - 1):
        nb(p, p, l, l, r)
        pp, pl = sub(p, l, 2 * dl)
        nb(pp, p, pl, l, r)
        nb(p, pp, l, pl, 2 * r)
        p, l, dl, r = pp, pl, dl * 2, r * 2
    nb(p, p, l, l, r)
    dt = time.perf_counter() - t0
    return {"value": len(pts) / dt, "unit": "points/s", "cores": 1, "kind": kind,
            "sample": f"INPUT PYRAMID ONLY (grid subsampling + radius neighbours, 5 levels) of one batch of {len(lens)} x "
                      f"{int(lens[0])} points with the reference's own C++ operators, single thread as the TF op runs them "
                      f"({dt * 1e3:.0f} ms); the TF network itself cannot run here (no TensorFlow)"}



# --------------------------------------------------------------------------------------------------
# config 5: KNN + gather microbench grid -> profiles/knn_gather_sweep.json
# --------------------------------------------------------------------------------------------------
def run_sweep(args):
    import torch
    from contrastboundary_b200 import _lib, fused, synthetic
    _lib.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def timed(fn, iters):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            flush_r.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            fn()
            b.record(st)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return float(np.mean(ts))

    clocks = Clocks(0)
    clocks.start()
    points = []
    for e in range(14, 21):
        n = 1 << e
        xyz = torch.from_numpy(synthetic.make_scene(n, 4242 + e)[0]).to(dev)
        off = torch.tensor([n], dtype=torch.int32, device=dev)
        for c in (64, 256):
            feat = torch.randn(n, c, device=dev)
            for k in (16, 32, 64):
                alg = 12 * n + 4 * n * c + 8 * n * k + 4 * n * k * c
                rec = {"N": n, "K": k, "C": c, "alg_bytes": alg}
                try:
                    grid = fused.grid_build(xyz, off, k)
                    out = fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off)
                    rec["replayed_queries"] = int(grid[:32].view(torch.int32)[1])     # CbGridHeader.flagged_count: exact-tie replays
                    iters = 10 if alg < (4 << 30) else 3
                    tk = timed(lambda: fused.knn_gather_grid(grid, k, xyz, xyz, feat, off, off, out), iters)
                    del out
                    tc = timed(lambda: fused.knn_gather(k, xyz, xyz, feat, off, off), iters)
                    rec.update({"kernel_us": tk * 1e6, "kernel_gbs": alg / tk / 1e9, "kernel_frac": alg / tk / 1e9 / peak,
                                "operator_us": tc * 1e6, "operator_gbs": alg / tc / 1e9, "operator_frac": alg / tc / 1e9 / peak})
                    del grid
                except Exception as ex:                      # e.g. the 68.7 GB output of N=2^20, K=64, C=256 on a busy GPU
                    rec["skipped"] = repr(ex)[:200]
                    torch.cuda.empty_cache()
                points.append(rec)
                print(json.dumps(rec), file=sys.stderr, flush=True)
            del feat
            torch.cuda.empty_cache()
    clk = clocks.stop()
    ok = [p for p in points if "kernel_frac" in p]
    res = {"what": "config 5: fused KNN + neighbour-feature gather (cb_knn_gather), single synthetic S3DIS-shape scene of N points, "
                   "self query; kernel = cb_knn_gather_grid on a prebuilt grid, operator = cb_knn_gather incl. the grid build",
           "alg_bytes": "12N + 4NC + 8NK + 4NKC (SURVEY 8d)", "peak_gbs": peak, "peak_source": peak_src,
           "l2": "between timed iterations a 256 MiB buffer is zeroed and another is read", "clocks": clk, "points": points,
           "summary": {"n_points": len(points), "n_measured": len(ok),
                       "operator_frac_min": min(p["operator_frac"] for p in ok) if ok else None,
                       "operator_frac_median": float(np.median([p["operator_frac"] for p in ok])) if ok else None,
                       "operator_frac_max": max(p["operator_frac"] for p in ok) if ok else None,
                       "points_at_or_above_0.70_operator": sum(p["operator_frac"] >= 0.70 for p in ok),
                       "points_at_or_above_0.70_kernel": sum(p["kernel_frac"] >= 0.70 for p in ok)}}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "profiles", "knn_gather_sweep.json"), "w"), indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "knn_gather_sweep.json"), "w"), indent=1)
    print(json.dumps({"sweep": "profiles/knn_gather_sweep.json", **res["summary"]}))


if __name__ == "__main__":
    a = parse()
    if a.sweep:
        run_sweep(a)
    elif a.impl == "reference":
        run_reference(a)
    elif a.impl == "stock-gpu":
        run_stock_gpu(a)
    elif a.config == 3:
        from contrastboundary_b200 import bench_convnet
        bench_convnet.run(a, Clocks, ROOT, convnet_cpu_pyramid_baseline)
    else:
        run_ours(a)
