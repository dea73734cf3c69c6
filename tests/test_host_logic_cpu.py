"""CPU tests of the host-side logic around the CUDA path (no compute call into libcbops): level sizes of the
TransitionDown chain, the stacked batch-index matrices of the TF pyramid, the CPU fall-back of the BatchNorm wrappers
(state_dict compatibility with nn.BatchNorm1d) and the IoU histograms of the boundary evaluation."""
import numpy as np
import torch
import torch.nn as nn

from contrastboundary_b200 import boundary_eval, linear_ops, model, tf_pyramid, trainer


def test_level_offsets_follow_the_reference_rule():
    """per-scene floor(n_b / stride), accumulated (pytorch/model/blocks.py:64-67): remainders are lost at every level"""
    cfg = model.CBLConfig()
    ohs = model.level_offsets_host([40960, 40960 + 30001, 40960 + 30001 + 7], cfg)
    assert ohs[0] == [40960, 70961, 70968]
    lens = [40960, 30001, 7]
    for l in range(1, 5):
        lens = [x // 4 for x in lens]
        assert ohs[l] == list(np.cumsum(lens)), l
    assert ohs[4][-1] - ohs[4][-2] == 0            # the 7-point scene has vanished by level 2


def test_stack_batch_inds_matches_reference_semantics():
    """tf_stack_batch_inds_while (tensorflow/datasets/base.py:694-737): rows padded with n = sum(lens); one extra shadow
    column only when no row is padded"""
    a = tf_pyramid.stack_batch_inds(torch.tensor([3, 2, 5], dtype=torch.int32)).numpy()
    assert np.array_equal(a, np.array([[0, 1, 2, 10, 10], [3, 4, 10, 10, 10], [5, 6, 7, 8, 9]], np.int32))
    b = tf_pyramid.stack_batch_inds(torch.tensor([2, 2], dtype=torch.int32)).numpy()
    assert np.array_equal(b, np.array([[0, 1, 4], [2, 3, 4]], np.int32))
    c = tf_pyramid.stack_batch_inds(torch.tensor([2, 2], dtype=torch.int32), tight=True).numpy()
    assert np.array_equal(c, np.array([[0, 1], [2, 3]], np.int32))


def test_pyramid_host_logic_matches_executed_reference_source():
    """The reference's input-pyramid builder (tensorflow/datasets/base.py:694-737,767-842), EXECUTED on the NumPy TF stand-in with its
    two custom ops served by the CPU oracle of the reference's C++ (tests/golden/make_golden_tf_ops.py): the golden script asserts
    that its five levels ARE the arrays stored as net/* (the oracle composition that tf_pyramid.segmentation_inputs_radius is
    tested against on the GPU); here the product's host-side pieces reproduce its batch index matrices and batch weights."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_ops_ref.npz"))
    lens0, lens4 = g["cbl/batches_len/0"], g["cbl/batches_len/4"]
    a = tf_pyramid.stack_batch_inds(torch.from_numpy(lens0.astype(np.int32))).numpy()
    b = tf_pyramid.stack_batch_inds(torch.from_numpy(lens4.astype(np.int32))).numpy()
    assert np.array_equal(a, g["pyr/in_batches"]) and np.array_equal(b, g["pyr/out_batches"])
    inds = np.repeat(np.arange(len(lens0)), lens0)
    w = (lens0.min().astype(np.float32) / lens0.astype(np.float32))[inds]                   # tf_pyramid.py:52-53 (base.py:776-779)
    assert np.allclose(w, g["pyr/batch_weights"], rtol=1e-6, atol=0)
    assert sum(len(g[f"net/points/{l}"]) for l in range(5)) == 2183 and g["net/upsamples/0"].shape == (0, 1) and g["net/pools/4"].shape == (0, 1)


def test_batchnorm_wrapper_is_a_drop_in_on_cpu():
    """linear_ops.BatchNorm1d / bn_act: same parameters, buffers, state_dict and results as nn.BatchNorm1d (+ add + relu)"""
    torch.manual_seed(0)
    ref, ours = nn.BatchNorm1d(8), linear_ops.BatchNorm1d(8)
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    x, r = torch.randn(50, 8), torch.randn(50, 8)
    for training in (True, False):
        ref.train(training); ours.train(training)
        want = torch.relu(ref(x) + r)
        got = linear_ops.bn_act(ours, x, residual=r, relu=True)
        assert torch.allclose(got, want, atol=1e-6)
    linear_ops.flush_bn_counters()
    for k, v in ref.state_dict().items():
        assert torch.allclose(ours.state_dict()[k].float(), v.float(), atol=1e-6), k


def test_network_state_dict_names_match_the_oracle_network():
    """the product network and the op-by-op restatement of the reference network share every parameter / buffer name
    (the restatement's names are pinned to the REAL reference by tests/golden/model_ref.npz)"""
    from oracle import cpu_pointops, ref_model
    ours = model.PointTransformerSeg(model.CBLConfig())
    ref = ref_model.RefSeg(cpu_pointops)
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b


def test_intersection_and_union_histograms():
    rng = np.random.default_rng(0)
    label = rng.integers(0, 13, 5000); label[rng.random(5000) < 0.05] = 255
    pred = np.where(rng.random(5000) < 0.7, label, rng.integers(0, 13, 5000))
    i, u, t = boundary_eval.intersection_and_union(torch.from_numpy(pred), torch.from_numpy(label), 13, 255)
    p = pred.copy(); p[label == 255] = 255
    inter = p[p == label]
    ai = np.bincount(inter[inter < 13], minlength=13)[:13]
    ao = np.bincount(p[p < 13], minlength=13)[:13]
    at = np.bincount(label[label < 13], minlength=13)[:13]
    assert np.array_equal(i.numpy(), ai) and np.array_equal(u.numpy(), ao + at - ai) and np.array_equal(t.numpy(), at)


def test_trainer_checkpoint_round_trip_and_schedule(tmp_path):
    """optimizer / scheduler / checkpoint format of the reference trainer (pytorch/tool/train.py:154-165,209-224,289-296)"""
    from contrastboundary_b200 import trainer
    torch.manual_seed(0)
    net = torch.nn.Sequential(model._LatentMLP(6, 8), torch.nn.Linear(8, 13))
    opt = trainer.build_optimizer(net, fused=False)
    sch = trainer.build_scheduler(opt, epochs=100)
    assert sch.milestones == {60: 1, 80: 1}
    x, y = torch.randn(32, 6), torch.randint(0, 13, (32,))
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(net(x), y).backward()
        opt.step()
        sch.step()
    path = str(tmp_path / "model" / "model_last.pth")
    trainer.save_checkpoint(path, 3, net, opt, sch, best_iou=0.42, is_best=True)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"epoch", "state_dict", "optimizer", "scheduler", "best_iou", "is_best"}     # train.py:292-293
    assert all(k.startswith("module.") for k in ckpt["state_dict"])                                  # DDP keys, test.py:107
    assert (tmp_path / "model" / "model_best.pth").exists()
    net2 = torch.nn.Sequential(model._LatentMLP(6, 8), torch.nn.Linear(8, 13))
    opt2 = trainer.build_optimizer(net2, fused=False)
    sch2 = trainer.build_scheduler(opt2, epochs=100)
    epoch, best = trainer.resume(path, net2, opt2, sch2)
    assert (epoch, best) == (3, 0.42) and sch2.last_epoch == 3
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    assert torch.equal(opt.state_dict()["state"][0]["momentum_buffer"], opt2.state_dict()["state"][0]["momentum_buffer"])
    net3 = torch.nn.Sequential(model._LatentMLP(6, 8), torch.nn.Linear(8, 13))
    assert trainer.load_weights(path, net3) == 3 and torch.equal(net3[1].weight, net[1].weight)
    # learning rate: 0.5 until epoch 60, 0.05 until 80, then 0.005
    for _ in range(3, 85):
        opt2.step(); sch2.step()
    assert abs(opt2.param_groups[0]["lr"] - 0.005) < 1e-9



def test_trainer_metrics_scalars_and_sync_bn_option(tmp_path):
    """train.py:148-149 (sync_bn), :265-284 (scalars), :328-364 (per-step metrics -> epoch averages)"""
    import json
    import torch.nn as nn
    rng = np.random.default_rng(0)
    acc = trainer.MetricsAccumulator(13, reduce_every=2)
    tot_n, tot_loss, ti, tu, tt = 0, np.zeros(6), np.zeros(13), np.zeros(13), np.zeros(13)
    for step in range(5):
        n = int(rng.integers(50, 90))
        logits = torch.from_numpy(rng.standard_normal((n, 13)).astype(np.float32))
        target = torch.from_numpy(rng.integers(0, 13, n))
        target[:3] = 255                                                     # ignored points
        loss = torch.from_numpy(rng.random(6).astype(np.float32))
        acc.update(trainer.pack_step_metrics(loss, logits, target, 13, 255))
        pred = logits.max(1)[1].numpy().copy()
        tg = target.numpy()
        pred[tg == 255] = 255                                               # util/common_util.py:40-52
        tot_n += n
        tot_loss += loss.numpy().astype(np.float64) * n
        for c in range(13):
            ti[c] += np.sum((pred == c) & (tg == c)); tt[c] += np.sum(tg == c)
            tu[c] += np.sum(pred == c) + np.sum(tg == c) - np.sum((pred == c) & (tg == c))
    loss_avg, miou, macc, allacc = acc.summary()
    assert np.allclose(loss_avg, tot_loss / tot_n, rtol=1e-12)
    assert abs(miou - np.mean(ti / (tu + 1e-10))) < 1e-12 and abs(macc - np.mean(ti / (tt + 1e-10))) < 1e-12
    assert abs(allacc - ti.sum() / (tt.sum() + 1e-10)) < 1e-12
    log = trainer.ScalarLog(str(tmp_path / "run"))
    log.log_epoch("train", loss_avg, miou, macc, allacc, 1)
    log.close()
    tags = [json.loads(l)["tag"] for l in open(tmp_path / "run" / "scalars.jsonl")]
    assert tags == ["loss_train"] + [f"loss_train_{i}" for i in range(6)] + ["mIoU_train", "mAcc_train", "allAcc_train"]
    m = trainer.build_model(sync_bn=True)
    assert any(isinstance(x, nn.SyncBatchNorm) for x in m.modules()) and not any(type(x) is nn.BatchNorm1d for x in m.modules())
    assert all(not x.fused for x in m.modules() if hasattr(x, "fused") and isinstance(x, (model.PointTransformerLayer, model.TransitionDown)))
    assert set(trainer.build_model().state_dict()) == set(m.state_dict())   # same checkpoint keys either way


def test_reference_constructor_signatures():
    """pointtransformer_seg_repro(c=, k=, config=) / Loss(config) / ContrastHead(head_cfg, config) take the reference's own
    config node (pytorch/util/config.py CfgNode of the shipped yaml; pointtransformer_seg.py:15-66,139-143, heads.py:66)"""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    refpy = next((p for p in ("/root/reference/pytorch", os.path.join(root, "baseline", "_ref", "pytorch"))
                  if os.path.exists(os.path.join(p, "util", "config.py"))), None)
    cfg_dict = {"base_fdim": 32, "nsample": [36, 24, 24, 24, 24], "nstride": [4, 4, 4, 4], "ignore_label": 255, "voxel_size": 0.04,
                "contrast": {"stage": "Ua", "contrast": "softnn", "ftype": "latent", "sample": "label", "pos": "cnt", "dist": "l2",
                             "temperature": 1, "weight": "w.1"},
                "multi": {"stage": "Ua", "ftype": "latent", "combine": "concat"}}
    if refpy is not None:
        sys.path.insert(0, refpy)
        from util.config import CfgNode
        cfg = CfgNode(json.loads(json.dumps(cfg_dict)), default="")
    else:
        cfg = cfg_dict
    m = model.pointtransformer_seg_repro(c=6, k=13, config=cfg)
    crit = model.Loss(cfg)
    head = model.ContrastHead(cfg["contrast"] if isinstance(cfg, dict) else cfg.contrast, cfg)
    assert m.cfg.nsample == [36, 24, 24, 24, 24] and m.cfg.contrast.weight == 0.1 and m.cfg.contrast.temperature == 1.0
    assert head.stages == [("up", i) for i in range(5)] and crit.contrast_head is not None
    assert len(m.state_dict()) == 960                                       # the reference network's parameter / buffer count
    bad = dict(cfg_dict, contrast=dict(cfg_dict["contrast"], contrast="nce"))
    import pytest
    with pytest.raises(NotImplementedError):
        model.Loss(bad)
    plain = {k: v for k, v in cfg_dict.items() if k not in ("contrast", "multi")}
    m2 = model.pointtransformer_seg_repro(c=6, k=13, config=plain)           # origin_4gpu.yaml: no heads, plain classifier
    assert m2.head is None and m2.cls is not None and model.Loss(plain).contrast_head is None


def test_cpu_tensors_take_torch_paths_and_the_wgrad_fork_is_opt_in():
    """Host logic around the CUDA-only kernels: on CPU tensors the drop-in wrappers fall through to torch (the KERNELS have no
    CPU fallback, the nn.Module-level wrappers stay usable for state_dict / shape work), and the weight-gradient fork is off
    unless engine._backward turns it on."""
    import torch
    from contrastboundary_b200 import linear_ops, model
    assert linear_ops._fork["on"] is False and not linear_ops._fork["used"]
    with linear_ops.wgrad_fork():
        assert linear_ops._fork["on"] is True
    assert linear_ops._fork["on"] is False
    torch.manual_seed(0)
    x = torch.randn(17, 8, requires_grad=True)
    lin = linear_ops.Linear(8, 5)
    ref = torch.nn.functional.linear(x, lin.weight, lin.bias)
    got = lin(x)
    assert torch.equal(got, ref)
    got.sum().backward()
    assert lin.weight.grad is not None and x.grad is not None
    logits = torch.randn(11, 13, requires_grad=True)
    target = torch.randint(0, 13, (11,))
    target[3] = 255
    a = model.cross_entropy(logits, target, 255)
    b = torch.nn.functional.cross_entropy(logits, target, ignore_index=255)
    assert torch.equal(a, b)
