"""world_size-2 gloo test (CPU) of the N>1 host logic: data-parallel sharding of scenes by rank, DDP gradient
averaging through the network's modules, and the max-over-ranks step timing used by bench.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from contrastboundary_b200 import model, synthetic
    torch.manual_seed(0)
    # a CPU-capable slice of the network: the MultiHead latent MLP + classifier (our Linear subclass falls back
    # to torch off-GPU; the fused operators themselves are CUDA-only by design)
    net = torch.nn.Sequential(model._LatentMLP(6, 8), torch.nn.Linear(8, 13))
    ddp = DDP(net)
    # scenes shard by rank: rank r gets its own seeds (bench.py: 5000 + 97 * rank + i)
    b = synthetic.make_batch(1, [256], 5000 + 97 * rank)
    x = torch.from_numpy(np.concatenate([b["points"], b["features"]], 1))
    y = torch.from_numpy(b["point_labels"])
    loss = torch.nn.functional.cross_entropy(ddp(x), y)
    loss.backward()
    g = net[1].weight.grad.clone()
    # reference: average of the per-rank gradients computed without DDP
    net2 = torch.nn.Sequential(model._LatentMLP(6, 8), torch.nn.Linear(8, 13))
    net2.load_state_dict(net.state_dict())
    net2.zero_grad()
    torch.nn.functional.cross_entropy(net2(x), y).backward()
    local = net2[1].weight.grad.clone()
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    local /= world
    assert torch.allclose(g, local, rtol=1e-5, atol=1e-7), "DDP gradient != mean of rank gradients"
    # max-over-ranks timing
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert abs(float(t) - 0.010 * world) < 1e-12
    # distinct shards
    h = torch.tensor([float(x.sum())])
    hs = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(hs, h)
    assert len({round(float(v), 4) for v in hs}) == world, "ranks got identical scenes"
    # one coalesced metrics all-reduce per step (trainer.MetricsAccumulator) == the reference's five (train.py:328-338)
    from contrastboundary_b200 import trainer
    logits = ddp(x).detach()
    lossv = torch.stack([torch.nn.functional.cross_entropy(logits, y), torch.tensor(0.25 * (rank + 1))])
    acc = trainer.MetricsAccumulator(13)
    acc.update(trainer.pack_step_metrics(lossv, logits, y, 13))
    loss_avg, miou, macc, allacc = acc.summary()
    n = torch.tensor([float(len(y))])
    ln = lossv.double() * len(y)
    dist.all_reduce(n), dist.all_reduce(ln)
    assert np.allclose(loss_avg, (ln / n).numpy(), rtol=1e-12)
    pred = logits.max(1)[1]
    inter = torch.bincount(pred[pred == y], minlength=13).double()
    tgt = torch.bincount(y, minlength=13).double()
    dist.all_reduce(inter), dist.all_reduce(tgt)
    assert abs(allacc - float(inter.sum() / (tgt.sum() + 1e-10))) < 1e-12
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
