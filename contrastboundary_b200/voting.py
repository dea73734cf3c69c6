"""SURVEY §8(f) row 3 (remainder) — test-time voting of the reference (pytorch/tool/test.py:128-148,157-239,330-335):
every full-resolution room is enumerated voxel by voxel (`data_load`), rooms larger than `voxel_max` are covered by
"spatially regular" nearest-`voxel_max` crops driven by a potential field (`test`, :196-216), the crops are batched,
pushed through the network and their logits accumulated per original point (`cumulate_probs`).

Everything stays on the device: the voxel enumeration comes from cb_voxelize, the crop loop is argmin / distance /
top-k over device tensors, logits are accumulated with index_add_.  The only host reads are the loop conditions
(the reference's loop has the same data-dependent trip counts)."""
import torch

from . import dataprep


def enumerate_voxel_points(coord, voxel_size):
    """idx_data of data_load (test.py:135-146): the i-th list holds, for every occupied voxel, its (i mod count)-th point.
    coord: (n,3) tensor already shifted to min = 0 (as data_load does).  -> list of int64 index tensors (n_voxels each)"""
    if not voxel_size:
        return [torch.arange(coord.shape[0], device=coord.device)]
    idx_sort, count = dataprep.voxelize(coord, voxel_size, mode=1)
    start = torch.cumsum(count, 0) - count
    return [idx_sort[start + (i % count)] for i in range(int(count.max()))]


def regular_crops(coord_part, voxel_max, generator=None, potentials=None):
    """the `while idx_uni.size != idx_part.shape[0]` loop of test.py:199-216 for one enumeration of a room.
    Yields (idx_crop int64 (voxel_max,), coord_sub (voxel_max,3) shifted to min 0).  `potentials`: initial values
    (the reference: np.random.rand(n) * 1e-3); default = torch.rand with `generator`."""
    n = coord_part.shape[0]
    dev = coord_part.device
    if potentials is None:
        potentials = torch.rand(n, device=dev, generator=generator, dtype=coord_part.dtype) * 1e-3
    coord_p = potentials.clone()
    covered = torch.zeros(n, dtype=torch.bool, device=dev)
    while not bool(covered.all()):
        init_idx = torch.argmin(coord_p)                                        # centre = lowest potential
        dist = ((coord_part - coord_part[init_idx]) ** 2).sum(1)
        d_crop, idx_crop = torch.topk(dist, voxel_max, largest=False, sorted=True)   # argsort(dist)[:voxel_max]
        coord_p[idx_crop] += (1 - d_crop / d_crop.max()) ** 2                   # update potentials
        covered[idx_crop] = True
        sub = coord_part[idx_crop]
        yield idx_crop, sub - sub.min(0)[0]                                     # input_normalize


@torch.no_grad()
def vote_room(model, coord, feat, num_classes, voxel_size=0.04, voxel_max=None, batch_size_test=10, generator=None):
    """accumulated logits (n, num_classes) of one room (test.py:157-239).  coord (n,3), feat (n,3) in 0..255, device tensors.
    `model(inputs) -> (logits, stage_list)` with inputs = {'points','features','offset'} (the reference's model call)."""
    dev = coord.device
    coord = coord - coord.min(0)[0]                                             # data_load :137-138
    cum = torch.zeros((coord.shape[0], num_classes), dtype=torch.float32, device=dev)
    samples = []                                                                # (idx into the room, coord, feat)
    for idx_part in enumerate_voxel_points(coord, voxel_size):
        c, f = coord[idx_part], feat[idx_part]
        if voxel_max and c.shape[0] > voxel_max:
            for idx_crop, c_sub in regular_crops(c, voxel_max, generator):
                samples.append((idx_part[idx_crop], c_sub, f[idx_crop] / 255.0))
        else:
            samples.append((idx_part, c - c.min(0)[0], f / 255.0))
    for s in range(0, len(samples), batch_size_test):
        chunk = samples[s:s + batch_size_test]
        inds = torch.cat([c[0] for c in chunk])
        offset = torch.cumsum(torch.tensor([c[0].numel() for c in chunk]), 0).to(torch.int32).to(dev)
        inputs = {"points": torch.cat([c[1] for c in chunk]).float().contiguous(),
                  "features": torch.cat([c[2] for c in chunk]).float().contiguous(), "offset": offset,
                  "offset_host": [int(v) for v in offset.tolist()]}
        pred, _ = model(inputs)
        cum.index_add_(0, inds, pred.float())                                   # cumulate_probs :333
    return cum
