"""SURVEY §8(f) row 3 — the test-time boundary evaluation of the reference (pytorch/tool/test.py:250-257,392-428):
K nearest neighbours over FULL-RESOLUTION rooms (10^5-10^6 points, kr in {16, 32, 64}; the reference's brute-force
kernel is O(N^2) there), `get_boundary_mask` (pytorch/model/basic_operators.py:69-97) and the per-mask
intersection / union / target histograms (pytorch/util/common_util.py:40-52) — all on the device."""
import ctypes as C

import torch

from . import _lib as L
from . import pointops


def room_neighbors(coord, kr):
    """neighbor_idx (n, kr) int32 of one full-resolution room (test.py:250-257)"""
    xyz = coord.contiguous().float()
    off = torch.tensor([xyz.shape[0]], dtype=torch.int32, device=xyz.device)
    idx, _ = pointops.knn_raw(int(kr), xyz, xyz, off, off, True)
    return idx


def get_boundary_mask(labels, neighbor_label=None, neighbor_idx=None, valid_mask=None, get_plain=False, get_cnt=False):
    """same signature and results as the reference's get_boundary_mask for 1-D integer labels with neighbor_idx given
    (the form test.py uses); `neighbor_label` (pre-gathered labels) is not needed on this path and must be None"""
    if neighbor_label is not None or neighbor_idx is None:
        raise L.CbopsError("boundary_eval.get_boundary_mask: pass neighbor_idx (the gather is fused into the kernel)")
    L.require_cuda(labels, neighbor_idx)
    lab = labels.contiguous().long()
    idx = neighbor_idx.contiguous().int()
    n, kr = idx.shape
    dev = lab.device
    vm = valid_mask.contiguous().to(torch.uint8) if valid_mask is not None else None
    cnt = torch.empty(n, dtype=torch.int32, device=dev) if get_cnt else None
    bound = torch.empty(n, dtype=torch.uint8, device=dev) if not get_cnt else None
    plain = torch.empty(n, dtype=torch.uint8, device=dev) if get_plain else None
    rc = L.lib().cb_boundary_mask(C.c_longlong(n), C.c_int(kr), L.ptr(lab), L.ptr(idx), L.ptr(vm), L.ptr(cnt), L.ptr(bound),
                                  L.ptr(plain), L.stream())
    L.check(rc, "cb_boundary_mask")
    out = cnt.long() if get_cnt else bound.bool()
    if get_plain:
        return out, plain.bool()
    return out


def intersection_and_union(pred, target, k, ignore_index=255):
    """area_intersection, area_union, area_target (k each) — intersectionAndUnionGPU, util/common_util.py:40-52"""
    pred = pred.reshape(-1).clone()
    target = target.reshape(-1)
    pred[target == ignore_index] = ignore_index
    inter = pred[pred == target]
    ai = torch.bincount(inter[(inter >= 0) & (inter < k)], minlength=k)[:k]
    ao = torch.bincount(pred[(pred >= 0) & (pred < k)], minlength=k)[:k]
    at = torch.bincount(target[(target >= 0) & (target < k)], minlength=k)[:k]
    return ai, ao + at - ai, at


def boundary_iou(pred, label, coord, krs=(16, 32, 64), num_classes=13, ignore_label=255):
    """{kr: {'bound-i','bound-u','bound-t','plain-i','plain-u','plain-t'}} for one room (test.py:392-412)"""
    out = {}
    for kr in krs:
        idx = room_neighbors(coord, kr)
        bound, plain = get_boundary_mask(label, neighbor_idx=idx, get_plain=True)
        d = {}
        for name, mask in (("bound", bound), ("plain", plain)):
            i, u, t = intersection_and_union(pred[mask], label[mask], num_classes, ignore_label)
            d[f"{name}-i"], d[f"{name}-u"], d[f"{name}-t"] = i, u, t
        out[kr] = d
    return out
