"""oracle/boundary.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
NumPy restatement of get_boundary_mask (pytorch/model/basic_operators.py:69-97) for 1-D integer labels and a
neighbour-index matrix; pinned by tests/golden/boundary_ref.npz, which tests/golden/make_golden_boundary.py produced by
importing the REAL reference function."""
import numpy as np


def get_boundary_mask(labels, neighbor_idx, valid_mask=None, get_plain=False, get_cnt=False):
    neighbor_label = labels[neighbor_idx.reshape(-1).astype(np.int64)].reshape(neighbor_idx.shape)   # :74-75
    valid_neighbor = neighbor_label >= 0                                                             # :78
    lab = labels[:, None]
    neq = np.logical_and(lab != neighbor_label, valid_neighbor)                                      # :81-82
    if get_cnt:
        bound = neq.sum(-1)                                                                          # :84
        bound = bound * valid_mask if valid_mask is not None else bound
    else:
        bound = neq.any(-1)                                                                          # :87
        bound = np.logical_and(bound, valid_mask) if valid_mask is not None else bound
    if get_plain:
        eq = np.logical_or(lab == neighbor_label, np.logical_not(valid_neighbor))                    # :93-94
        plain = eq.all(-1)
        plain = np.logical_and(plain, valid_mask) if valid_mask is not None else plain
        return bound, plain
    return bound
