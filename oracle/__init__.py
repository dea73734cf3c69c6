"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy-facing loaders for the CPU restatements of the reference's operators
(`pointops_oracle.c`, `tfops_oracle.cpp`) and, when present, for the reference's own sources
compiled unmodified into `oracle/_ref/` (`build_ref.sh`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.  The product (`contrastboundary_b200/`) never does.

Parity status: the reference has no golden vectors for this path.  The
restatements are pinned against (i) `oracle/_ref/libref_cpu.so` / `libref_cpy.so` (the
reference's C++ cores, run in the build container; vectors in tests/golden/tf_*.npz) and
(ii) `oracle/_ref/pointops_cuda.so` (the reference's CUDA kernels for sm_100a, run on the B200
box; vectors in tests/golden/pointops_*.npz).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C/C++ restatements (and oracle/_ref when /root/reference exists)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, f))
                            for f in ("liboracle_pointops.so", "liboracle_tfops.so"))
    if need or any(os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(os.path.join(_HERE, l))
                   for s, l in (("pointops_oracle.c", "liboracle_pointops.so"),
                                ("tfops_oracle.cpp", "liboracle_tfops.so"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    if os.path.isdir(os.environ.get("REF_ROOT", "/root/reference")):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])


_libs = {}


def _lib(name):
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _libs[name] = C.CDLL(path)
    return _libs[name]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------------------------
# pointops restatement (liboracle_pointops.so)
# ----------------------------------------------------------------------------------------------

def num_threads():
    return int(_lib("liboracle_pointops.so").oracle_num_threads())


def set_num_threads(t):
    _lib("liboracle_pointops.so").oracle_set_num_threads(int(t))


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """-> idx (m,nsample) int32, dist2 (m,nsample) float32 — SQUARED distances (kernel output)."""
    lib = _lib("liboracle_pointops.so")
    xyz, offset, new_offset = _f32(xyz), _i32(offset), _i32(new_offset)
    new_xyz = xyz if new_xyz is None else _f32(new_xyz)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    d2 = np.zeros((m, nsample), np.float32)
    lib.oracle_knnquery.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _i32p, _i32p, _i32p, _f32p]
    lib.oracle_knnquery(m, int(nsample), xyz, new_xyz, offset, new_offset, idx, d2)
    return idx, d2


def furthestsampling(xyz, offset, new_offset):
    lib = _lib("liboracle_pointops.so")
    xyz, offset, new_offset = _f32(xyz), _i32(offset), _i32(new_offset)
    n, b = xyz.shape[0], offset.shape[0]
    lens = np.diff(np.concatenate([[0], offset]))
    n_max = int(lens.max())
    idx = np.zeros(int(new_offset[-1]), np.int32)
    tmp = np.full(n, 1e10, np.float32)
    lib.oracle_furthestsampling.argtypes = [C.c_int, C.c_int, _f32p, _i32p, _i32p, _f32p, _i32p]
    lib.oracle_furthestsampling(b, n_max, xyz, offset, new_offset, tmp, idx)
    return idx


def grouping_forward(inp, idx):
    lib = _lib("liboracle_pointops.so")
    inp, idx = _f32(inp), _i32(idx)
    m, k = idx.shape
    c = inp.shape[1]
    out = np.zeros((m, k, c), np.float32)
    lib.oracle_grouping_forward.argtypes = [C.c_int] * 3 + [_f32p, _i32p, _f32p]
    lib.oracle_grouping_forward(m, k, c, inp, idx, out)
    return out


def grouping_backward(grad_out, idx, n):
    lib = _lib("liboracle_pointops.so")
    grad_out, idx = _f32(grad_out), _i32(idx)
    m, k, c = grad_out.shape
    gi = np.zeros((n, c), np.float32)
    lib.oracle_grouping_backward.argtypes = [C.c_int] * 3 + [_f32p, _i32p, _f32p]
    lib.oracle_grouping_backward(m, k, c, grad_out, idx, gi)
    return gi


def subtraction_forward(a, b, idx):
    lib = _lib("liboracle_pointops.so")
    a, b, idx = _f32(a), _f32(b), _i32(idx)
    n, c = a.shape
    k = idx.shape[1]
    out = np.zeros((n, k, c), np.float32)
    lib.oracle_subtraction_forward.argtypes = [C.c_int] * 3 + [_f32p, _f32p, _i32p, _f32p]
    lib.oracle_subtraction_forward(n, k, c, a, b, idx, out)
    return out


def subtraction_backward(idx, grad_out):
    lib = _lib("liboracle_pointops.so")
    grad_out, idx = _f32(grad_out), _i32(idx)
    n, k, c = grad_out.shape
    g1 = np.zeros((n, c), np.float32)
    g2 = np.zeros((n, c), np.float32)
    lib.oracle_subtraction_backward.argtypes = [C.c_int] * 3 + [_i32p, _f32p, _f32p, _f32p]
    lib.oracle_subtraction_backward(n, k, c, idx, grad_out, g1, g2)
    return g1, g2


def aggregation_forward(inp, position, weight, idx):
    lib = _lib("liboracle_pointops.so")
    inp, position, weight, idx = _f32(inp), _f32(position), _f32(weight), _i32(idx)
    n, k, c = position.shape
    w_c = weight.shape[-1]
    out = np.zeros((n, c), np.float32)
    lib.oracle_aggregation_forward.argtypes = [C.c_int] * 4 + [_f32p, _f32p, _f32p, _i32p, _f32p]
    lib.oracle_aggregation_forward(n, k, c, w_c, inp, position, weight, idx, out)
    return out


def aggregation_backward(inp, position, weight, idx, grad_out):
    lib = _lib("liboracle_pointops.so")
    inp, position, weight, idx, grad_out = _f32(inp), _f32(position), _f32(weight), _i32(idx), _f32(grad_out)
    n, k, c = position.shape
    w_c = weight.shape[-1]
    gi = np.zeros_like(inp)
    gp = np.zeros_like(position)
    gw = np.zeros_like(weight)
    lib.oracle_aggregation_backward.argtypes = [C.c_int] * 4 + [_f32p, _f32p, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p]
    lib.oracle_aggregation_backward(n, k, c, w_c, inp, position, weight, idx, grad_out, gi, gp, gw)
    return gi, gp, gw


def interpolation_forward(inp, idx, weight):
    lib = _lib("liboracle_pointops.so")
    inp, idx, weight = _f32(inp), _i32(idx), _f32(weight)
    n, k = idx.shape
    c = inp.shape[1]
    out = np.zeros((n, c), np.float32)
    lib.oracle_interpolation_forward.argtypes = [C.c_int] * 3 + [_f32p, _i32p, _f32p, _f32p]
    lib.oracle_interpolation_forward(n, c, k, inp, idx, weight, out)
    return out


def interpolation_backward(grad_out, idx, weight, m):
    lib = _lib("liboracle_pointops.so")
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    n, k = idx.shape
    c = grad_out.shape[1]
    gi = np.zeros((m, c), np.float32)
    lib.oracle_interpolation_backward.argtypes = [C.c_int] * 3 + [_f32p, _i32p, _f32p, _f32p]
    lib.oracle_interpolation_backward(n, c, k, grad_out, idx, weight, gi)
    return gi


# ----------------------------------------------------------------------------------------------
# TF-side restatement (liboracle_tfops.so) and the compiled reference cores (oracle/_ref)
# ----------------------------------------------------------------------------------------------

def _grid_subsample(lib, fn, points, features, labels, dl):
    points = _f32(points)
    n = points.shape[0]
    feats = None if features is None else _f32(features).reshape(n, -1)
    labs = None if labels is None else _i32(labels).reshape(n, -1)
    fdim = 0 if feats is None else feats.shape[1]
    ldim = 0 if labs is None else labs.shape[1]
    op = np.zeros((n, 3), np.float32)
    of = np.zeros((n, max(fdim, 1)), np.float32)
    oc = np.zeros((n, max(ldim, 1)), np.int32)
    f = getattr(lib, fn)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float,
                  C.c_void_p, C.c_void_p, C.c_void_p]
    m = f(points.ctypes.data, n, None if feats is None else feats.ctypes.data, fdim,
          None if labs is None else labs.ctypes.data, ldim, C.c_float(dl), op.ctypes.data, of.ctypes.data,
          oc.ctypes.data)
    out = [op[:m].copy()]
    if feats is not None:
        out.append(of.reshape(-1)[:m * fdim].reshape(m, fdim).copy())
    if labs is not None:
        lab = oc.reshape(-1)[:m * ldim].reshape(m, ldim).copy()
        out.append(lab[:, 0] if np.ndim(labels) == 1 else lab)
    return out[0] if len(out) == 1 else tuple(out)


def grid_subsampling(points, features=None, labels=None, sampleDl=0.1):
    return _grid_subsample(_lib("liboracle_tfops.so"), "oracle_grid_subsample", points, features, labels, sampleDl)


def _batch_grid(lib, fn, points, batches, dl):
    points, batches = _f32(points), _i32(batches)
    n, b = points.shape[0], batches.shape[0]
    op = np.zeros((n, 3), np.float32)
    ob = np.zeros(b, np.int32)
    f = getattr(lib, fn)
    f.restype = C.c_int
    f.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_float, _f32p, _i32p]
    m = f(points, n, batches, b, C.c_float(dl), op, ob)
    return op[:m].copy(), ob


def batch_grid_subsampling(points, batches_len, sampleDl):
    return _batch_grid(_lib("liboracle_tfops.so"), "oracle_batch_grid_subsample", points, batches_len, sampleDl)


def _batch_neighbors(lib, fn, free_fn, queries, supports, q_batches, s_batches, radius):
    queries, supports = _f32(queries), _f32(supports)
    q_batches, s_batches = _i32(q_batches), _i32(s_batches)
    f = getattr(lib, fn)
    f.restype = C.POINTER(C.c_int)
    f.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, _i32p, C.c_int, C.c_float, C.POINTER(C.c_int)]
    mc = C.c_int(0)
    nq = queries.shape[0]
    p = f(queries, nq, supports, supports.shape[0], q_batches, s_batches, q_batches.shape[0],
          C.c_float(radius), C.byref(mc))
    out = np.ctypeslib.as_array(p, shape=(max(nq * mc.value, 1),))[:nq * mc.value].reshape(nq, mc.value).copy()
    fr = getattr(lib, free_fn)
    fr.argtypes = [C.c_void_p]
    fr(p)
    return out


def batch_neighbors(queries, supports, q_batches, s_batches, radius):
    return _batch_neighbors(_lib("liboracle_tfops.so"), "oracle_batch_radius_neighbors", "oracle_free",
                            queries, supports, q_batches, s_batches, radius)


def knn_batch(supports, queries, k):
    lib = _lib("liboracle_tfops.so")
    supports, queries = _f32(supports), _f32(queries)
    B, N, _ = supports.shape
    M = queries.shape[1]
    out = np.zeros((B, M, k), np.int64)
    lib.oracle_knn_batch.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i64p]
    lib.oracle_knn_batch(supports, queries, B, N, M, int(k), out)
    return out


# ---- the reference itself (oracle/_ref), CPU side -------------------------------------------

def have_ref_cpu():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cpu.so"))


def have_ref_gpu():
    return os.path.exists(os.path.join(_HERE, "_ref", "pointops_cuda.so"))


def ref_grid_subsampling(points, features=None, labels=None, sampleDl=0.1):
    return _grid_subsample(_lib("_ref/libref_cpy.so"), "ref_grid_subsample", points, features, labels, sampleDl)


def ref_batch_grid_subsampling(points, batches_len, sampleDl):
    return _batch_grid(_lib("_ref/libref_cpu.so"), "ref_batch_grid_subsample", points, batches_len, sampleDl)


def ref_batch_neighbors(queries, supports, q_batches, s_batches, radius):
    return _batch_neighbors(_lib("_ref/libref_cpu.so"), "ref_batch_radius_neighbors", "ref_free",
                            queries, supports, q_batches, s_batches, radius)


def ref_pointops_cuda():
    """Import the reference's own CUDA extension (GPU box only)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = os.path.join(_HERE, "_ref", "pointops_cuda.so")
    spec = importlib.util.spec_from_file_location("pointops_cuda", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
