"""A NumPy stand-in for the small part of the TensorFlow-1 API that the reference's ConvNet operators use, so that the
reference's OWN source files (tensorflow/models/local_aggregation_operators.py, basic_operators.py, ...) can be imported and
EXECUTED here, where TensorFlow is not installable, to produce golden vectors (tests/golden/make_golden_tf_ops.py).

Test infrastructure only.  Every function is the eager NumPy meaning of the TF-1 op of the same name (graph mode is not
modelled: the reference's functions are straight-line tensor code); variables live in a dictionary keyed by their
variable-scope path, created by the initializer the reference asks for from a seeded generator.  Tensors are float64 /
int64 NumPy arrays, so the goldens carry no float32 rounding of their own."""
import contextlib
import types

import numpy as np

__version__ = "1.15.0-numpy-shim"
float32, float16, float64, int32, int64, bool = np.float64, np.float64, np.float64, np.int64, np.int64, np.bool_


class _State:
    scope = []
    variables = {}
    taps = {}              # name -> last input of a named layer (e.g. the tensor entering 'pool_bn')
    rng = np.random.default_rng(0)
    collections = {}


def reset(seed=0):
    _State.scope, _State.variables, _State.taps, _State.collections = [], {}, {}, {}
    _State.rng = np.random.default_rng(seed)


def variables():
    return _State.variables


def preset(values):
    """start the next run from given variable values (get_variable returns an existing entry instead of initialising one)"""
    _State.variables = {k: np.array(v, dtype=np.float64) for k, v in values.items()}


def taps():
    return _State.taps


@contextlib.contextmanager
def variable_scope(name_or_scope, *args, **kwargs):
    pushed = isinstance(name_or_scope, str) and name_or_scope != ""
    if pushed:
        _State.scope.append(name_or_scope)
    try:
        yield "/".join(_State.scope)
    finally:
        if pushed:
            _State.scope.pop()


name_scope = variable_scope


@contextlib.contextmanager
def device(_):
    yield


def _path(name):
    return "/".join(_State.scope + [name])


def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True, **kwargs):
    key = _path(name)
    if key not in _State.variables:
        shape = tuple(int(s) for s in shape)
        _State.variables[key] = np.asarray(initializer(shape), dtype=np.float64).reshape(shape)
    return _State.variables[key]


# ---- initializers (values only matter in that they are generic: the tests read the variables back) -------------------
def glorot_uniform_initializer(**_):
    def init(shape):
        fan_in, fan_out = (shape[0], shape[-1]) if len(shape) > 1 else (shape[0], shape[0])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _State.rng.uniform(-lim, lim, shape)
    return init


def truncated_normal_initializer(stddev=1.0, **_):
    return lambda shape: np.clip(_State.rng.normal(0, stddev, shape), -2 * stddev, 2 * stddev)


def constant_initializer(value=0.0, **_):
    # a generic (non-zero) value around the requested constant: a zero bias would hide a missing bias term
    return lambda shape: np.full(shape, float(value)) + 0.1 * _State.rng.standard_normal(shape)


def zeros_initializer(**_):
    return lambda shape: 0.1 * _State.rng.standard_normal(shape)


def ones_initializer(**_):
    return lambda shape: 1.0 + 0.1 * _State.rng.standard_normal(shape)


def _variance_scaling_initializer(factor=2.0, mode="FAN_IN", uniform=False, **_):
    def init(shape):
        fan_in = shape[0] if len(shape) > 1 else shape[0]
        return _State.rng.normal(0, np.sqrt(factor / fan_in), shape)
    return init


def _vs(scale=1.0, **_):
    return _variance_scaling_initializer(factor=scale)


initializers = types.SimpleNamespace(glorot_uniform=glorot_uniform_initializer, glorot_normal=glorot_uniform_initializer,
                                     variance_scaling=_vs, zeros=zeros_initializer, ones=ones_initializer)
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(variance_scaling_initializer=_variance_scaling_initializer))


# ---- tensor ops ----------------------------------------------------------------------------------------------------------
def shape(x):
    return np.shape(x)


def concat(values, axis=0, **_):
    return np.concatenate([np.asarray(v) for v in values], axis=axis)


def zeros_like(x, dtype=None, **_):
    return np.zeros_like(x, dtype=dtype)


def ones_like(x, dtype=None, **_):
    return np.ones_like(x, dtype=dtype)


def zeros(shape, dtype=np.float64, **_):
    return np.zeros(shape, dtype=dtype)


def ones(shape, dtype=np.float64, **_):
    return np.ones(shape, dtype=dtype)


def constant(value, dtype=None, **_):
    return np.asarray(value, dtype=dtype)


Tensor = np.ndarray


def TensorShape(dims):
    return tuple(dims)


def gather(params, indices, axis=0, batch_dims=0, **_):
    """tf.gather, GPU-kernel semantics for an out-of-range index: the output row is zero (the CPU kernel raises).  The
    reference relies on it: multiscale_head gathers with the one-past-the-end shadow index of a radius search that found
    nothing (heads/head.py:455-458), and it trains on GPU."""
    assert batch_dims == 0 and axis == 0
    params, idx = np.asarray(params), np.asarray(indices).astype(np.int64)
    n = params.shape[0]
    assert idx.min() >= 0 and idx.max() <= n, "only the one-past-the-end shadow index is expected out of range"
    if idx.max() < n:
        return params[idx]
    return np.concatenate([params, np.zeros_like(params[:1])], 0)[idx]


def one_hot(indices, depth, axis=-1, dtype=np.float64, **_):
    idx = np.asarray(indices).astype(np.int64)
    out = (idx[..., None] == np.arange(int(depth))).astype(dtype)        # out-of-range (e.g. -1) -> all zeros, as TF
    assert axis in (-1, idx.ndim)
    return out


def argmax(x, axis=None, **_):
    return np.argmax(x, axis=axis)                                        # first maximum, as TF


def boolean_mask(tensor, mask, name=None, axis=None, **_):
    tensor, mask = np.asarray(tensor), np.asarray(mask).astype(np.bool_)
    return tensor[mask]


def reduce_any(x, axis=None, keepdims=False, **_):
    return np.any(x, axis=axis, keepdims=keepdims)


def reduce_all(x, axis=None, keepdims=False, **_):
    return np.all(x, axis=axis, keepdims=keepdims)


def cond(pred, true_fn=None, false_fn=None, name=None, **_):
    return true_fn() if np.asarray(pred).all() else false_fn()


def while_loop(cond, body, loop_vars, shape_invariants=None, name=None, **_):
    vars_ = list(loop_vars)
    while np.asarray(cond(*vars_)).all():
        vars_ = list(body(*vars_))
    return vars_


def map_fn(fn, elems, dtype=None, **_):
    return np.stack([fn(e) for e in elems])


def range(start, limit=None, delta=1, dtype=None, **_):          # noqa: A001  (tf.range)
    return np.arange(start, limit, delta) if limit is not None else np.arange(start)


def pad(tensor, paddings, mode="CONSTANT", constant_values=0, **_):
    assert mode == "CONSTANT"
    return np.pad(np.asarray(tensor), [(int(a), int(b)) for a, b in paddings], constant_values=constant_values)


def stack(values, axis=0, **_):
    return np.stack(values, axis=axis)


def stop_gradient(x, **_):
    return x


not_equal, pow = np.not_equal, np.power


def _xlogy(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return np.where(x == 0, 0.0, x * np.log(np.where(x == 0, 1.0, y)))


math = types.SimpleNamespace(xlogy=_xlogy, log=np.log, exp=np.exp)


def expand_dims(x, axis, **_):
    return np.expand_dims(x, axis)


def squeeze(x, axis=None, **_):
    return np.squeeze(x, axis=axis)


def tile(x, multiples, **_):
    return np.tile(x, [int(m) for m in multiples])


def reshape(x, shape, **_):
    return np.reshape(x, [int(s) for s in shape])


def transpose(x, perm=None, **_):
    return np.transpose(x, perm)


def cast(x, dtype, **_):
    return np.asarray(x).astype(dtype)


sqrt, square, exp, log, abs, maximum, minimum = np.sqrt, np.square, np.exp, np.log, np.abs, np.maximum, np.minimum
def multiply(x, y, name=None):
    return np.multiply(x, y)


add, subtract, divide = np.add, np.subtract, np.divide
less, greater, greater_equal, less_equal, equal, logical_and, logical_or, logical_not = (
    np.less, np.greater, np.greater_equal, np.less_equal, np.equal, np.logical_and, np.logical_or, np.logical_not)


def where(cond, x=None, y=None, **_):
    return np.where(cond, x, y)


def reduce_sum(x, axis=None, keepdims=False, **_):
    return np.sum(x, axis=axis, keepdims=keepdims)


def reduce_max(x, axis=None, keepdims=False, **_):
    return np.max(x, axis=axis, keepdims=keepdims)


def reduce_min(x, axis=None, keepdims=False, **_):
    return np.min(x, axis=axis, keepdims=keepdims)


def reduce_mean(x, axis=None, keepdims=False, **_):
    return np.mean(x, axis=axis, keepdims=keepdims)


def matmul(a, b, **_):
    return np.matmul(a, b)


def tensordot(a, b, axes, **_):
    return np.tensordot(a, b, axes)


def add_to_collection(name, value):
    _State.collections.setdefault(name, []).append(value)


def get_collection(name, scope=None):
    assert scope is None
    return list(_State.collections.get(name, []))


def add_n(inputs, name=None):
    out = 0.0
    for v in inputs:
        out = out + v
    return out


accumulate_n = add_n


def _relu(x, **_):
    return np.maximum(x, 0)


def _leaky_relu(x, alpha=0.2, **_):
    return np.where(x > 0, x, alpha * x)


def _softmax(x, axis=-1, **_):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def _l2_loss(x, **_):
    for name, v in _State.variables.items():            # remember WHICH variables the reference regularises
        if v is x:
            _State.taps.setdefault("l2_loss_variables", []).append(name)
    return 0.5 * np.sum(np.square(x))


def _l2_normalize(x, axis=-1, epsilon=1e-12, **_):
    return x / np.sqrt(np.maximum(np.sum(np.square(x), axis=axis, keepdims=True), epsilon))


def _sparse_xent(labels=None, logits=None, name=None, **_):
    z = logits - logits.max(axis=-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=-1, keepdims=True))
    return -logp[np.arange(logp.shape[0]), np.asarray(labels).astype(np.int64)]


nn = types.SimpleNamespace(relu=_relu, leaky_relu=_leaky_relu, softmax=_softmax, l2_loss=_l2_loss, l2_normalize=_l2_normalize,
                           sparse_softmax_cross_entropy_with_logits=_sparse_xent)


def _batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, training=False, trainable=True, name=None, fused=None, **_):
    """tf.layers.batch_normalization: statistics over every axis but the last; biased variance in training"""
    x = np.asarray(inputs, dtype=np.float64)
    c = x.shape[-1]
    with variable_scope(name or "batch_normalization"):
        _State.taps[_path("input")] = x
        gamma = get_variable("gamma", [c], ones_initializer())
        beta = get_variable("beta", [c], zeros_initializer())
        mm = get_variable("moving_mean", [c], zeros_initializer())
        mv = get_variable("moving_variance", [c], lambda s: 1.0 + 0.1 * _State.rng.random(s))
    red = tuple(range(x.ndim - 1))
    if training:
        mean, var = x.mean(axis=red), x.var(axis=red)
    else:
        mean, var = mm, mv
    return (x - mean) / np.sqrt(var + epsilon) * gamma + beta


layers = types.SimpleNamespace(batch_normalization=_batch_normalization)
