"""The NumPy stand-in for the TF-1 API (tests/golden/tf_numpy_shim.py) is what the reference's TF source is executed on to make
tests/golden/tf_ops_ref.npz.  Its ops are checked here against an independent implementation of the documented TensorFlow
semantics — torch on the CPU — so that the pin does not rest on the stand-in's author reading the TF docs the same way twice."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tf_numpy_shim as tf  # noqa: E402


def test_batch_normalization_training_and_inference():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 7)) * 3 + 1
    tf.reset(seed=1)
    y = tf.layers.batch_normalization(x, momentum=0.99, epsilon=1e-3, training=True, name="bn")
    v = tf.variables()
    ref = F.batch_norm(torch.from_numpy(x), None, None, torch.from_numpy(v["bn/gamma"]), torch.from_numpy(v["bn/beta"]), True, 0.0, 1e-3)
    assert np.abs(y - ref.numpy()).max() < 1e-12                      # batch mean, BIASED variance
    y3 = tf.layers.batch_normalization(rng.standard_normal((4, 5, 7)), training=True, name="bn3")      # statistics over all but the last axis
    assert np.allclose(y3.reshape(-1, 7).mean(0), tf.variables()["bn3/beta"], atol=1e-12)
    yi = tf.layers.batch_normalization(x, epsilon=1e-3, training=False, name="bn")
    ref_i = F.batch_norm(torch.from_numpy(x), torch.from_numpy(v["bn/moving_mean"]), torch.from_numpy(v["bn/moving_variance"]),
                         torch.from_numpy(v["bn/gamma"]), torch.from_numpy(v["bn/beta"]), False, 0.0, 1e-3)
    assert np.abs(yi - ref_i.numpy()).max() < 1e-12


def test_losses_and_activations():
    rng = np.random.default_rng(1)
    logits, labels = rng.standard_normal((40, 13)) * 4, rng.integers(0, 13, 40)
    a = tf.nn.sparse_softmax_cross_entropy_with_logits(labels=labels, logits=logits)
    b = F.cross_entropy(torch.from_numpy(logits), torch.from_numpy(labels), reduction="none").numpy()
    assert np.abs(a - b).max() < 1e-12
    x = rng.standard_normal((6, 5, 4))
    assert np.abs(tf.nn.softmax(x, 1) - torch.softmax(torch.from_numpy(x), 1).numpy()).max() < 1e-12
    assert np.abs(tf.nn.leaky_relu(x, alpha=0.2) - F.leaky_relu(torch.from_numpy(x), 0.2).numpy()).max() == 0
    assert abs(tf.nn.l2_loss(x) - 0.5 * float((torch.from_numpy(x) ** 2).sum())) < 1e-12           # sum(x^2) / 2
    n = tf.nn.l2_normalize(x, axis=-1, epsilon=1e-12)
    assert np.abs(n - F.normalize(torch.from_numpy(x), dim=-1, eps=1e-6).numpy()).max() < 1e-9
    assert np.abs(tf.math.xlogy(np.array([0.0, 2.0]), np.array([0.0, 3.0])) - np.array([0.0, 2 * np.log(3.0)])).max() < 1e-15


def test_indexing_ops():
    rng = np.random.default_rng(2)
    p = rng.standard_normal((9, 3))
    idx = rng.integers(0, 9, (5, 4))
    assert np.array_equal(tf.gather(p, idx), p[idx])
    idx[0, 0] = 9                                                      # the one-past-the-end shadow index: GPU kernel -> zero row
    assert np.array_equal(tf.gather(p, idx)[0, 0], np.zeros(3)) and np.array_equal(tf.gather(p, idx)[1:], p[idx[1:]])
    oh = tf.one_hot(np.array([[0, 2], [-1, 1]]), depth=3)
    assert np.array_equal(oh, np.array([[[1, 0, 0], [0, 0, 1]], [[0, 0, 0], [0, 1, 0]]], float))       # -1 -> all zeros
    assert tf.argmax(np.array([[1, 3, 3], [2, 2, 0]]), axis=-1).tolist() == [1, 0]                     # first maximum
    m = np.array([True, False, True])
    assert np.array_equal(tf.boolean_mask(p[:3], m), p[:3][m])
    assert np.array_equal(tf.pad(np.array([[1, 2]]), [[0, 0], [0, 3]], "CONSTANT", constant_values=7), np.array([[1, 2, 7, 7, 7]]))
    assert np.array_equal(tf.tensordot(p, rng.standard_normal((3, 2)) * 0 + 1, 1), p.sum(1, keepdims=True).repeat(2, 1))
    assert tf.reduce_any(np.array([[False, True], [False, False]]), axis=-1, keepdims=True).tolist() == [[True], [False]]


def test_control_flow_scopes_and_variables():
    out = tf.while_loop(lambda i, acc: tf.less(i, 4), lambda i, acc: (i + 1, acc + i), [0, 0])
    assert out == [4, 6]
    assert tf.cond(np.array(True), true_fn=lambda: 1, false_fn=lambda: 2) == 1 and tf.cond(np.array(False), lambda: 1, lambda: 2) == 2
    tf.reset(seed=3)
    with tf.variable_scope("a"):
        with tf.variable_scope("b"):
            w = tf.get_variable("weights", [4, 2], initializer=tf.glorot_uniform_initializer())
        with tf.variable_scope("b"):                                   # re-entering a scope finds the same variable
            assert tf.get_variable("weights", [4, 2], initializer=tf.glorot_uniform_initializer()) is w
    assert list(tf.variables()) == ["a/b/weights"] and np.abs(w).max() <= np.sqrt(6 / 6)
    tf.add_to_collection("weight_losses", tf.multiply(tf.nn.l2_loss(w), 1e-3, name="weight_loss"))
    assert tf.taps()["l2_loss_variables"] == ["a/b/weights"] and len(tf.get_collection("weight_losses")) == 1
    assert abs(tf.add_n(tf.get_collection("weight_losses") + [1.0]) - (1.0 + 1e-3 * 0.5 * float((w ** 2).sum()))) < 1e-15
