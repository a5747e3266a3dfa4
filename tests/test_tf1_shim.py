"""The TF1-API shim that generated tests/golden/*.npz (tests/golden/tf1_shim.py), op by op against independent NumPy
statements of the TensorFlow 1.x documented semantics -- the ops the reference's model.py / aggregators.py call.
This is the "op numerics pinned against TF's documented semantics" half of the oracle's pin."""
import math

import numpy as np
import pytest

from tests.golden import tf1_shim as tf


def run(node, feed=None):
    return tf.Session().run(node, feed or {})


def ph(arr, dtype="float32"):
    p = tf.placeholder(dtype, shape=list(arr.shape), name="x")
    return p, {p: arr}


RNG = np.random.RandomState(0)


def test_matmul_batches_over_leading_dims():
    """model.py:214: [B, m, d, d] x [B, m, d, 1] -> [B, m, d, 1] (R.h, not R^T.h); :312: plain 2-D."""
    R = RNG.randn(3, 5, 4, 4).astype(np.float32)
    h = RNG.randn(3, 5, 4, 1).astype(np.float32)
    pr, fr = ph(R)
    p_h, fh = ph(h)
    got = run(tf.matmul(pr, p_h), {**fr, **fh})
    want = np.einsum("bmij,bmjk->bmik", R.astype(np.float64), h.astype(np.float64))
    assert got.shape == (3, 5, 4, 1) and np.allclose(got, want, atol=1e-5)
    assert not np.allclose(got, np.einsum("bmji,bmjk->bmik", R, h), atol=1e-3)


def test_softmax_axis_and_gather_axis0():
    x = RNG.randn(2, 3, 7).astype(np.float32)
    p, f = ph(x)
    e = np.exp(x - x.max(-1, keepdims=True))
    assert np.allclose(run(tf.nn.softmax(p), f), e / e.sum(-1, keepdims=True), atol=1e-6)       # default: last axis
    assert np.allclose(run(tf.nn.softmax(p, dim=-1), f), e / e.sum(-1, keepdims=True), atol=1e-6)
    e1 = np.exp(x - x.max(1, keepdims=True))
    assert np.allclose(run(tf.nn.softmax(p, dim=1), f), e1 / e1.sum(1, keepdims=True), atol=1e-6)
    table = RNG.randn(11, 4).astype(np.float32)
    ids = RNG.randint(0, 11, size=(2, 5)).astype(np.int64)
    pi, fi = ph(ids, "int64")
    pt, ft = ph(table)
    assert np.array_equal(run(tf.nn.embedding_lookup(pt, pi), {**fi, **ft}), table[ids])         # [2, 5, 4]
    adj = RNG.randint(0, 11, size=(11, 3)).astype(np.int64)                                       # constant (model.py:251)
    assert np.array_equal(run(tf.gather(adj, pi), fi), adj[ids])
    ids32 = ids.astype(np.int32)                                                                  # int32 feed (memories)
    p32, f32 = ph(ids32, "int32")
    assert np.array_equal(run(tf.gather(pt, p32), {**f32, **ft}), table[ids])


def test_reductions_and_losses():
    x = RNG.randn(4, 6).astype(np.float32)
    p, f = ph(x)
    assert np.isclose(run(tf.reduce_mean(p), f), x.mean(), atol=1e-6)                              # all elements
    assert np.allclose(run(tf.reduce_mean(p, axis=-1), f), x.mean(-1), atol=1e-6)
    assert np.allclose(run(tf.reduce_sum(p, axis=1), f), x.sum(1), atol=1e-5)
    assert np.isclose(run(tf.reduce_mean(tf.reduce_sum(p * p)), f), (x.astype(np.float64) ** 2).sum(), rtol=1e-6)
    assert np.isclose(run(tf.nn.l2_loss(p), f), (x.astype(np.float64) ** 2).sum() / 2, rtol=1e-6)  # sum(t^2) / 2
    logits = np.array([-80.0, -3.0, 0.0, 2.5, 90.0], dtype=np.float32)
    labels = np.array([1.0, 0.0, 1.0, 1.0, 0.0], dtype=np.float32)
    pl, fl = ph(logits)
    pz, fz = ph(labels)
    got = run(tf.nn.sigmoid_cross_entropy_with_logits(labels=pz, logits=pl), {**fl, **fz})
    x64, z64 = logits.astype(np.float64), labels.astype(np.float64)
    # documented definition z * -log(sigmoid(x)) + (1 - z) * -log(1 - sigmoid(x)), evaluated stably in float64
    want = z64 * np.logaddexp(0, -x64) + (1 - z64) * np.logaddexp(0, x64)
    assert np.all(np.isfinite(got)) and np.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert np.allclose(run(tf.sigmoid(pl), fl), 1 / (1 + np.exp(-x64)), atol=1e-6)
    assert np.array_equal(run(tf.nn.relu(p), f), np.maximum(x, 0))


def test_shape_ops():
    x = RNG.randn(2, 1, 3).astype(np.float32)
    p, f = ph(x)
    assert np.array_equal(run(tf.tile(p, [1, 4, 1]), f), np.tile(x, (1, 4, 1)))
    assert np.array_equal(run(tf.tile(p, multiples=[1, 2, 2]), f), np.tile(x, (1, 2, 2)))
    assert run(tf.expand_dims(p, axis=2), f).shape == (2, 1, 1, 3)
    assert run(tf.expand_dims(p, axis=-1), f).shape == (2, 1, 3, 1)
    assert run(tf.squeeze(p, axis=1), f).shape == (2, 3)
    y = RNG.randn(2, 5, 3).astype(np.float32)
    py, fy = ph(y)
    assert np.array_equal(run(tf.concat([p, py], axis=1), {**f, **fy}), np.concatenate([x, y], axis=1))
    assert np.array_equal(run(tf.concat([py, py], axis=-1), fy), np.concatenate([y, y], axis=-1))
    assert np.array_equal(run(tf.reshape(py, [-1, 3]), fy), y.reshape(-1, 3))                      # row-major, -1 inferred
    assert np.array_equal(run(tf.reshape(py, [2, 5, 1, 3]), fy), y.reshape(2, 5, 1, 3))
    assert tf.reshape(py, [-1, 3]).shape == (10, 3) and tf.concat([p, py], axis=1).get_shape()[1] == 6   # static shapes


def test_xavier_uniform_limits_follow_tf_contrib_fans():
    """tf.contrib.layers.xavier_initializer(uniform=True): U(-l, l), l = sqrt(6 / (fan_in + fan_out)); for rank > 2
    both fans are multiplied by the receptive field prod(shape[:-2]) -- the [n_rel, d, d] KGE table (model.py:84-86)."""
    import torch
    gen = torch.Generator().manual_seed(0)
    for shape, fans in (((400, 16), (400, 16)), ((7, 16, 16), (7 * 16, 7 * 16)), ((48, 1), (48, 1)), ((16,), (16, 16))):
        t = tf._xavier(list(shape), gen).numpy()
        lim = math.sqrt(6.0 / sum(fans))
        assert t.shape == shape and np.abs(t).max() <= lim
        if t.size >= 400:
            assert np.abs(t).max() > 0.9 * lim and abs(t.mean()) < 0.1 * lim
            assert abs(t.std() - lim / math.sqrt(3)) < 0.1 * lim                                   # uniform, not normal


def test_adam_matches_the_tf1_update_rule():
    """tf.train.AdamOptimizer docs: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
    var -= lr_t m / (sqrt(v) + eps)  (epsilon outside the bias correction -- 'epsilon hat')."""
    import torch
    tf.reset_default_graph()
    with tf.variable_scope("t"):
        w = tf.get_variable(name="w", shape=[5, 3])
    x = RNG.randn(4, 5).astype(np.float32)
    px, fx = ph(x)
    loss = tf.reduce_mean(tf.matmul(px, w) * tf.matmul(px, w))
    opt = tf.train.AdamOptimizer(0.05)
    step = opt.minimize(loss)
    sess = tf.Session()
    w0 = w.value.detach().numpy().astype(np.float64).copy()
    m = np.zeros_like(w0)
    v = np.zeros_like(w0)
    ref = w0.copy()
    for t in range(1, 4):
        y = x.astype(np.float64) @ ref
        g = 2.0 * x.astype(np.float64).T @ y / y.size
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        ref = ref - 0.05 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        sess.run([step, loss], fx)
        assert np.allclose(w.value.detach().numpy(), ref, rtol=2e-5, atol=2e-6), t
    tf.reset_default_graph()


def test_unfed_placeholder_raises():
    p = tf.placeholder("float32", shape=[None, 3], name="p")
    with pytest.raises(KeyError):
        run(p * 2.0)
    assert p.shape == (tf.DEFAULT_BATCH, 3)
