"""The Keras stand-in (oracle/keras_shim.py) on which the reference's source is executed, checked on its own: every layer's
arithmetic against the independent shifted-sum NumPy layers of oracle/naive_numpy.py or a closed form, the merge layers'
rank broadcasting, K.gradients / K.function, and Adam.get_updates against the published update rule written out by hand.
(The golden vectors of tests/golden/reference_vectors.npz are only as good as this file's subject.)"""
import math

import numpy as np
import torch

from oracle import keras_shim as ks
from oracle import naive_numpy as NN

RNG = np.random.default_rng(5)


def _val(sym, feed):
    return ks.evaluate([sym], {k: torch.as_tensor(v, dtype=ks.DT) for k, v in feed.items()})[0].detach().numpy()


def _set(layer, **w):
    for k, v in w.items():
        with torch.no_grad():
            layer.weights[k].copy_(torch.as_tensor(v, dtype=ks.DT))


def test_conv2d_same_and_1x1_match_the_shifted_sum_convolution():
    for ks_, ci, co in ((3, 2, 5), (5, 1, 4), (1, 6, 3)):
        x = RNG.standard_normal((2, 7, 6, ci))
        k, b = RNG.standard_normal((ks_, ks_, ci, co)), RNG.standard_normal(co)
        inp = ks.Input((7, 6, ci))
        layer = ks.Conv2D(co, (ks_, ks_), padding="same")
        y = layer(inp)
        assert [tuple(t.shape) for t in layer.weights.values()] == [(ks_, ks_, ci, co), (co,)]   # HWIO kernel, bias
        _set(layer, kernel=k, bias=b)
        np.testing.assert_allclose(_val(y, {inp: x}), NN.conv_same(x, k, b), rtol=1e-12, atol=1e-12)


def test_conv2d_transpose_k2_s2_valid_scatters_each_pixel_to_its_2x2_block():
    ci, co = 3, 4
    x = RNG.standard_normal((2, 3, 5, ci))
    k, b = RNG.standard_normal((2, 2, co, ci)), RNG.standard_normal(co)   # Keras: (kh, kw, out, in)
    inp = ks.Input((3, 5, ci))
    layer = ks.Conv2DTranspose(co, (2, 2), strides=(2, 2), padding="valid")
    y = layer(inp)
    assert tuple(layer.weights["kernel"].shape) == (2, 2, co, ci) and y.shape == (None, 6, 10, co)
    _set(layer, kernel=k, bias=b)
    want = np.zeros((2, 6, 10, co))
    for a in range(2):
        for c in range(2):
            want[:, a::2, c::2, :] = np.einsum("nhwi,oi->nhwo", x, k[a, c]) + b
    np.testing.assert_allclose(_val(y, {inp: x}), want, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(want, NN.deconv2(x, k, b), rtol=1e-12, atol=1e-12)


def test_batchnorm_inference_formula_dense_on_last_axis_pool_flatten():
    x = RNG.standard_normal((2, 4, 4, 3))
    g, be, mu, var = RNG.uniform(0.5, 1.5, 3), RNG.standard_normal(3), RNG.standard_normal(3), RNG.uniform(0.5, 2.0, 3)
    inp = ks.Input((4, 4, 3))
    bn = ks.BatchNormalization()
    y = bn(inp)
    assert list(bn.weights) == ["gamma", "beta", "moving_mean", "moving_variance"] and bn.non_trainable == {"moving_mean", "moving_variance"}
    _set(bn, gamma=g, beta=be, moving_mean=mu, moving_variance=var)
    np.testing.assert_allclose(_val(y, {inp: x}), (x - mu) / np.sqrt(var + 1e-3) * g + be, rtol=1e-12)
    # Dense acts on the last axis of a rank-3 input (the noise path: (N, 32, 1) -> (N, 32, F))
    z = RNG.standard_normal((2, 5, 1))
    zi = ks.Input((5, 1))
    d = ks.Dense(4)
    yz = d(zi)
    k, b = RNG.standard_normal((1, 4)), RNG.standard_normal(4)
    _set(d, kernel=k, bias=b)
    np.testing.assert_allclose(_val(yz, {zi: z}), z @ k + b, rtol=1e-12)
    # Flatten is row-major over the non-batch axes; MaxPooling2D(2, 2) takes the 2x2 maximum; Dropout is the identity
    f = ks.Flatten()(yz)
    np.testing.assert_allclose(_val(f, {zi: z}), (z @ k + b).reshape(2, -1), rtol=1e-12)
    p = ks.MaxPooling2D(pool_size=(2, 2))(inp)
    np.testing.assert_allclose(_val(p, {inp: x}), NN.pool2(x), rtol=0)
    np.testing.assert_allclose(_val(ks.Dropout(0.25)(inp), {inp: x}), x, rtol=0)


def test_merge_layers_broadcast_a_lower_rank_input_over_the_spatial_axes():
    x = RNG.standard_normal((2, 3, 4, 5))
    v = RNG.standard_normal((2, 5))
    a, b = ks.Input((3, 4, 5)), ks.Input((5,))
    np.testing.assert_allclose(_val(ks.multiply([a, b]), {a: x, b: v}), x * v[:, None, None, :], rtol=1e-12)
    np.testing.assert_allclose(_val(ks.add([a, b]), {a: x, b: v}), x + v[:, None, None, :], rtol=1e-12)
    c = ks.Input((3, 4, 2))
    y = RNG.standard_normal((2, 3, 4, 2))
    np.testing.assert_allclose(_val(ks.concatenate([a, c], axis=-1), {a: x, c: y}), np.concatenate([x, y], -1), rtol=0)


def test_model_reapplication_shares_weights_and_gradients_are_of_the_sum():
    inp = ks.Input((4, 4, 1))
    conv = ks.Conv2D(2, (3, 3), padding="same")
    out = ks.Flatten()(ks.Activation("relu")(conv(inp)))
    m = ks.Model(inputs=inp, outputs=out)
    k, b = RNG.standard_normal((3, 3, 1, 2)), RNG.standard_normal(2)
    m.set_named_weights({conv.name + "/kernel": k, conv.name + "/bias": b})
    x = RNG.standard_normal((2, 4, 4, 1))
    other = ks.Input((4, 4, 1))
    y2 = m(other * 2.0)                                   # the model applied to a new tensor
    want = np.maximum(NN.conv_same(2.0 * x, k, b), 0).reshape(2, -1)
    np.testing.assert_allclose(_val(y2, {other: x}), want, rtol=1e-12)
    np.testing.assert_allclose(m.predict(2.0 * x), want, rtol=1e-12)
    # K.gradients(y, [x]) = d sum(y) / dx, checked by central differences
    grad = ks.K.gradients(m(other), [other])[0]
    got = _val(grad, {other: x})
    f = lambda t: float(np.maximum(NN.conv_same(t, k, b), 0).sum())
    for idx in [(0, 1, 2, 0), (1, 3, 0, 0), (0, 0, 0, 0)]:
        e = np.zeros_like(x); e[idx] = 1e-6
        assert abs((f(x + e) - f(x - e)) / 2e-6 - got[idx]) < 1e-6
    assert [tuple(t.shape) for t in m.trainable_weights] == [(3, 3, 1, 2), (2,)]


def test_function_outputs_use_the_pre_update_weights_and_adam_follows_the_published_rule():
    inp = ks.Input((3,))
    d = ks.Dense(1)
    y = d(inp)
    m = ks.Model(inputs=inp, outputs=y)
    w0, b0 = np.array([[0.5], [-1.0], [2.0]]), np.array([0.25])
    m.set_named_weights({d.name + "/kernel": w0, d.name + "/bias": b0})
    loss = ks.K.mean(ks.K.square(y))
    lr, b1, b2, eps = 1e-2, 0.0, 0.9, 1e-7
    upd = ks.Adam(lr=lr, beta_1=b1, beta_2=b2).get_updates(m.trainable_weights, [], loss)   # keras 2.0 / 2.1 argument order
    fn = ks.K.function([inp], [loss], upd)
    x = np.array([[1.0, 2.0, -1.0], [0.5, 0.0, 1.0]])
    w, b = w0.copy(), b0.copy()
    mw, vw, mb, vb = 0.0, 0.0, 0.0, 0.0
    for t in (1, 2, 3):
        pred = x @ w + b
        want_loss = float((pred ** 2).mean())
        got_loss = float(fn([x])[0])
        assert abs(got_loss - want_loss) < 1e-12          # evaluated BEFORE this call's update
        gw, gb = x.T @ (2 * pred) / 2.0, (2 * pred).mean(axis=0)
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        mw, vw = b1 * mw + (1 - b1) * gw, b2 * vw + (1 - b2) * gw * gw
        mb, vb = b1 * mb + (1 - b1) * gb, b2 * vb + (1 - b2) * gb * gb
        w, b = w - lr_t * mw / (np.sqrt(vw) + eps), b - lr_t * mb / (np.sqrt(vb) + eps)
        np.testing.assert_allclose(d.weights["kernel"].detach().numpy(), w, rtol=1e-12)
        np.testing.assert_allclose(d.weights["bias"].detach().numpy(), b, rtol=1e-12)
    # keras >= 2.1.3 argument order gives the same object
    assert ks.Adam().get_updates(loss, m.trainable_weights).params == m.trainable_weights
