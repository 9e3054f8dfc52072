"""Independent naive NumPy forward of Gen_UNet2D / Dis_C2D_FCN1 (test infrastructure, NOT product code).

Purpose: the torch restatement in depgan_oracle.py restates the Keras layers' arithmetic (the reference cannot run here), so this
file re-derives the same forward passes from the Keras layer definitions with explicit shifted-sum
convolutions in NHWC and no torch.  It shares *no code* with depgan_oracle.py beyond the layer-name tables.
Only usable on small H, W (pure NumPy).  Citations as in depgan_oracle.py (TG:255-498).
"""
from __future__ import annotations

import numpy as np

from .depgan_oracle import CRITIC_CONVS, CRITIC_POOL_AFTER, FILM_HEADS, GEN_BLOCKS, GEN_DECONVS

EPS = 1e-3


def conv_same(x, k, b):
    """x (N,H,W,Ci), k (kh,kw,Ci,Co) Keras HWIO, 'same' zero padding, stride 1 (cross-correlation)."""
    n, h, w, ci = x.shape
    kh, kw, _, co = k.shape
    ph, pw = kh // 2, kw // 2
    xp = np.zeros((n, h + 2 * ph, w + 2 * pw, ci), dtype=x.dtype)
    xp[:, ph:ph + h, pw:pw + w, :] = x
    out = np.zeros((n, h, w, co), dtype=x.dtype)
    for a in range(kh):
        for c in range(kw):
            out += np.einsum("nhwi,io->nhwo", xp[:, a:a + h, c:c + w, :], k[a, c])
    return out + b


def bn(P, name, x):
    g, be = P[name + "/gamma"], P[name + "/beta"]
    mu, var = P[name + "/moving_mean"], P[name + "/moving_variance"]
    return g * (x - mu) / np.sqrt(var + EPS) + be


def pool2(x):
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def deconv2(x, k, b):
    """Conv2DTranspose k=2,s=2,'valid'; k (2,2,Co,Ci): out[n,2i+a,2j+c,o] = sum_i x[n,i,j,ci] k[a,c,o,ci] + b."""
    n, h, w, ci = x.shape
    co = k.shape[2]
    out = np.zeros((n, 2 * h, 2 * w, co), dtype=x.dtype)
    for a in range(2):
        for c in range(2):
            out[:, a::2, c::2, :] = np.einsum("nhwi,oi->nhwo", x, k[a, c])
    return out + b


def relu(x):
    return np.maximum(x, 0)


def film(P, z):
    h = relu(bn(P, "dense_bn_noise_1_add_f0", z @ P["dense_noise_1_add_f0/kernel"] + P["dense_noise_1_add_f0/bias"]))
    h = relu(bn(P, "dense_bn_noise_1_add_f1", h @ P["dense_noise_1_add_f1/kernel"] + P["dense_noise_1_add_f1/bias"]))
    h = h.reshape(h.shape[0], -1)
    out = {}
    for suf, _ in FILM_HEADS:
        mul = bn(P, "dense_bn_noise_2_mul" + suf, h @ P["dense_noise_2_mul%s/kernel" % suf] + P["dense_noise_2_mul%s/bias" % suf])
        add = bn(P, "dense_bn_noise_2_add" + suf, h @ P["dense_noise_2_add%s/kernel" % suf] + P["dense_noise_2_add%s/bias" % suf])
        out[suf] = (mul, add)
    return out


def gen_forward(P, x, z, head="tanh"):
    fp = film(P, z)
    skips = []
    for bi, (c_in, c_noise, c_out, suf, mult) in enumerate(GEN_BLOCKS):
        a = relu(bn(P, "bn_" + c_in, conv_same(x, P["conv2d_%s/kernel" % c_in], P["conv2d_%s/bias" % c_in])))
        y = bn(P, "bn_" + c_noise, conv_same(a, P["conv2d_%s/kernel" % c_noise], P["conv2d_%s/bias" % c_noise]))
        mul, add = fp[suf]
        r = relu(y * mul[:, None, None, :] + add[:, None, None, :]) + a
        o = relu(bn(P, "bn_" + c_out, conv_same(r, P["conv2d_%s/kernel" % c_out], P["conv2d_%s/bias" % c_out])))
        if bi < 3:
            skips.append(o)
            x = pool2(o)
        elif bi < 6:
            d = GEN_DECONVS[bi - 3]
            u = relu(bn(P, "bn_" + d, deconv2(o, P["deconv2d_%s/kernel" % d], P["deconv2d_%s/bias" % d])))
            x = np.concatenate([u, skips[5 - bi]], axis=-1)
        else:
            x = o
    seg = conv_same(x, P["gen_segmentation/kernel"], P["gen_segmentation/bias"])
    if head == "tanh":
        return np.tanh(seg)
    e = np.exp(seg - seg.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def critic_forward(P, x):
    for name, k, ci, co in CRITIC_CONVS:
        x = relu(conv_same(x, P[name + "/kernel"], P[name + "/bias"]))
        if name in CRITIC_POOL_AFTER:
            x = pool2(x)
    x = conv_same(x, P["dis_9/kernel"], P["dis_9/bias"])
    return x.reshape(x.shape[0], -1) @ P["dense_1/kernel"] + P["dense_1/bias"]
