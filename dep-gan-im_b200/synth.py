"""Seeded synthetic weights and inputs for the DEP-GAN hot path (NumPy only).

The reference's trained ``.h5`` weights and NIfTI cohort are not available, so
benchmarks and tests run on synthetic tensors with the value ranges the reference's preprocessing produces
(EG:533-613: IM/PM masked, clamped to [0,1]; TU:493-512: z-scored FLAIR) and on weights drawn with the
initialisers the reference names (TG:255-312: Conv2D default glorot_uniform, Dense/dis_9 he_normal).
"""
from __future__ import annotations

import numpy as np


def init_weights(manifest, seed=0, trained_like=False):
    """manifest: iterable of (layer, weight, shape).  Returns dict 'layer/weight' -> float32 array.

    trained_like=False reproduces a fresh Keras model (bias 0, BN gamma 1 / beta 0 / moving stats 0,1).
    trained_like=True perturbs every vector so that no term of the arithmetic is trivially 0 or 1
    (used by parity tests so BN/bias handling is actually exercised).
    """
    rng = np.random.default_rng(seed)
    out = {}
    for layer, weight, shape in manifest:
        key = "%s/%s" % (layer, weight)
        shape = tuple(int(s) for s in shape)
        if weight == "kernel":
            if len(shape) == 4:
                rf = shape[0] * shape[1]
                fan_in, fan_out = rf * shape[2], rf * shape[3]
            else:
                fan_in, fan_out = shape[0], shape[1]
            he = layer.startswith("dense_") or layer == "dis_9"
            if he:  # he_normal: truncated normal, stddev sqrt(2/fan_in)
                std = np.sqrt(2.0 / fan_in)
                v = rng.standard_normal(shape)
                bad = np.abs(v) > 2
                while bad.any():
                    v[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(v) > 2
                v = v * std
            else:   # glorot_uniform
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                v = rng.uniform(-lim, lim, shape)
        elif weight in ("gamma", "moving_variance"):
            v = np.ones(shape)
            if trained_like:
                v = v + (0.2 * rng.uniform(-1, 1, shape) if weight == "gamma" else 0.5 * rng.uniform(0, 1, shape))
        else:  # bias, beta, moving_mean
            v = np.zeros(shape)
            if trained_like:
                v = 0.1 * rng.standard_normal(shape)
        out[key] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def _smooth_field(rng, n, h, w, sigma):
    """Low-pass random field in [0,1] per slice (FFT Gaussian filter)."""
    f = rng.standard_normal((n, h, w))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    g = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fy * fy + fx * fx))
    s = np.fft.irfft2(np.fft.rfft2(f) * g, s=(h, w))
    s = (s - s.mean(axis=(1, 2), keepdims=True)) / (s.std(axis=(1, 2), keepdims=True) + 1e-12)
    return s


def brain_mask(n, h, w, rng=None):
    """Axial brain-like ellipse mask (float32 0/1), slightly varying per slice."""
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((n, h, w), np.float32)
    for i in range(n):
        ry = 0.40 * h * (1.0 - 0.25 * abs((i + 0.5) / n - 0.5))
        rx = 0.33 * w * (1.0 - 0.25 * abs((i + 0.5) / n - 0.5))
        out[i] = (((yy - h / 2) / ry) ** 2 + ((xx - w / 2) / rx) ** 2 <= 1.0)
    return out


def make_im_pair(n, h=256, w=256, nicg=1, thr=0.178, seed=0):
    """Synthetic (real_1tp (n,h,w,nicg), real_2tp (n,h,w,1), mask (n,h,w)) float32, SURVEY 8d configs 1/3/4.

    Channel 0 = IM/PM in [0,1] with ~1-3 % of in-brain voxels >= thr; channel 1 (nicg=2) = FLAIR in [0,1].
    real_2tp = clip(base + smooth perturbation, 0, 1) * mask.
    """
    rng = np.random.default_rng(seed)
    mask = brain_mask(n, h, w)
    s = _smooth_field(rng, n, h, w, sigma=max(h, w) / 40.0)
    # shift so that roughly 2 % of the standard-normal field exceeds thr: P(s*a + b >= thr) ~ 0.02 at z = 2.05
    a = 0.12
    base = np.clip(s * a + (thr - 2.05 * a), 0.0, 1.0) * mask
    d = _smooth_field(rng, n, h, w, sigma=max(h, w) / 25.0) * 0.05
    y2 = np.clip(base + d * (base > 0.02), 0.0, 1.0) * mask
    x = base[..., None]
    if nicg == 2:
        fl = _smooth_field(rng, n, h, w, sigma=max(h, w) / 60.0)
        fl = (fl - fl.min()) / (fl.max() - fl.min() + 1e-12) * mask
        x = np.concatenate([x, fl[..., None]], axis=-1)
    return x.astype(np.float32), y2[..., None].astype(np.float32), mask.astype(np.float32)


def make_flair(n, h=256, w=256, seed=0):
    """Synthetic z-scored FLAIR volume (n,h,w,1) float32 + mask, SURVEY 8d config 2 (TU:493-512)."""
    rng = np.random.default_rng(seed)
    mask = brain_mask(n, h, w)
    v = (_smooth_field(rng, n, h, w, sigma=max(h, w) / 60.0) * 0.15 + 0.5) * mask
    v = (v - v.mean()) / (v.std() + 1e-12)
    return v[..., None].astype(np.float32), mask.astype(np.float32)


def make_noise(n, seed=0, noise_len=32):
    return np.random.default_rng(seed).standard_normal((n, noise_len, 1)).astype(np.float32)


def make_eps(n, seed=0):
    return np.random.default_rng(seed).uniform(size=(n, 1, 1, 1)).astype(np.float32)
