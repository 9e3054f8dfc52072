"""Bit-exact DEM post-processing (EG:616-628, 673-741; EU:570, 597-600) through the C ABI vs the NumPy oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import depgan_oracle as O

pytestmark = pytest.mark.gpu


def _run_gan(base, preds, mask, thr, nicg=1):
    from depgan_b200 import postproc
    return postproc.dem_pipeline(base, preds, mask, thr, nicg=nicg)


@pytest.mark.parametrize("thr", [0.178, 0.5])
def test_dem_postproc_bit_exact_random(thr):
    rng = np.random.default_rng(0)
    Z, H, W = 6, 64, 64
    base = rng.uniform(0, 1, (Z, H, W)).astype(np.float32)
    mask = (rng.uniform(size=(Z, H, W)) > 0.3).astype(np.float32)
    preds = [rng.uniform(-1, 1, (Z, H, W)).astype(np.float32) for _ in range(10)]
    dem, fake2, labels, count = _run_gan(base[..., None], preds, mask, thr)
    dem_o = O.inference_mean(preds, mask)
    count_o, labels_o, fake2_o = O.dem_postproc(base, dem_o, mask, thr)
    assert np.array_equal(dem, dem_o)            # float64 accumulation order is fixed -> bit-exact
    assert np.array_equal(fake2, fake2_o)
    assert np.array_equal(labels, labels_o.astype(np.uint8))
    assert count == count_o


def test_dem_postproc_adversarial_values():
    """Values exactly at T, at +-1 after clipping, mask = 0, float32(T) vs float64(T) boundary (SURVEY a-10)."""
    thr = 0.178
    t32 = np.float32(thr)
    base = np.array([thr, t32, np.nextafter(t32, np.float32(0)), 0.0, 1.0, 0.9, 0.1, t32], np.float32)
    d = np.array([0.0, 0.0, 0.0, thr, 0.5, -2.5, float(np.float64(thr) - np.float64(np.float32(0.1))), -1e-12])
    base = base.reshape(1, 1, -1)
    preds = [np.broadcast_to(d.astype(np.float32).reshape(1, 1, -1), base.shape).copy()]
    mask = np.ones_like(base)
    mask[0, 0, 4] = 0.0
    dem, fake2, labels, count = _run_gan(base[..., None], preds, mask, thr)
    dem_o = O.inference_mean(preds, mask)
    count_o, labels_o, fake2_o = O.dem_postproc(base, dem_o, mask, thr)
    assert np.array_equal(fake2, fake2_o) and np.array_equal(labels, labels_o.astype(np.uint8)) and count == count_o


def test_uresnet_labels_first_max_wins():
    from depgan_b200 import postproc
    rng = np.random.default_rng(1)
    Z, H, W = 3, 32, 32
    preds = [rng.dirichlet(np.ones(4), (Z, H, W)).astype(np.float32) for _ in range(10)]
    preds[0][0, 0, 0] = [0.25, 0.25, 0.25, 0.25]
    for p in preds[1:]:
        p[0, 0, 0] = [0.25, 0.25, 0.25, 0.25]
        p[0, 0, 1] = [0.1, 0.4, 0.4, 0.1]
    preds[0][0, 0, 1] = [0.1, 0.4, 0.4, 0.1]
    mask = (rng.uniform(size=(Z, H, W)) > 0.2).astype(np.float32)
    mean, labels, count = postproc.uresnet_pipeline(preds, mask)
    mean_o = O.inference_mean(preds, mask[..., None])
    lab_o, cnt_o = O.uresnet_labels(mean_o)
    assert np.array_equal(mean, mean_o) and np.array_equal(labels, lab_o) and count == cnt_o


def test_postproc_empty():
    from depgan_b200 import postproc
    base = np.zeros((0, 16, 16, 1), np.float32)
    dem, fake2, labels, count = postproc.dem_pipeline(base, [np.zeros((0, 16, 16), np.float32)],
                                                      np.zeros((0, 16, 16), np.float32), 0.5)
    assert dem.shape == (0, 16, 16) and count == 0
