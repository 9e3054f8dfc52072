"""Whole-network parity through the Keras-like surface / C ABI against the CPU oracle."""
import numpy as np
import pytest
import torch

from depgan_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nicg,nc_out,head", [(1, 1, "tanh"), (2, 1, "tanh"), (1, 4, "softmax")])
def test_generator_fp32_matches_oracle(nicg, nc_out, head):
    from depgan_b200 import Gen_UNet2D
    H = W = 64
    P = util.gen_weights(nicg, nc_out, seed=3)
    x, _, _ = synth.make_im_pair(3, H, W, nicg=nicg, seed=1)
    z = synth.make_noise(3, seed=2)
    g = Gen_UNet2D((H, W, nicg), (32, 1), 32, nc_out, precision="fp32", max_batch=4)
    g.set_weights(P)
    got = g.predict([x, z])
    want = util.oracle_gen(P, x, z, head)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() <= 1e-4, np.abs(got - want).max()  # FP32-accumulate variant (BASELINE.json)


@pytest.mark.parametrize("nicg,nc_out,head", [(1, 1, "tanh"), (2, 1, "tanh"), (1, 4, "softmax")])
def test_generator_bf16_tcgen05_matches_oracle(nicg, nc_out, head):
    from depgan_b200 import Gen_UNet2D
    H = W = 64
    P = util.gen_weights(nicg, nc_out, seed=3)
    x, _, _ = synth.make_im_pair(3, H, W, nicg=nicg, seed=1)
    z = synth.make_noise(3, seed=2)
    g = Gen_UNet2D((H, W, nicg), (32, 1), 32, nc_out, precision="bf16", max_batch=4)
    g.set_weights(P)
    got = g.predict([x, z])
    want = util.oracle_gen(P, x, z, head)
    assert np.abs(got - want).max() <= 1e-2, np.abs(got - want).max()  # DEM tolerance of BASELINE.json


@pytest.mark.parametrize("nicg,nc_out,head", [(1, 1, "tanh"), (2, 1, "tanh"), (1, 4, "softmax")])
def test_generator_f16_tcgen05_matches_oracle(nicg, nc_out, head):
    """precision='f16': the same tcgen05 kernels with IEEE-half activations / weights (inference handles).  64 x 64 also
    walks the CUDA-core fallbacks of the 8 x 8 bottleneck in that format."""
    from depgan_b200 import Gen_UNet2D
    H = W = 64
    P = util.gen_weights(nicg, nc_out, seed=3)
    x, _, _ = synth.make_im_pair(3, H, W, nicg=nicg, seed=1)
    z = synth.make_noise(3, seed=2)
    g = Gen_UNet2D((H, W, nicg), (32, 1), 32, nc_out, precision="f16", max_batch=4)
    g.set_weights(P)
    got = g.predict([x, z])
    want = util.oracle_gen(P, x, z, head)
    assert np.abs(got - want).max() <= 2.5e-3, np.abs(got - want).max()
    with pytest.raises(Exception):  # training handles keep the bf16 range
        Gen_UNet2D((H, W, nicg), (32, 1), 32, nc_out, precision="f16", max_batch=4, training=True)


def test_generator_intermediate_activations_fp32():
    from depgan_b200 import Gen_UNet2D
    from oracle import depgan_oracle as O
    H = W = 32
    P = util.gen_weights(1, 1, seed=5)
    x, _, _ = synth.make_im_pair(2, H, W, seed=1)
    z = synth.make_noise(2, seed=2)
    g = Gen_UNet2D((H, W, 1), precision="fp32", max_batch=2)
    g.set_weights(P)
    g.predict([x, z])
    Pt = O.to_torch(P, torch.float64)
    with torch.no_grad():
        film = O.film_params(Pt, torch.as_tensor(z, dtype=torch.float64))
        _, acts = O.gen_forward(Pt, torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(z, dtype=torch.float64),
                                return_acts=True)
    got_film = g.debug_activation("film", 2).reshape(2, -1)
    off = 0
    for bi, suf in enumerate(["_m1", "_m2", "_m3", "", "_p3", "_p2", "_p1"]):
        for k in range(2):
            want = film[suf][k].numpy()
            c = want.shape[1]
            assert np.abs(got_film[:, off:off + c] - want).max() <= 1e-4, (suf, k)
            off += c
    for name in ["gen_0", "gen_1", "gen_3", "gen_9", "de_gen_9", "gen_10", "gen_17"]:
        want = acts[name].permute(0, 2, 3, 1).numpy()
        got = g.debug_activation(name, 2).reshape(want.shape)
        assert np.abs(got - want).max() <= 2e-4, name


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_critic_matches_oracle(precision, tol):
    from depgan_b200 import Dis_C2D_FCN1
    H = W = 64
    P = util.critic_weights(H, W, seed=7)
    _, y2, _ = synth.make_im_pair(3, H, W, seed=4)
    d = Dis_C2D_FCN1((H, W, 1), precision=precision, max_batch=4)
    d.set_weights(P)
    got = d.predict(y2)
    want = util.oracle_critic(P, y2)
    scale = max(1.0, float(np.abs(want).max()))
    assert got.shape == (3, 1)
    assert np.abs(got - want).max() <= tol * scale, (got.ravel(), want.ravel())


def test_predict_batching_does_not_change_results():
    from depgan_b200 import Gen_UNet2D
    H = W = 32
    x, _, _ = synth.make_im_pair(5, H, W, seed=1)
    z = synth.make_noise(5, seed=2)
    g = Gen_UNet2D((H, W, 1), precision="bf16", max_batch=8)
    a = g.predict([x, z], batch_size=8)
    b = g.predict([x, z], batch_size=2)
    assert np.array_equal(a, b)


def test_full_size_generator_bf16_256():
    """BASELINE config 1 shape (256x256; batch reduced to 2 for oracle time) within the DEM tolerance.
    (Measured on B200, 4 slices: max-abs 5.4e-3 with these weights; 9.7e-3 with fresh-init weights whose
    un-normalised activations are the worst case for bf16 storage -- see DESIGN.md.)"""
    from depgan_b200 import Gen_UNet2D
    H = W = 256
    P = util.gen_weights(1, 1, seed=11, trained_like=True)
    x, _, _ = synth.make_im_pair(2, H, W, seed=1)
    z = synth.make_noise(2, seed=2)
    g = Gen_UNet2D((H, W, 1), precision="bf16", max_batch=2)
    g.set_weights(P)
    got = g.predict([x, z])
    want = util.oracle_gen(P, x, z, dtype=torch.float32)
    assert np.abs(got - want).max() <= 1e-2, np.abs(got - want).max()


# ---- BASELINE batch sizes through size-independent properties (the oracle is too slow at 64 x 256 x 256) ----
@pytest.mark.parametrize("nc_out,B", [(4, 64), (1, 16)])
def test_full_batch_slices_are_independent_and_order_free(nc_out, B):
    """configs[1] (DEP-UResNet, batch 64) and configs[0] (DEP-GAN generator, batch 16) at 256 x 256: slices never mix
    (BN runs on its moving statistics), so (i) a permuted batch gives the permuted result bit for bit although every
    slice lands on other CTAs / ring slots / accumulator stages, and (ii) one slice replicated over the batch gives
    identical outputs, equal to that slice computed alone."""
    from depgan_b200 import Gen_UNet2D
    H = W = 256
    P = util.gen_weights(1, nc_out, seed=5, trained_like=True)
    if nc_out == 4:
        x, _ = synth.make_flair(B, H, W, seed=3)
    else:
        x, _, _ = synth.make_im_pair(B, H, W, seed=3)
    z = synth.make_noise(B, seed=4)
    g = Gen_UNet2D((H, W, 1), (32, 1), 32, nc_out, precision="bf16", max_batch=B)
    g.set_weights(P)
    a = g.predict([x, z], batch_size=B)
    assert np.isfinite(a).all()
    perm = np.random.default_rng(0).permutation(B)
    b = g.predict([x[perm], z[perm]], batch_size=B)
    assert np.array_equal(b, a[perm])
    xr, zr = np.repeat(x[7:8], B, axis=0), np.repeat(z[7:8], B, axis=0)
    c = g.predict([xr, zr], batch_size=B)
    assert all(np.array_equal(c[i], c[0]) for i in range(1, B))
    assert np.array_equal(c[0], g.predict([x[7:8], z[7:8]], batch_size=B)[0])
    assert np.array_equal(c[0], a[7])
    if nc_out == 4:  # softmax rows sum to one
        assert np.abs(a.sum(-1) - 1.0).max() <= 1e-5


def test_full_batch_critic_rows_are_independent():
    """configs[2] critic batch (96 rows = real | fake | mixed of batch 32) at 256 x 256: no BatchNorm, strictly
    per-sample (TG:316-345)."""
    from depgan_b200 import Dis_C2D_FCN1
    H = W = 256
    rows = 96
    x, _, _ = synth.make_im_pair(rows, H, W, seed=9)
    d = Dis_C2D_FCN1((H, W, 1), precision="bf16", max_batch=rows, seed=3)
    a = d.predict(x[..., :1], batch_size=rows)
    assert a.shape == (rows, 1) and np.isfinite(a).all()
    perm = np.random.default_rng(1).permutation(rows)
    b = d.predict(x[perm][..., :1], batch_size=rows)
    assert np.array_equal(b, a[perm])
