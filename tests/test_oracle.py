"""Pins the CPU oracle (oracle/depgan_oracle.py).  The reference ships no tests or golden vectors and cannot
run here (its executed source pins the oracle in tests/test_reference_vectors.py); the restatement is additionally guarded by: an independent naive NumPy forward,
finite differences of the loss graphs in fp64, the parameter-count identities, hand-derived known answers for the
integer post-processing and Keras-Adam, and committed golden vectors (tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from depgan_b200 import synth
from oracle import depgan_oracle as O
from oracle import naive_numpy as NN

GOLD = Path(__file__).resolve().parent / "golden"


def _w(man, seed):
    return synth.init_weights(man, seed=seed, trained_like=True)


def test_param_counts():
    assert O.manifest_count(O.gen_manifest(1, 1)) == 2491969
    assert O.manifest_count(O.gen_manifest(2, 1)) == 2492257
    assert O.manifest_count(O.gen_manifest(1, 4)) == 2492068
    assert O.manifest_count(O.critic_manifest(256, 256)) == 1798002
    assert len(O.gen_manifest(1, 1)) == 242 and len(O.critic_manifest()) == 26


@pytest.mark.parametrize("nicg,nc_out,head", [(1, 1, "tanh"), (2, 1, "tanh"), (1, 4, "softmax")])
def test_generator_matches_independent_numpy_forward(nicg, nc_out, head):
    P = _w(O.gen_manifest(nicg, nc_out), 1)
    x, _, _ = synth.make_im_pair(2, 16, 16, nicg=nicg, seed=2)
    z = synth.make_noise(2, seed=3)
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    want = NN.gen_forward(P64, x.astype(np.float64), z.astype(np.float64), head)
    with torch.no_grad():
        got = O.gen_forward(O.to_torch(P), torch.as_tensor(x, dtype=torch.float64),
                            torch.as_tensor(z, dtype=torch.float64), head).numpy()
    assert np.abs(got - want).max() < 1e-10


def test_critic_matches_independent_numpy_forward():
    P = _w(O.critic_manifest(32, 32), 4)
    _, y2, _ = synth.make_im_pair(2, 32, 32, seed=5)
    want = NN.critic_forward({k: v.astype(np.float64) for k, v in P.items()}, y2.astype(np.float64))
    with torch.no_grad():
        got = O.critic_forward(O.to_torch(P), torch.as_tensor(y2, dtype=torch.float64)).numpy()
    assert np.abs(got - want).max() < 1e-10


def test_golden_vectors():
    g = np.load(GOLD / "oracle_vectors.npz")
    H = W = 32
    for tag, nicg, nc_out, head in [("gan_im", 1, 1, "tanh"), ("gan_pf", 2, 1, "tanh"), ("uresnet", 1, 4, "softmax")]:
        P = synth.init_weights(O.gen_manifest(nicg, nc_out), seed=21, trained_like=True)
        x, _, _ = synth.make_im_pair(2, H, W, nicg=nicg, seed=3)
        z = synth.make_noise(2, seed=4)
        with torch.no_grad():
            y = O.gen_forward(O.to_torch(P), torch.as_tensor(x, dtype=torch.float64),
                              torch.as_tensor(z, dtype=torch.float64), head).numpy()
        assert np.allclose(y, g[tag + "_out"], rtol=0, atol=1e-12), tag
    Pc = synth.init_weights(O.critic_manifest(H, W), seed=22, trained_like=True)
    x, y2, _ = synth.make_im_pair(2, H, W, seed=3)
    with torch.no_grad():
        c = O.critic_forward(O.to_torch(Pc), torch.as_tensor(y2, dtype=torch.float64)).numpy()
    assert np.allclose(c, g["critic_out"], rtol=0, atol=1e-12)
    PG = synth.init_weights(O.gen_manifest(1, 1), seed=21, trained_like=True)
    z, ep = synth.make_noise(2, seed=4), synth.make_eps(2, seed=5)
    tr = O.OracleTrainer(PG, Pc, synth.init_weights(O.critic_manifest(H, W), seed=23, trained_like=True), thr=0.178)
    assert np.allclose(tr.netD_y2_train([y2, x, z, ep], update=False) + [tr.last_gp], g["critic_y2_losses"], atol=1e-10)
    assert np.allclose(tr.netD_dem_train([y2, x, z, ep], update=False) + [tr.last_gp], g["critic_dem_losses"], atol=1e-10)
    assert np.allclose(tr.netG_no_update([x, y2, z]), g["gen_losses"], atol=1e-10)


def _fd_check(loss_fn, P, keys, n_probe=3, h=1e-6, seed=0):
    rng = np.random.default_rng(seed)
    loss = loss_fn()
    grads = torch.autograd.grad(loss, [P[k] for k in keys], allow_unused=True)
    for k, g in zip(keys, grads):
        flat = P[k].detach().view(-1)
        for idx in rng.choice(flat.numel(), size=min(n_probe, flat.numel()), replace=False):
            old = float(flat[idx])
            def at(v):  # set under no_grad, evaluate with autograd on (the GP needs autograd.grad inside)
                with torch.no_grad():
                    flat[idx] = v
                return float(loss_fn().detach())
            lp, lm = at(old + h), at(old - h)
            at(old)
            num = (lp - lm) / (2 * h)
            ana = 0.0 if g is None else float(g.reshape(-1)[idx])
            assert abs(num - ana) <= 1e-5 * max(1.0, abs(num), abs(ana)), (k, int(idx), num, ana)


@pytest.mark.parametrize("which", ["y2", "dem"])
def test_critic_loss_gradient_incl_gradient_penalty_vs_finite_differences(which):
    H = W = 16
    PG = O.to_torch(_w(O.gen_manifest(1, 1), 1))
    PD = O.to_torch(_w(O.critic_manifest(H, W), 2), requires_grad=True)
    x, y2, _ = synth.make_im_pair(2, H, W, seed=3)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    z, ep = t(synth.make_noise(2, seed=4)), t(synth.make_eps(2, seed=5))
    fn = lambda: O.critic_loss(PD, PG, t(y2), t(x), z, ep, which)[0]
    _fd_check(fn, PD, ["conv2d_dis_0a/kernel", "conv2d_dis_3/kernel", "conv2d_dis_8/bias", "dis_9/kernel",
                       "dense_1/kernel"])


def test_generator_loss_gradient_vs_finite_differences():
    H = W = 16
    PG = O.to_torch(_w(O.gen_manifest(1, 1), 1), requires_grad=True)
    PD1 = O.to_torch(_w(O.critic_manifest(H, W), 2))
    PD2 = O.to_torch(_w(O.critic_manifest(H, W), 3))
    x, y2, _ = synth.make_im_pair(2, H, W, seed=3)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    z = t(synth.make_noise(2, seed=4))
    fn = lambda: O.gen_loss(PG, PD1, PD2, t(x), t(y2), z, 0.178)[0]
    _fd_check(fn, PG, ["conv2d_gen_0/kernel", "bn_gen_noise_m2/gamma", "dense_noise_2_mul_p1/kernel",
                       "deconv2d_de_gen_11/kernel", "gen_segmentation/kernel", "dense_noise_1_add_f0/kernel"], h=1e-6)


def test_generator_loss_terms_known_answers():
    """M3 / M4 use hard >= thresholds on float32 values and are batch-global (TG:581-589)."""
    a = torch.tensor([[1.0, 0.0, 1.0, 1.0]], dtype=torch.float64)
    b = torch.tensor([[1.0, 1.0, 0.0, 1.0]], dtype=torch.float64)
    assert float(O.dice_coef(a, b)) == pytest.approx((2 * 2 + 1e-7) / (3 + 3 + 1e-7))


def test_keras_adam_known_answer():
    p = {"l/kernel": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    opt = O.KerasAdam(p, lr=1e-4, beta_1=0.0, beta_2=0.9, eps=1e-7)
    g = torch.tensor([0.5, -4.0], dtype=torch.float64)
    opt.step(p, {"l/kernel": g})
    lr_t = 1e-4 * np.sqrt(1 - 0.9) / (1 - 0.0)
    want = np.array([1.0, -2.0]) - lr_t * g.numpy() / (np.sqrt(0.1 * g.numpy() ** 2) + 1e-7)
    assert np.allclose(p["l/kernel"].numpy(), want, rtol=0, atol=1e-15)
    assert opt.iterations == 1


def test_dem_postproc_known_answers():
    thr = 0.178
    t32 = float(np.float32(thr))  # nearest float32 lies ABOVE 0.178
    assert t32 > thr
    base = np.array([[[0.1, 0.3, 0.3, 0.1, t32, 0.9]]], np.float32)
    dem = np.array([[[0.2, -0.2, 0.1, 0.0, -1e-9, 5.0]]], np.float64)
    mask = np.array([[[1, 1, 1, 1, 1, 0]]], np.float32)
    count, labels, fake2 = O.dem_postproc(base, dem, mask, thr)
    # fake2 = [0.3, 0.1, 0.4, 0.1, ~0.178, 1.0(clipped)] -> grow, shrink, stay, none, stay, stay
    assert labels.ravel().tolist() == [2.0, 1.0, 3.0, 0.0, 3.0, 3.0]
    assert fake2.ravel()[5] == 1.0
    assert count == 3  # voxels 0, 2, 4 are > thr inside the mask; voxel 5 is masked out
    # strict '>' for the volume, '>=' for labels: a voxel exactly at thr counts as 'stay'/'grow' but not as volume
    c2, l2, _ = O.dem_postproc(np.array([[[0.0]]], np.float32), np.array([[[thr]]]), np.ones((1, 1, 1), np.float32), thr)
    assert c2 == 0 and l2.ravel().tolist() == [2.0]


def test_inference_mean_is_float64_of_float32_products():
    p = [np.array([[[0.1]]], np.float32), np.array([[[0.2]]], np.float32)]
    m = np.array([[[1.0]]], np.float32)
    got = O.inference_mean(p, m)
    assert got.dtype == np.float64
    assert got.ravel()[0] == (np.float64(np.float32(0.1)) + np.float64(np.float32(0.2))) / 2.0


def test_uresnet_labels_first_max():
    prob = np.array([[[[0.25, 0.25, 0.25, 0.25], [0.1, 0.4, 0.4, 0.1]]]])
    lab, cnt = O.uresnet_labels(prob)
    assert lab.dtype == np.uint8 and lab.ravel().tolist() == [0, 1] and cnt == 1
