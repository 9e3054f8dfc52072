"""The CUDA path (through the C ABI) directly against golden vectors produced by executing the reference's own source
(tests/golden/reference_vectors.npz, see tests/golden/make_reference_vectors.py and tests/test_reference_vectors.py): no
oracle in between.  fp32 path: forward outputs, the two-round sequence of the four step functions (TG:513-598) with
their Adam updates, every weight afterwards; the tensor-core formats on the forward outputs within their budgets."""
import json
from pathlib import Path

import numpy as np
import pytest

from depgan_b200 import synth

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).parent / "golden" / "reference_vectors.npz", allow_pickle=False)
H, N = 32, 3


def _man(key):
    return [(l, w, tuple(s)) for l, w, s in json.loads(str(G[key]))]


def _weights(man, seed, rename=None):
    P = synth.init_weights(man, seed=seed, trained_like=True)
    return {k.replace(*rename): v for k, v in P.items()} if rename else P


def _digest(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    return np.array([a.sum(), np.square(a).sum()] + list(a[:6]) + [0.0] * max(0, 6 - a.size))[:8]


def _inputs(tag, nicg, thr):
    a, b, c, d = [int(v) for v in G[tag + "/input_seeds"]]
    x1, y2, _ = synth.make_im_pair(N, H, H, nicg=nicg, thr=thr, seed=a)
    return x1, y2, synth.make_noise(N, seed=b), synth.make_eps(N, seed=c), synth.make_noise(N, seed=d)


@pytest.mark.parametrize("tag,nicg,thr", [("gan_im", 1, 0.178), ("gan_pf", 2, 0.5)])
def test_fp32_forward_and_training_sequence_match_the_executed_reference(tag, nicg, thr):
    from depgan_b200 import Dis_C2D_FCN1, Gen_UNet2D
    from depgan_b200.trainer import DepGanTrainer
    sg, s1, s2 = [int(v) for v in G[tag + "/weight_seeds"]]
    x1, y2, z, ep, z2 = _inputs(tag, nicg, thr)
    g = Gen_UNet2D((H, H, nicg), (32, 1), 32, 1, precision="fp32", max_batch=N, training=True)
    d1 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * N, training=True)
    d2 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * N, training=True)
    g.set_weights(_weights(_man(tag + "/manifest_G"), sg))
    d1.set_weights(_weights(_man(tag + "/manifest_Dy2"), s1))
    d2.set_weights(_weights(_man(tag + "/manifest_Ddem"), s2, rename=("dense_2/", "dense_1/")))
    assert np.abs(g.predict([x1, z]) - G[tag + "/gen_out"]).max() <= 1e-4
    assert np.allclose(d1.predict(y2), G[tag + "/critic_y2_out"], rtol=1e-4, atol=1e-4)
    assert np.allclose(d2.predict(y2 - x1[..., :1]), G[tag + "/critic_dem_out"], rtol=1e-4, atol=1e-4)
    tr = DepGanTrainer(g, d1, d2, thr)
    tr.netD_y2_train([y2, x1, z, ep], update=False)
    assert abs(tr.last_gp - float(G[tag + "/gp_y2"])) <= 1e-4 * max(1.0, abs(float(G[tag + "/gp_y2"])))
    tr.netD_dem_train([y2, x1, z, ep], update=False)
    assert abs(tr.last_gp - float(G[tag + "/gp_dem"])) <= 1e-4 * max(1.0, abs(float(G[tag + "/gp_dem"])))
    i = 0
    for it in range(2):
        zz = z if it == 0 else z2
        for name, args in (("netD_y2_train", [y2, x1, zz, ep]), ("netD_dem_train", [y2, x1, zz, ep]),
                           ("netG_no_update", [x1, y2, zz]), ("netG_train", [x1, y2, zz])):
            got = np.array(getattr(tr, name)(args), dtype=np.float64)
            want = G[tag + "/seq%d" % i]
            assert np.allclose(got, want, rtol=2e-4, atol=2e-5), (it, name, got, want)
            i += 1
    # every weight after two Adam steps per network (lr 1e-4: the updates are ~2e-4 per element, so the tolerance
    # below resolves them -- a missing or doubled update, a wrong beta or epsilon placement fails)
    for net, key, rename in ((g, "G", None), (d1, "Dy2", None), (d2, "Ddem", ("dense_2/", "dense_1/"))):
        W = net.get_weights()
        for (l, w, s), want in zip(_man(tag + "/manifest_" + key), G[tag + "/final_digest_" + key]):
            k = (l + "/" + w).replace(*rename) if rename else l + "/" + w
            got = _digest(W[k])
            n = int(np.prod(s))
            if key != "G" and k in ("dis_9/bias", "dense_1/bias"):
                # The critics' two output biases have a structurally zero WGAN-GP gradient (+1/N over the fake rows
                # cancels -1/N over the real rows; the penalty does not see an additive constant).  What reaches Adam
                # is rounding noise, which Adam (beta_1 = 0) normalises to steps of up to lr_t * sqrt(10) each -- in
                # float32 here as in the reference's float32 TF graph, but not in the float64 run that made the golden
                # vectors.  Bounded by the two steps taken, not compared digit for digit.
                assert np.abs(got[2:] - want[2:]).max() <= 2 * 3.2e-4, (key, k, got, want)
                continue
            # Adam (beta_1 = 0) turns a gradient element's RELATIVE error into an update error of lr_t * sqrt(10) * rel:
            # float32 gradients of small elements (a few per cent off, and not bit-reproducible: atomics) move a weight
            # by up to ~1e-5 differently from the float64 run; a missing / doubled update or a wrong beta is >= 3e-4
            assert np.allclose(got[2:], want[2:], rtol=1e-5, atol=2e-5), (key, k, got, want)       # first elements
            assert abs(got[0] - want[0]) <= 2e-5 * np.sqrt(n) + 2e-6 * n + 1e-5 * abs(want[0]), (key, k, got[0], want[0])
    moved = np.abs(g.predict([x1, z]) - G[tag + "/gen_out"]).max()
    assert np.abs(g.predict([x1, z]) - G[tag + "/gen_out_after"]).max() <= 2e-4 and moved > 1e-5


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("f16", 3e-3)])
def test_tensor_core_forward_against_the_executed_reference(precision, tol):
    from depgan_b200 import Gen_UNet2D
    x1, y2, z, ep, _ = _inputs("gan_im", 1, 0.178)
    g = Gen_UNet2D((H, H, 1), (32, 1), 32, 1, precision=precision, max_batch=N)
    g.set_weights(_weights(_man("gan_im/manifest_G"), int(G["gan_im/weight_seeds"][0])))
    assert np.abs(g.predict([x1, z]) - G["gan_im/gen_out"]).max() <= tol


@pytest.mark.parametrize("tag,nc", [("TU", 4), ("EG", 1)])
def test_other_scripts_generator_forward_fp32(tag, nc):
    from depgan_b200 import Gen_UNet2D
    x, _, _ = synth.make_im_pair(N, H, H, nicg=1, thr=0.178, seed=21)
    z = synth.make_noise(N, seed=22)
    g = Gen_UNet2D((H, H, 1), (32, 1), 32, nc, precision="fp32", max_batch=N)
    g.set_weights(_weights(_man("topo_%s/manifest" % tag), 201))
    assert np.abs(g.predict([x, z]) - G["topo_%s/out" % tag]).max() <= 1e-4
