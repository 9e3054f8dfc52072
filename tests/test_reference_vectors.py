"""The oracle against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (tests/golden/reference_vectors.npz,
made in the build container by tests/golden/make_reference_vectors.py: the reference's layer helpers, Dis_C2D_FCN1,
Gen_UNet2D, the WGAN-GP / generator-loss graph construction with its Adam optimizers and the testing scripts' per-subject
evaluation blocks, exec'ed unmodified on the Keras stand-in oracle/keras_shim.py).  This pins oracle/depgan_oracle.py --
the checker of every GPU parity test -- to what the reference's code does: manifests (names, shapes, creation order),
topology, loss composition, gradient penalty, which weights each step updates and how (Keras Adam, two iterations),
the 10-repeat float64 mean, DEM post-processing, labels, volumes and the 18-column evaluation row.
CPU only; reads nothing outside the repository."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from depgan_b200 import synth
from oracle import depgan_oracle as O

G = np.load(Path(__file__).parent / "golden" / "reference_vectors.npz", allow_pickle=False)
H, N, Z = 32, 3, 42


def _man(key):
    return [(l, w, tuple(s)) for l, w, s in json.loads(str(G[key]))]


def _weights(man, seed, rename=None):
    P = synth.init_weights(man, seed=seed, trained_like=True)
    if rename:
        P = {k.replace(rename[0], rename[1]): v for k, v in P.items()}
    return P


def _digest(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    return np.array([a.sum(), np.square(a).sum()] + list(a[:6]) + [0.0] * max(0, 6 - a.size))[:8]


def _close(a, b, rtol=1e-9, atol=1e-11):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag,nicg", [("gan_im", 1), ("gan_pf", 2)])
def test_manifests_are_the_ones_the_reference_code_creates(tag, nicg):
    """Names, shapes AND creation order of every weight tensor, as Keras would list them for the reference's models."""
    assert _man(tag + "/manifest_G") == [(l, w, tuple(s)) for l, w, s in O.gen_manifest(nicg, 1)]
    want_d = [(l, w, tuple(s)) for l, w, s in O.critic_manifest(H, H)]
    assert _man(tag + "/manifest_Dy2") == want_d
    # the second critic's unnamed Dense gets Keras' next automatic name (dense_2); everything else is identical
    assert [(l.replace("dense_2", "dense_1"), w, s) for l, w, s in _man(tag + "/manifest_Ddem")] == want_d


def test_native_manifest_matches_the_reference_code():
    """The C ABI's manifest (host-only calls) lists the same tensors in the same order."""
    from depgan_b200 import _lib, api
    for tag, nicg, nc in (("gan_im/manifest_G", 1, 1), ("gan_pf/manifest_G", 2, 1), ("topo_TU/manifest", 1, 4)):
        cfg = _lib.Cfg(256, 256, nicg, nc, 32, 1, _lib.PREC_FP32, 0)
        got = [(n.split("/")[0], n.split("/")[1], tuple(s)) for n, s, _, _ in api.manifest(_lib.MODEL_GEN, cfg)[0]]
        assert got == _man(tag)
    cfg = _lib.Cfg(H, H, 1, 1, 32, 1, _lib.PREC_FP32, 0)
    got = [(n.split("/")[0], n.split("/")[1], tuple(s)) for n, s, _, _ in api.manifest(_lib.MODEL_CRITIC, cfg)[0]]
    assert got == _man("gan_im/manifest_Dy2")


def _setup(tag, nicg, thr):
    sg, s1, s2 = [int(v) for v in G[tag + "/weight_seeds"]]
    PG = _weights(_man(tag + "/manifest_G"), sg)
    P1 = _weights(_man(tag + "/manifest_Dy2"), s1)
    P2 = _weights(_man(tag + "/manifest_Ddem"), s2, rename=("dense_2/", "dense_1/"))
    a, b, c, d = [int(v) for v in G[tag + "/input_seeds"]]
    x1, y2, _ = synth.make_im_pair(N, H, H, nicg=nicg, thr=thr, seed=a)
    return PG, P1, P2, x1, y2, synth.make_noise(N, seed=b), synth.make_eps(N, seed=c), synth.make_noise(N, seed=d)


@pytest.mark.parametrize("tag,nicg,thr", [("gan_im", 1, 0.178), ("gan_pf", 2, 0.5)])
def test_forward_passes_match_the_executed_reference_graph(tag, nicg, thr):
    PG, P1, P2, x1, y2, z, ep, _ = _setup(tag, nicg, thr)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    with torch.no_grad():
        _close(O.gen_forward(O.to_torch(PG), t(x1), t(z)).numpy(), G[tag + "/gen_out"])
        _close(O.critic_forward(O.to_torch(P1), t(y2)).numpy(), G[tag + "/critic_y2_out"])
        _close(O.critic_forward(O.to_torch(P2), t(y2 - x1[..., :1])).numpy(), G[tag + "/critic_dem_out"])


@pytest.mark.parametrize("tag,nicg,thr", [("gan_im", 1, 0.178), ("gan_pf", 2, 0.5)])
def test_step_functions_gradients_and_adam_updates_match_the_executed_reference_graph(tag, nicg, thr):
    """TG:513-598 executed: two rounds of netD_y2_train, netD_dem_train, netG_no_update, netG_train -- the returned
    losses, the gradient penalty, every gradient tensor of the first round and every weight after the second."""
    PG, P1, P2, x1, y2, z, ep, z2 = _setup(tag, nicg, thr)
    tr = O.OracleTrainer(PG, P1, P2, thr)
    tr.netD_y2_train([y2, x1, z, ep], update=False)
    _close(tr.last_gp, G[tag + "/gp_y2"])
    tr.netD_dem_train([y2, x1, z, ep], update=False)
    _close(tr.last_gp, G[tag + "/gp_dem"])

    def grads_in_order(man_key, rename=None):
        keys = [l + "/" + w for l, w, _ in _man(man_key) if O.is_trainable(w)]
        if rename:
            keys = [k.replace(*rename) for k in keys]
        return np.stack([_digest(tr.last_grads[k].numpy()) for k in keys])

    names = json.loads(str(G[tag + "/sequence"]))
    assert names == ["netD_y2_train", "netD_dem_train", "netG_no_update", "netG_train"] * 2
    i = 0
    for it in range(2):
        zz = z if it == 0 else z2
        _close(tr.netD_y2_train([y2, x1, zz, ep]), G[tag + "/seq%d" % i]); i += 1
        if it == 0:
            _close(grads_in_order(tag + "/manifest_Dy2"), G[tag + "/grad_digest_Dy2"], rtol=1e-7, atol=1e-12)
        _close(tr.netD_dem_train([y2, x1, zz, ep]), G[tag + "/seq%d" % i]); i += 1
        if it == 0:
            _close(grads_in_order(tag + "/manifest_Ddem", ("dense_2/", "dense_1/")), G[tag + "/grad_digest_Ddem"],
                   rtol=1e-7, atol=1e-12)
        _close(tr.netG_no_update([x1, y2, zz]), G[tag + "/seq%d" % i], rtol=1e-8); i += 1
        _close(tr.netG_train([x1, y2, zz]), G[tag + "/seq%d" % i], rtol=1e-8); i += 1
        if it == 0:
            _close(grads_in_order(tag + "/manifest_G"), G[tag + "/grad_digest_G"], rtol=1e-6, atol=1e-12)
    assert list(G[tag + "/optimizer_iterations"]) == [2, 2] and tr.optG.iterations == 2 and tr.optDy2.iterations == 2
    for P, key, rename in ((tr.PG, "G", None), (tr.PDy2, "Dy2", None), (tr.PDdem, "Ddem", ("dense_2/", "dense_1/"))):
        keys = [l + "/" + w for l, w, _ in _man(tag + "/manifest_" + key)]
        if rename:
            keys = [k.replace(*rename) for k in keys]
        got = np.stack([_digest(P[k].detach().numpy()) for k in keys])
        _close(got, G[tag + "/final_digest_" + key], rtol=1e-8, atol=1e-12)
    with torch.no_grad():
        out = O.gen_forward(tr.PG, torch.as_tensor(x1, dtype=torch.float64), torch.as_tensor(z, dtype=torch.float64))
    _close(out.numpy(), G[tag + "/gen_out_after"], rtol=1e-8, atol=1e-10)
    assert np.abs(G[tag + "/gen_out_after"] - G[tag + "/gen_out"]).max() > 1e-5   # the updates did move the generator


@pytest.mark.parametrize("tag,nicg,nc,head", [("TU", 1, 4, "softmax"), ("EU", 1, 4, "softmax"), ("EG", 1, 1, "tanh"),
                                              ("EG2", 2, 1, "tanh")])
def test_generator_as_each_other_script_defines_it(tag, nicg, nc, head):
    man = _man("topo_%s/manifest" % tag)
    assert man == [(l, w, tuple(s)) for l, w, s in O.gen_manifest(nicg, nc)]
    P = _weights(man, 201)
    x, _, _ = synth.make_im_pair(N, H, H, nicg=nicg, thr=0.5 if nicg == 2 else 0.178, seed=21)
    z = synth.make_noise(N, seed=22)
    with torch.no_grad():
        y = O.gen_forward(O.to_torch(P), torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(z, dtype=torch.float64), head)
    _close(y.numpy(), G["topo_%s/out" % tag])


def test_uresnet_compile_arguments():
    """TU:427: Adam(lr=1e-4) with Keras' default moments and categorical cross-entropy -- what fit() trains with."""
    c = json.loads(str(G["topo_TU/compile"]))
    assert c == {"loss": "categorical_crossentropy", "lr": 1e-4, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-7}


def _pix():
    return np.prod(np.array([0.9375, 0.9375, 4.0], dtype=np.float32))


def test_depgan_subject_evaluation_block_matches_the_executed_testing_script():
    """EG:615-807 executed on a synthetic 42-slice subject: 10-repeat float64 mean of masked float32 predictions, the
    clipped follow-up map, the strict / non-strict thresholds, labels, voxel count and the 18-column evaluation row."""
    thr = 0.178
    P = _weights(_man("EG_eval/manifest"), int(G["EG_eval/weight_seed"]))
    base, y2, mask = synth.make_im_pair(Z, H, H, nicg=1, thr=thr, seed=31)
    rs = np.random.RandomState(int(G["EG_eval/np_random_seed"]))
    preds = []
    Pt = O.to_torch(P, torch.float64)
    for _ in range(10):
        noise = rs.normal(size=(Z, 32, 1)).astype("float32")
        preds.append(np.squeeze(O.predict(Pt, base, noise, "tanh", dtype=torch.float64)))
    mean = O.inference_mean(preds, mask.reshape(Z, H, H))
    _close(mean, G["EG_eval/mean_map"], rtol=0, atol=1e-7)
    count, labels, fake2 = O.dem_postproc(base[..., 0], mean, mask.reshape(Z, H, H), thr)
    _close(fake2, G["EG_eval/fake2"], rtol=0, atol=1e-7)
    assert np.array_equal(labels, G["EG_eval/labels"]) and count == int(G["EG_eval/count_out"])
    w1 = (base[..., 0] >= thr).astype(np.float32)
    w2 = (y2[..., 0] >= thr).astype(np.float32)
    v1 = np.count_nonzero(np.multiply(G["EG_eval/mask1"], w1)) * _pix() / 1000
    v2 = np.count_nonzero(np.multiply(mask.reshape(Z, H, H), w2)) * _pix() / 1000
    vo = np.int64(count) * _pix() / 1000   # np.count_nonzero returns a NumPy integer: int64 x float32 -> float64
    row = O.evaluation_row(labels, G["EG_eval/code"], v1, v2, vo)
    _close(row, G["EG_eval/row"], rtol=1e-12)
    assert 0 < count < Z * H * H and len(np.unique(labels)) >= 3   # a non-degenerate case


def test_uresnet_subject_evaluation_block_matches_the_executed_testing_script():
    """EU:553-700 executed: softmax predictions x mask accumulated in float64, convert_from_1hot, count(label > 0), row."""
    man = [(l, w, tuple(s)) for l, w, s in O.gen_manifest(1, 4)]
    P = _weights(man, int(G["EU_eval/weight_seed"]))
    flair, _ = synth.make_flair(Z, H, H, seed=33)
    base, y2, mask = synth.make_im_pair(Z, H, H, nicg=1, thr=0.178, seed=31)
    rs = np.random.RandomState(int(G["EU_eval/np_random_seed"]))
    Pt = O.to_torch(P, torch.float64)
    preds = [O.predict(Pt, flair, rs.normal(size=(Z, 32, 1)).astype("float32"), "softmax", dtype=torch.float64)
             for _ in range(10)]
    mean = O.inference_mean(preds, mask.reshape(Z, H, H, 1))
    _close(mean, G["EU_eval/mean_map"], rtol=0, atol=1e-7)
    labels, count = O.uresnet_labels(mean)
    assert np.array_equal(labels, G["EU_eval/labels"]) and count == int(G["EU_eval/count_out"])
    w1 = (base[..., 0] >= 0.178).astype(np.float32)
    w2 = (y2[..., 0] >= 0.178).astype(np.float32)
    v1 = np.count_nonzero(np.multiply(G["EG_eval/mask1"], w1)) * _pix() / 1000
    v2 = np.count_nonzero(np.multiply(mask.reshape(Z, H, H), w2)) * _pix() / 1000
    row = O.evaluation_row(labels, G["EG_eval/code"], v1, v2, np.int64(count) * _pix() / 1000)
    _close(row, G["EU_eval/row"], rtol=1e-12)


def test_host_side_helpers_match_the_executed_reference_functions():
    """data_prep / data_prep_save / map_image_to_intensity_range (TG:105-149) executed on random inputs: the product's
    rewritten versions (depgan_b200/preproc.py) return the same arrays; convert_to_1hot (TU:209-223) is what fit() is fed."""
    from depgan_b200 import preproc
    assert np.array_equal(preproc.data_prep(G["host/vol"]), G["host/data_prep"])
    assert np.array_equal(preproc.data_prep_save(G["host/stack"]), G["host/data_prep_save"])
    _close(preproc.map_image_to_intensity_range(G["host/img64"], 0, 1, percentiles=0), G["host/map_p0"], rtol=1e-13)
    _close(preproc.map_image_to_intensity_range(G["host/img64"], -1, 1, percentiles=5), G["host/map_p5"], rtol=1e-13)
    _close(preproc.map_image_to_intensity_range(G["host/u8"], 0, 255, percentiles=2), G["host/map_u8"], rtol=1e-13)
    lab = G["host/labels"]
    onehot = np.eye(4, dtype=np.int16)[lab.astype(int)[..., 0]][..., None, :]   # (N,H,W,1,C) like the reference returns
    assert np.array_equal(onehot, G["host/onehot"]) and G["host/onehot"].shape == lab.shape + (4,)


@pytest.mark.parametrize("nicg,pm", [(1, True), (1, False), (2, True)])
def test_depgan_subject_preparation_matches_the_executed_testing_script(nicg, pm):
    """EG:533-611 executed on synthetic volumes (masks, stroke lesions, clamping, FLAIR min-max normalisation, channel
    order) against depgan_b200.preproc.prepare_subject_dem."""
    from depgan_b200 import preproc
    v = {k: G["prep/vol_" + k] for k in ("pm1", "im1", "flair", "icv1", "icv2", "sl1", "sl2")}
    x, m1, m2 = preproc.prepare_subject_dem(v["pm1"] if pm else v["im1"], v["icv1"], v["icv2"],
                                            flair_1tp=v["flair"] if nicg == 2 else None, sl_1tp=v["sl1"], sl_2tp=v["sl2"],
                                            nicg=nicg)
    tag = "prep/EG_nicg%d_%s" % (nicg, "pm" if pm else "im")
    want = G[tag + "/x"]
    assert x.shape == want.shape and x.dtype == np.float32
    np.testing.assert_allclose(x, want, rtol=0, atol=1e-7)     # the script keeps float64 where NumPy 2 promotes; values equal
    assert np.array_equal(m1, G[tag + "/mask1"]) and np.array_equal(m2, G[tag + "/mask2"])


def test_uresnet_subject_preparation_matches_the_executed_testing_script():
    """EU:506-538 executed: masked FLAIR, whole-volume z-score, nan_to_num; masks keep their channel axis."""
    from depgan_b200 import preproc
    v = {k: G["prep/vol_" + k] for k in ("flair", "icv1", "icv2", "sl1", "sl2")}
    x, m1, m2 = preproc.prepare_subject_uresnet(v["flair"], v["icv1"], v["icv2"], sl_1tp=v["sl1"], sl_2tp=v["sl2"])
    want = G["prep/EU/x"]
    assert x.shape == want.shape
    np.testing.assert_allclose(x, want, rtol=2e-6, atol=2e-6)
    assert np.array_equal(m1, G["prep/EU/mask1"]) and np.array_equal(m2, G["prep/EU/mask2"])
