"""Golden vectors from the REFERENCE'S OWN SOURCE, executed (run from the repo root in the build container, where
/root/reference exists: `python tests/golden/make_reference_vectors.py`; writes tests/golden/reference_vectors.npz).

The reference is four Python-2 / Keras-2 / TF-1 scripts; Keras and TensorFlow do not exist here.  This script reads the
scripts' text from /root/reference, cuts out -- by anchor lines, nothing is copied into the repository -- the pieces that
define the hot path and exec's them UNMODIFIED against oracle/keras_shim.py (a torch-float64 stand-in for the few Keras
layers / backend calls they use):

  TG (DEP-GAN training):  dice_coef / dice_coef_loss, the layer helpers, Dis_C2D_FCN1, Gen_UNet2D (TG:153-162, 255-498)
                          and the graph construction of the four step functions netD_y2_train, netD_dem_train,
                          netG_no_update, netG_train with their Adam optimizers (TG:513-598), at imageSize 32;
  TU / EU / EG:           each script's own helpers + Gen_UNet2D (softmax head with compile() in TU; the testing copies);
  EG (DEP-GAN testing):   the per-subject block from the 10-repeat prediction loop to the evaluation row
                          (EG:615-807), fed with a synthetic 42-slice subject and the shim-built netG;
  EU (UResNet testing):   convert_from_1hot and the same block of that script (EU:553-~690).

Python-2 print STATEMENTS inside the cut blocks are blanked (they do not parse under Python 3); nothing else is touched.
What is recorded: the layer / weight manifests the code creates (names, shapes, creation order), forward outputs, the
outputs of a short sequence of the step functions, gradient and post-update weight digests, the gradient penalty, and
the post-processing results.  tests/test_reference_vectors.py checks the oracle (and the native manifest) against them.
"""
import json
import re
import sys
import textwrap
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from depgan_b200 import synth  # noqa: E402
from oracle import keras_shim as ks  # noqa: E402

REF = Path("/root/reference")
TG = REF / "DEP-GAN_PROB_IM_twoCritics_training_4fold.py"
EG = REF / "DEP-GAN_testing_4fold.py"
TU = REF / "DEP-UResNet-wNoises-training-4fold.py"
EU = REF / "DEP-UResNet_testing_4fold.py"
OUT = Path(__file__).resolve().parent / "reference_vectors.npz"
H = 32          # imageSize of the fixtures (the critics need H % 16 == 0)
N = 3           # batch
Z_SLICES = 42   # the testing scripts draw noise for exactly 42 slices (EG:620, EU:557)


def cut(path, start_pat, end_pat, include_end=False):
    """Lines of `path` from the first line matching start_pat up to the next line matching end_pat."""
    lines = path.read_text(encoding="utf-8", errors="replace").splitlines()
    i0 = next(i for i, l in enumerate(lines) if re.search(start_pat, l))
    i1 = next(i for i in range(i0 + 1, len(lines)) if re.search(end_pat, lines[i]))
    block = lines[i0:i1 + (1 if include_end else 0)]
    # Python-2 print statements -> pass (same indentation); print(...) calls are left alone
    block = [re.sub(r"^(\s*)print\b(?!\s*\().*$", r"\1pass", l) for l in block]
    return textwrap.dedent("\n".join(block)) + "\n", (i0 + 1, i1)


def run(src, ns, what):
    ns.setdefault("print", lambda *a, **k: None)   # the blocks' print(...) calls: silenced
    exec(compile(src, "<reference:%s>" % what, "exec"), ns)


def manifest_of(model):
    return [(l, w, tuple(int(s) for s in t.shape)) for l, w, t in model.named_weights()]


def digest(t):
    a = np.asarray(t.detach().numpy() if hasattr(t, "detach") else t, dtype=np.float64).ravel()
    return np.array([a.sum(), np.square(a).sum()] + list(a[:6]) + [0.0] * max(0, 6 - a.size))[:8]


def build_networks(path, tag, ns):
    """exec one script's layer helpers + network definitions."""
    if path in (TG, EG):
        src, span = cut(path, r"^def dice_coef\(", r"^(###|'''|# Class|class )")
        run(src, ns, tag + ":dice")
    src, span = cut(path, r"^def dense_bn\(", r"^''' SECTION 5")
    run(src, ns, tag + ":networks")
    return span


def gan_training_graph(nicg, thr, out, tag):
    ns = ks.namespace()
    ns.update(imageSize=H, noiseSize=32, nicg=nicg, first_fm_G=32, delta=10, lrD=1e-4, lrG=1e-4, IM_TRSH=thr)
    build_networks(TG, "TG", ns)
    src, span = cut(TG, r"^\s+netD_y2 = Dis_C2D_FCN1\(", r"---- LOAD TRAINING DATA")
    run(src, ns, "TG:graph")
    out[tag + "/graph_lines"] = np.array(span)
    netG, Dy2, Ddem = ns["netG"], ns["netD_y2"], ns["netD_dem"]
    mans = {"G": manifest_of(netG), "Dy2": manifest_of(Dy2), "Ddem": manifest_of(Ddem)}
    for k, m in mans.items():
        out[tag + "/manifest_" + k] = np.array(json.dumps(m))
    seeds = {"G": 101, "Dy2": 102, "Ddem": 103}
    for k, model in (("G", netG), ("Dy2", Dy2), ("Ddem", Ddem)):
        model.set_named_weights(synth.init_weights(mans[k], seed=seeds[k], trained_like=True))
    out[tag + "/weight_seeds"] = np.array([seeds["G"], seeds["Dy2"], seeds["Ddem"]])
    out[tag + "/trainable_G"] = np.array(json.dumps([n for n, w, t in netG.named_weights() if t.requires_grad and False] or
                                                    [l + "/" + w for l, w, t in netG.named_weights() if t.requires_grad]))
    x1, y2, _ = synth.make_im_pair(N, H, H, nicg=nicg, thr=thr, seed=11)
    z, ep = synth.make_noise(N, seed=12), synth.make_eps(N, seed=13)
    z2 = synth.make_noise(N, seed=14)
    out[tag + "/input_seeds"] = np.array([11, 12, 13, 14])
    # forward passes (model.predict) before any update
    out[tag + "/gen_out"] = netG.predict([x1, z])
    out[tag + "/critic_y2_out"] = Dy2.predict(y2)
    out[tag + "/critic_dem_out"] = Ddem.predict(y2 - x1[..., :1])
    # gradient penalties of the two critic graphs, from the reference's own tensors
    K = ks.K
    gp = K.function([ns["netD_real_input"], ns["netG_real_input"], ns["noiseZ"], ns["ep_input"]], [ns["grad_penalty"]])
    gpd = K.function([ns["netD_real_input"], ns["netG_real_input"], ns["noiseZ"], ns["ep_input_dem"]],
                     [ns["grad_penalty_dem"]])
    out[tag + "/gp_y2"] = gp([y2, x1, z, ep])[0]
    out[tag + "/gp_dem"] = gpd([y2, x1, z, ep])[0]
    # the step functions, in the order of one (shortened) generator iteration, twice (Adam t = 1, 2)
    seq = []
    for it in range(2):
        zz = z if it == 0 else z2
        f = ns["netD_y2_train"]
        seq.append(("netD_y2_train", f([y2, x1, zz, ep])))
        if it == 0:
            out[tag + "/grad_digest_Dy2"] = np.stack([digest(g) for g in f.last_grads])
        f = ns["netD_dem_train"]
        seq.append(("netD_dem_train", f([y2, x1, zz, ep])))
        if it == 0:
            out[tag + "/grad_digest_Ddem"] = np.stack([digest(g) for g in f.last_grads])
        seq.append(("netG_no_update", ns["netG_no_update"]([x1, y2, zz])))
        f = ns["netG_train"]
        seq.append(("netG_train", f([x1, y2, zz])))
        if it == 0:
            out[tag + "/grad_digest_G"] = np.stack([digest(g) for g in f.last_grads])
    out[tag + "/sequence"] = np.array(json.dumps([s for s, _ in seq]))
    for i, (s, v) in enumerate(seq):
        out[tag + "/seq%d" % i] = np.array([float(np.asarray(a)) for a in v])
    for k, model in (("G", netG), ("Dy2", Dy2), ("Ddem", Ddem)):
        out[tag + "/final_digest_" + k] = np.stack([digest(t) for _, _, t in model.named_weights()])
    out[tag + "/optimizer_iterations"] = np.array([ns["training_updates"].opt.iterations,
                                                   ns["training_updates_dem"].opt.iterations])
    out[tag + "/gen_out_after"] = netG.predict([x1, z])


def generator_topologies(out):
    """Gen_UNet2D as each of the other three scripts defines it (softmax head in TU / EU, tanh in EG)."""
    for path, tag, nicg, nc in ((TU, "TU", 1, 4), (EU, "EU", 1, 4), (EG, "EG", 1, 1), (EG, "EG2", 2, 1)):
        ns = ks.namespace()
        ks.Model.compile = lambda self, **kw: setattr(self, "compiled", kw)   # TU:427 compiles inside Gen_UNet2D
        build_networks(path, tag, ns)
        g = ns["Gen_UNet2D"]((H, H, nicg), (32, 1), 32, nc)
        man = manifest_of(g)
        g.set_named_weights(synth.init_weights(man, seed=201, trained_like=True))
        x, _, _ = synth.make_im_pair(N, H, H, nicg=nicg, thr=0.5 if nicg == 2 else 0.178, seed=21)
        z = synth.make_noise(N, seed=22)
        out["topo_%s/manifest" % tag] = np.array(json.dumps(man))
        out["topo_%s/out" % tag] = g.predict([x, z])
        if tag == "TU":
            c = getattr(g, "compiled", {})
            opt = c.get("optimizer")
            out["topo_TU/compile"] = np.array(json.dumps({"loss": c.get("loss"), "lr": opt.lr, "beta_1": opt.beta_1,
                                                          "beta_2": opt.beta_2, "epsilon": opt.epsilon}))


class _Hdr:
    pixdim = np.array([0.9375, 0.9375, 4.0], dtype=np.float32)


def testing_blocks(out):
    # ---- EG: DEP-GAN per-subject evaluation ----
    ns = ks.namespace()
    build_networks(EG, "EG", ns)
    thr = 0.178
    netG = ns["Gen_UNet2D"]((H, H, 1), (32, 1), 32, 1)
    man = manifest_of(netG)
    netG.set_named_weights(synth.init_weights(man, seed=301, trained_like=True))
    base, y2, mask = synth.make_im_pair(Z_SLICES, H, H, nicg=1, thr=thr, seed=31)
    rng = np.random.default_rng(32)
    code = rng.integers(0, 4, size=(Z_SLICES, H, H, 1)).astype(np.float32)
    m1 = (rng.random((Z_SLICES, H, H)) > 0.2).astype(np.float32)
    w1 = (base[..., 0] >= thr).astype(np.float32)
    w2 = (y2[..., 0] >= thr).astype(np.float32)
    src, span = cut(EG, r"# Produce 10 results by using 10 different sets of noise", r"vol_dsc_best_all\.append")
    env = dict(np=np, netG=netG, noiseSize=32, nicg=1, PM=False, TRSH_VAL=thr, brain_prob__1tp=base.copy(),
               brain_prob__2tp=y2.copy(), icv_and_sl_mask_2tp=mask.reshape(Z_SLICES, H, H).copy(), icv_and_sl_mask_1tp=m1,
               brain_wmh_1tp=w1, brain_wmh_2tp=w2, brain_code_2tp=code, loaded_data_f_1tp=_Hdr())
    # model.predict returns the graph's float32 output; the shim evaluates in float64 and rounds once
    pred = netG.predict
    netG.predict = lambda xs, **kw: pred(xs, **kw).astype(np.float32)
    np.random.seed(4242)
    run(src, env, "EG:evaluation")
    out["EG_eval/lines"] = np.array(span)
    out["EG_eval/np_random_seed"] = np.array(4242)
    out["EG_eval/weight_seed"] = np.array(301)
    out["EG_eval/manifest"] = np.array(json.dumps(man))
    out["EG_eval/code"] = code
    out["EG_eval/mask1"] = m1
    out["EG_eval/mean_map"] = env["output_img_pred"]
    out["EG_eval/fake2"] = env["brain_prob__2tp_fake"]
    out["EG_eval/labels"] = env["wmh_change_mask_fake"]
    out["EG_eval/count_out"] = np.array(np.count_nonzero(env["wmh_from_out_2tp"]))
    out["EG_eval/row"] = np.array([float(v) for v in env["vol_dsc"]])
    out["EG_eval/vols_iam"] = np.array([env["vol_1tp__ml_iam"], env["vol_2tp__ml_iam"]])
    # ---- EU: DEP-UResNet per-subject evaluation ----
    ns = ks.namespace()
    build_networks(EU, "EU", ns)
    net = ns["Gen_UNet2D"]((H, H, 1), (32, 1), 32, 4)
    man = manifest_of(net)
    net.set_named_weights(synth.init_weights(man, seed=302, trained_like=True))
    src, _ = cut(EU, r"^def convert_from_1hot\(", r"^''' SECTION 4")
    env = dict(np=np)
    run(src, env, "EU:convert_from_1hot")
    flair, _ = synth.make_flair(Z_SLICES, H, H, seed=33)
    src, span = cut(EU, r"# Produce 10 results by using 10 different sets of noise", r"vol_dsc_all\.append")
    pred_u = net.predict
    net.predict = lambda xs, **kw: pred_u(xs, **kw).astype(np.float32)
    env.update(my_network=net, noiseSize=32, brain_flair_1tp=flair.copy(), icv_and_sl_mask_2tp=mask.reshape(Z_SLICES, H, H, 1).copy(),
               icv_and_sl_mask_1tp=m1[..., None], brain_wmh_1tp=w1[..., None], brain_wmh_2tp=w2[..., None], brain_cod_2tp=code,
               loaded_data_f_1tp=_Hdr())   # this script keeps every volume as (Z, H, W, 1)
    np.random.seed(4343)
    run(src, env, "EU:evaluation")
    out["EU_eval/lines"] = np.array(span)
    out["EU_eval/np_random_seed"] = np.array(4343)
    out["EU_eval/weight_seed"] = np.array(302)
    out["EU_eval/mean_map"] = env["output_img_pred"]
    out["EU_eval/labels"] = env["output_img_pred_lbl"]
    out["EU_eval/count_out"] = np.array(np.count_nonzero(env["wmh_from_out_2tp"]))
    out["EU_eval/row"] = np.array([float(v) for v in env["vol_dsc"]])


class _Rec:
    """Stand-ins for the step functions, the networks and the logger while the reference's training loop runs."""

    def __init__(self, env):
        self.ev, self.env, self.n_eval = [], env, 0

    def first_id(self, real_1tp):   # batch index of a mini-batch inside the (shuffled) training array of the moment
        pos = int(np.where(self.env["prob_flair_1tp_train"].ravel() == real_1tp.ravel()[0])[0][0])
        return pos // self.env["batchSize"]

    def critic(self, kind):
        def f(inputs):
            real_2tp, real_1tp, noise, ep = inputs
            assert noise.shape == (self.env["batchSize"], self.env["noiseSize"], 1) and ep.shape == (self.env["batchSize"], 1, 1, 1)
            self.ev.append([kind, self.first_id(real_1tp)])
            return 0.25, 0.75
        return f

    def no_update(self, inputs):
        self.n_eval += 1
        self.last_eval = self.first_id(inputs[0])
        return [float(self.n_eval % 7)] + [0.0] * 5

    def train(self, inputs):
        assert self.n_eval == 10 and self.last_eval == self.first_id(inputs[0])   # k_noise evaluations on the same batch
        self.n_eval = 0
        self.ev.append(["gen", self.first_id(inputs[0]), int(self.env["gen_iterations"])])
        return [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]

    # logger
    def log_scalar(self, tag, value, step): self.ev.append(["log", tag, int(step)])
    def log_images(self, tag, images, step, *a): self.ev.append(["img", tag, int(step)])
    # networks
    def predict(self, x, **kw): return np.zeros((1,), np.float32)
    def save(self, path): self.ev.append(["save"])


class _NoShuffleRandom:
    """np.random with shuffle disabled, so batch indices of the trace are positions in the caller's order."""
    def shuffle(self, a): pass
    def normal(self, size=None): return np.zeros(size)
    def uniform(self, size=None): return np.zeros(size)


class _NP:
    random = _NoShuffleRandom()
    def __getattr__(self, k): return getattr(np, k)


def training_loop_trace(out):
    """TG:778-894 (the `for epoch in range(niter)` loop) executed with recording stand-ins: which mini-batch every critic
    update, noise search and generator update touches, every logged tag with its step counter, validation, saves."""
    import time as _time
    src, span = cut(TG, r"^\s+for epoch in range\(niter\):", r"^\s+gen_iterations\+=1", include_end=True)
    out["loop/lines"] = np.array(span)
    for tag, g0, niter, batches in (("warmup_to_steady", 23, 2, 260), ("every_500th", 498, 1, 140), ("from_scratch", 0, 1, 230)):
        bs = 2
        data = np.arange(batches * bs + 1, dtype=np.float32).reshape(-1, 1, 1, 1)   # one sample left over: // batchSize
        env = dict(np=_NP(), time=_time, niter=niter, batchSize=bs, noiseSize=4, Diters=5, gen_iterations=g0, crit_iterations=0,
                   crit_dem_iterations=0, fold=1, errG=0, save_file_name="x", prob_flair_1tp_train=data, prob_2tp_train=data.copy(),
                   prob_flair_1tp_val=np.zeros((3, 2, 2, 1), np.float32), prob_2tp_val=np.zeros((3, 2, 2, 1), np.float32),
                   fixed_noise=np.zeros((3, 4, 1), np.float32), vn=3, vx=2, vy=2, vc=1, t0=0.0)
        rec = _Rec(env)
        env.update(netD_y2_train=rec.critic("y2"), netD_dem_train=rec.critic("dem"), netG_no_update=rec.no_update,
                   netG_train=rec.train, logger=rec, netD_y2=rec, netG=rec)
        run(src, env, "TG:loop")
        out["loop/%s" % tag] = np.array(json.dumps({"g0": g0, "niter": niter, "batches": batches, "batchSize": bs,
                                                    "events": rec.ev, "gen_iterations_end": int(env["gen_iterations"]),
                                                    "crit_iterations_end": int(env["crit_iterations"])}))


def host_functions(out):
    """The reference's NumPy-only helpers, executed: data_prep, data_prep_save, map_image_to_intensity_range (TG:105-149)
    and convert_to_1hot (TU:209-223)."""
    env = dict(np=np)
    src, _ = cut(TG, r"^def data_prep\(", r"^# Calculate Dice coefficient score")
    run(src, env, "TG:host")
    src, _ = cut(TU, r"^def convert_to_1hot\(", r"^''' SECTION 4")
    run(src, env, "TU:1hot")
    rng = np.random.default_rng(77)
    vol = rng.standard_normal((5, 7, 4)).astype(np.float32)

    class _Img:
        image = vol
    out["host/vol"] = vol
    out["host/data_prep"] = env["data_prep"](_Img())
    stack = rng.standard_normal((4, 5, 7, 1)).astype(np.float32)
    out["host/stack"] = stack
    out["host/data_prep_save"] = np.ascontiguousarray(env["data_prep_save"](stack))
    img64 = rng.standard_normal((6, 9)) * 3.0 + 1.0
    out["host/img64"] = img64
    out["host/map_p0"] = env["map_image_to_intensity_range"](img64, 0, 1, percentiles=0)
    out["host/map_p5"] = env["map_image_to_intensity_range"](img64, -1, 1, percentiles=5)
    u8 = rng.integers(0, 256, size=(5, 5)).astype(np.uint8)
    out["host/u8"] = u8
    out["host/map_u8"] = env["map_image_to_intensity_range"](u8, 0, 255, percentiles=2)
    lab = rng.integers(0, 4, size=(2, 3, 4, 1)).astype(np.float32)
    out["host/labels"] = lab
    out["host/onehot"] = env["convert_to_1hot"](lab, 4)


class _Vol:
    def __init__(self, image):
        self.image = image
        self.pixdim = np.array([0.9375, 0.9375, 4.0], dtype=np.float32)


class _FakeOs:
    class path:  # the stroke-lesion files "exist"
        @staticmethod
        def isfile(p): return True


def subject_preparation(out):
    """The per-subject preparation of the two testing scripts, executed on synthetic (X, Y, Z) volumes: masking with the
    intracranial-volume and stroke-lesion masks, clamping, FLAIR normalisation, channel concatenation (EG:533-611) and the
    whole-volume z-score of the DEP-UResNet script (EU:506-538)."""
    rng = np.random.default_rng(91)
    X, Y, Zs = 12, 10, 6
    vols = {"pm1": rng.uniform(-0.2, 1.0, (X, Y, Zs)).astype(np.float32), "pm2": rng.uniform(-0.2, 1.0, (X, Y, Zs)).astype(np.float32),
            "im1": rng.uniform(-0.2, 1.0, (X, Y, Zs)).astype(np.float32), "flair": (rng.standard_normal((X, Y, Zs)) * 90 + 300).astype(np.float32),
            "icv1": (rng.random((X, Y, Zs)) > 0.2).astype(np.float32), "icv2": (rng.random((X, Y, Zs)) > 0.25).astype(np.float32),
            "sl1": (rng.random((X, Y, Zs)) > 0.9).astype(np.float32), "sl2": (rng.random((X, Y, Zs)) > 0.92).astype(np.float32),
            "wmh1": (rng.random((X, Y, Zs)) > 0.8).astype(np.float32), "wmh2": (rng.random((X, Y, Zs)) > 0.8).astype(np.float32),
            "code2": rng.integers(0, 4, (X, Y, Zs)).astype(np.float32)}
    for k, v in vols.items():
        out["prep/vol_" + k] = v
    # ---- EG ----
    env = dict(np=np, os=_FakeOs)
    src, _ = cut(EG, r"^def data_prep\(", r"^# Calculate Dice coefficient score")
    run(src, env, "EG:host")
    dp = env["data_prep"]
    sq = lambda k: np.squeeze(dp(_Vol(vols[k])))   # EG:511-531: data_prep, then np.squeeze
    src, span = cut(EG, r"^\s+# Exclude non-brain tissues", r"# Produce 10 results by using 10 different sets of noise")
    for nicg, PM in ((1, True), (1, False), (2, True)):
        env2 = dict(env)
        env2.update(loaded_image_f_1tp=sq("flair"), loaded_image_im_1tp=sq("im1"), loaded_image_p_1tp=sq("pm1"),
                    loaded_image_p_2tp=sq("pm2"), loaded_image_i_1tp=sq("icv1"), loaded_image_w_1tp=sq("wmh1"),
                    loaded_image_w_2tp=sq("wmh2"), loaded_image_i_2tp=sq("icv2"), loaded_image_c_2tp=sq("code2"),
                    data_list_sl_1tp=["sl1"], data_list_sl_2tp=["sl2"], id=0, load_data=lambda p: _Vol(vols[p]),
                    TRSH_VAL=0.5 if PM else 0.178, nicg=nicg, PM=PM)
        run(src, env2, "EG:prepare")
        tag = "prep/EG_nicg%d_%s" % (nicg, "pm" if PM else "im")
        out[tag + "/x"] = env2["brain_prob__1tp"]
        out[tag + "/mask1"] = env2["icv_and_sl_mask_1tp"]
        out[tag + "/mask2"] = env2["icv_and_sl_mask_2tp"]
    out["prep/EG_lines"] = np.array(span)
    # ---- EU ----
    env = dict(np=np, os=_FakeOs)
    src, _ = cut(EU, r"^def data_prep\(", r"^# Change integer values|^def map_image_to_intensity_range|^# Calculate Dice")
    run(src, env, "EU:host")
    dp = env["data_prep"]
    src, span = cut(EU, r"^\s+# Exclude non-brain tissues", r"# Create output directories for each data")
    env.update(loaded_image_f_1tp=dp(_Vol(vols["flair"])), loaded_image_i_1tp=dp(_Vol(vols["icv1"])),
               loaded_image_w_1tp=dp(_Vol(vols["wmh1"])), loaded_image_w_2tp=dp(_Vol(vols["wmh2"])),
               loaded_image_i_2tp=dp(_Vol(vols["icv2"])), loaded_image_c_2tp=dp(_Vol(vols["code2"])),
               data_list_sl_1tp=["sl1"], data_list_sl_2tp=["sl2"], id=0, load_data=lambda p: _Vol(vols[p]))
    run(src, env, "EU:prepare")
    out["prep/EU/x"] = env["brain_flair_1tp"]
    out["prep/EU/mask1"] = env["icv_and_sl_mask_1tp"]
    out["prep/EU/mask2"] = env["icv_and_sl_mask_2tp"]
    out["prep/EU_lines"] = np.array(span)


def main():
    torch.set_num_threads(8)
    out = {}
    host_functions(out)
    subject_preparation(out)
    training_loop_trace(out)
    gan_training_graph(1, 0.178, out, "gan_im")
    gan_training_graph(2, 0.5, out, "gan_pf")
    generator_topologies(out)
    testing_blocks(out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "%.1f KB" % (OUT.stat().st_size / 1e3), len(out), "arrays")


if __name__ == "__main__":
    main()
