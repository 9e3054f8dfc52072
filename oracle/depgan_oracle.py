"""CPU ORACLE (test infrastructure, NOT product code) for the DEP-GAN / DEP-UResNet hot path.

PARITY: pinned to the reference's own source, EXECUTED.  The reference (febrianrachmadi/dep-gan-im) is four Python-2 /
Keras-2 / TF-1 scripts that cannot be imported in this environment (no tensorflow / keras / h5py; weights and data
absent) and ships no tests or golden vectors.  tests/golden/make_reference_vectors.py therefore cuts the hot path out of
the scripts under /root/reference by anchor lines (networks TG:255-498, loss / optimizer graph construction TG:513-598,
the testing scripts' evaluation blocks, the training loop) and exec's it unmodified on oracle/keras_shim.py, a torch-fp64
stand-in for the few Keras calls those lines make; tests/test_reference_vectors.py holds this file to the resulting
vectors (manifest order, forwards 1e-9, step functions + gradients + Adam updates 1e-8, post-processing exact).  What is
still restated rather than executed: the Keras layers' own arithmetic (inside the shim) and the Keras training phase of
the DEP-UResNet fit.  Further guards: an independent naive NumPy forward (oracle/naive_numpy.py), finite differences of
the three loss graphs (tests/test_oracle.py) and the parameter-count identities of SURVEY.md section 2a.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product (depgan_b200) never does.

Reference citations (TG = DEP-GAN_PROB_IM_twoCritics_training_4fold.py, EG = DEP-GAN_testing_4fold.py,
TU = DEP-UResNet-wNoises-training-4fold.py, EU = DEP-UResNet_testing_4fold.py):
  generator topology ........ TG:349-498 (GAN, tanh head), TU:291-428 (UResNet, softmax head)
  layer helpers ............. TG:255-312
  critic topology ........... TG:316-345
  critic WGAN-GP graphs ..... TG:523-571
  generator loss ............ TG:573-598, dice TG:153-162
  step schedule ............. TG:780-894
  inference repeat loop ..... EG:616-628, EU:553-564
  DEM post-processing ....... EG:673-686 (volume), EG:711-741 (labels)
  UResNet argmax/volume ..... EU:166-185, EU:570, EU:597-600
  Keras Adam ................ keras 2.x optimizers.Adam.get_updates (call sites TG:549,568,594; TU:427)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon (TG:257 et al. use defaults)

# ----------------------------------------------------------------------------------------------------------
# Manifests: (keras layer name, weight name, shape) in Keras layouts.
# ----------------------------------------------------------------------------------------------------------

# (suffix, multiplier of first_fm) for the 14 FiLM heads in the order the reference creates them (TG:362-395)
FILM_HEADS = [("_m3", 3), ("_m2", 2), ("_m1", 1), ("", 4), ("_p3", 3), ("_p2", 2), ("_p1", 1)]

# ResBlocks in graph order: (conv_in, conv_noise, conv_out, film suffix, width multiplier, input kind)
# TG:398-491.  Numbering follows the reference's layer names (jumps 5->8, 11->14).
GEN_BLOCKS = [
    ("gen_0", "gen_noise_m1", "gen_1", "_m1", 1),
    ("gen_2", "gen_noise_m2", "gen_3", "_m2", 2),
    ("gen_4", "gen_noise_m3", "gen_5", "_m3", 3),
    ("gen_8", "gen_noise_p4", "gen_9", "", 4),
    ("gen_10", "gen_noise_p3", "gen_11", "_p3", 3),
    ("gen_14", "gen_noise_p2", "gen_15", "_p2", 2),
    ("gen_16", "gen_noise_p1", "gen_17", "_p1", 1),
]
GEN_DECONVS = ["de_gen_9", "de_gen_11", "de_gen_15"]  # TG:449,464,478 (filters 4f,3f,2f)

CRITIC_CONVS = [  # (name, ksize, cin, cout)  TG:319-338
    ("conv2d_dis_0a", 5, 1, 16), ("conv2d_dis_0b", 5, 16, 16),
    ("conv2d_dis_1a", 5, 16, 32), ("conv2d_dis_1b", 5, 32, 32),
    ("conv2d_dis_2", 3, 32, 64), ("conv2d_dis_3", 3, 64, 64),
    ("conv2d_dis_4", 3, 64, 128), ("conv2d_dis_5", 3, 128, 128),
    ("conv2d_dis_6", 3, 128, 256), ("conv2d_dis_7", 3, 256, 256), ("conv2d_dis_8", 3, 256, 256),
]
CRITIC_POOL_AFTER = {"conv2d_dis_0b", "conv2d_dis_1b", "conv2d_dis_3", "conv2d_dis_5"}  # TG:321,325,329,333


def _bn(name, c):
    return [(name, "gamma", (c,)), (name, "beta", (c,)), (name, "moving_mean", (c,)), (name, "moving_variance", (c,))]


def gen_manifest(nicg=1, nc_out=1, first_fm=32, noise_len=32):
    """Weight tensors of Gen_UNet2D (TG:349-498) as (layer, weight, shape), Keras layouts."""
    f = first_fm
    m = []
    m += [("dense_noise_1_add_f0", "kernel", (1, f)), ("dense_noise_1_add_f0", "bias", (f,))]
    m += _bn("dense_bn_noise_1_add_f0", f)
    m += [("dense_noise_1_add_f1", "kernel", (f, f)), ("dense_noise_1_add_f1", "bias", (f,))]
    m += _bn("dense_bn_noise_1_add_f1", f)
    flat = noise_len * f
    for suf, mult in FILM_HEADS:
        for kind in ("add", "mul"):
            n = "noise_2_%s%s" % (kind, suf)
            m += [("dense_" + n, "kernel", (flat, f * mult)), ("dense_" + n, "bias", (f * mult,))]
            m += _bn("dense_bn_" + n, f * mult)
    cin = nicg
    skip_c = []
    for bi, (c_in, c_noise, c_out, suf, mult) in enumerate(GEN_BLOCKS):
        c = f * mult
        if bi >= 4:  # decoder: input is concat [deconv_out, skip]  (TG:450,465,479)
            cin = cin + skip_c[6 - bi]
        for nm, ci in ((c_in, cin), (c_noise, c), (c_out, c)):
            m += [("conv2d_" + nm, "kernel", (3, 3, ci, c)), ("conv2d_" + nm, "bias", (c,))]
            m += _bn("bn_" + nm, c)
        if bi < 3:
            skip_c.append(c)
        if 3 <= bi < 6:
            d = GEN_DECONVS[bi - 3]
            # Conv2DTranspose kernel layout (kh, kw, Cout, Cin), filters = same width (TG:449,464,478)
            m += [("deconv2d_" + d, "kernel", (2, 2, c, c)), ("deconv2d_" + d, "bias", (c,))]
            m += _bn("bn_" + d, c)
        cin = c
    m += [("gen_segmentation", "kernel", (1, 1, f, nc_out)), ("gen_segmentation", "bias", (nc_out,))]
    return m


def critic_manifest(h=256, w=256):
    """Weight tensors of Dis_C2D_FCN1 (TG:316-345).  Flatten width = (h/16)*(w/16) = 256 at 256x256."""
    m = []
    for name, k, ci, co in CRITIC_CONVS:
        m += [(name, "kernel", (k, k, ci, co)), (name, "bias", (co,))]
    m += [("dis_9", "kernel", (1, 1, 256, 1)), ("dis_9", "bias", (1,))]
    m += [("dense_1", "kernel", ((h // 16) * (w // 16), 1)), ("dense_1", "bias", (1,))]
    return m


def manifest_count(m):
    return int(sum(int(np.prod(s)) for _, _, s in m))


def is_trainable(weight_name):
    return weight_name not in ("moving_mean", "moving_variance")


# ----------------------------------------------------------------------------------------------------------
# Parameter handling
# ----------------------------------------------------------------------------------------------------------

def to_torch(params, dtype=torch.float64, requires_grad=False):
    """dict 'layer/weight' -> numpy  ==>  OrderedDict 'layer/weight' -> torch leaf tensors."""
    out = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad and is_trainable(k.split("/")[1]):
            t.requires_grad_(True)
        out[k] = t
    return out


def _bn_apply(P, name, x, channel_dim):
    """Keras inference-mode BN (learning phase 0): gamma*(x-mean)/sqrt(var+eps)+beta."""
    g, b = P[name + "/gamma"], P[name + "/beta"]
    mu, var = P[name + "/moving_mean"], P[name + "/moving_variance"]
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mu.view(shape)) * (g * inv).view(shape) + b.view(shape)


def _conv(P, name, x, pad):
    """Keras Conv2D 'same', stride 1.  x is NCHW here; kernel HWIO -> OIHW."""
    w = P[name + "/kernel"].permute(3, 2, 0, 1)
    return F.conv2d(x, w, P[name + "/bias"], padding=pad)


def _deconv(P, name, x):
    """Keras Conv2DTranspose k=2, s=2, 'valid'; kernel (kh,kw,Cout,Cin) -> torch (Cin,Cout,kh,kw)."""
    w = P[name + "/kernel"].permute(3, 2, 0, 1)
    return F.conv_transpose2d(x, w, P[name + "/bias"], stride=2)


def film_params(P, z):
    """Noise path TG:353-395.  z (N, L, 1) -> dict suffix -> (gamma (N,C), beta (N,C))."""
    h = z @ P["dense_noise_1_add_f0/kernel"] + P["dense_noise_1_add_f0/bias"]  # (N,L,f), Dense on last axis
    h = torch.relu(_bn_apply(P, "dense_bn_noise_1_add_f0", h, 2))
    h = h @ P["dense_noise_1_add_f1/kernel"] + P["dense_noise_1_add_f1/bias"]
    h = torch.relu(_bn_apply(P, "dense_bn_noise_1_add_f1", h, 2))
    h = h.reshape(h.shape[0], -1)  # Flatten: row-major (token, feature)   TG:360
    out = {}
    for suf, _ in FILM_HEADS:
        r = []
        for kind in ("mul", "add"):
            n = "noise_2_%s%s" % (kind, suf)
            v = h @ P["dense_" + n + "/kernel"] + P["dense_" + n + "/bias"]
            r.append(_bn_apply(P, "dense_bn_" + n, v, 1))
        out[suf] = (r[0], r[1])
    return out


def gen_forward(P, x_nhwc, z, head="tanh", return_acts=False):
    """Gen_UNet2D forward in Keras learning-phase 0 (Dropout = identity, BN = moving stats).

    x_nhwc (N,H,W,nicg), z (N,L,1).  Returns (N,H,W,nc_out): tanh head (TG:494-495) or softmax (TU:423-424).
    """
    film = film_params(P, z)
    x = x_nhwc.permute(0, 3, 1, 2)
    acts = {}
    skips = []
    for bi, (c_in, c_noise, c_out, suf, mult) in enumerate(GEN_BLOCKS):
        a = torch.relu(_bn_apply(P, "bn_" + c_in, _conv(P, "conv2d_" + c_in, x, 1), 1))        # conv2d_bn_relu
        y = _bn_apply(P, "bn_" + c_noise, _conv(P, "conv2d_" + c_noise, a, 1), 1)             # conv2d_bn
        gam, bet = film[suf]
        b = torch.relu(y * gam[:, :, None, None] + bet[:, :, None, None])                     # mul, add, relu
        r = b + a                                                                             # add_noiseZres
        o = torch.relu(_bn_apply(P, "bn_" + c_out, _conv(P, "conv2d_" + c_out, r, 1), 1))
        acts[c_in], acts[c_noise], acts[c_out] = a, b, o
        if bi < 3:
            skips.append(o)
            x = F.max_pool2d(o, 2)                                                            # MaxPooling2D(2,2)
        elif bi < 6:
            d = GEN_DECONVS[bi - 3]
            u = torch.relu(_bn_apply(P, "bn_" + d, _deconv(P, "deconv2d_" + d, o), 1))        # deconv2d_bn_relu
            x = torch.cat([u, skips[5 - bi]], dim=1)                                          # [deconv, skip]
            acts[d] = u
        else:
            x = o
    seg = _conv(P, "gen_segmentation", x, 0)
    if head == "tanh":
        out = torch.tanh(seg)
    elif head == "softmax":
        out = torch.softmax(seg, dim=1)
    else:
        out = seg
    out = out.permute(0, 2, 3, 1)
    if return_acts:
        return out, acts
    return out


def critic_forward(P, x_nhwc):
    """Dis_C2D_FCN1 forward TG:316-345.  (N,H,W,1) -> (N,1).  H=W=256 gives Flatten(256)."""
    x = x_nhwc.permute(0, 3, 1, 2)
    for name, k, ci, co in CRITIC_CONVS:
        x = torch.relu(_conv(P, name, x, k // 2))
        if name in CRITIC_POOL_AFTER:
            x = F.max_pool2d(x, 2)
    x = _conv(P, "dis_9", x, 0)                    # (N,1,h,w)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)   # Keras Flatten on NHWC
    return flat @ P["dense_1/kernel"] + P["dense_1/bias"]


# ----------------------------------------------------------------------------------------------------------
# Loss graphs  (TG:523-598)
# ----------------------------------------------------------------------------------------------------------

def _base(x1):
    """net_G_real_IM: channel 0 of the generator input, reshaped (N,H,W,1).  TG:528-529."""
    return x1[..., 0:1]


def critic_loss(PD, PG, real2, x1, z, ep, which, delta=10.0):
    """Y2-critic (which='y2', TG:532-552) or DEM-critic (which='dem', TG:554-571) WGAN-GP loss.

    Returns (loss, loss_real, loss_fake, grad_penalty).  The generator runs without gradient tracking
    because only critic weights are updated (TG:549, 568).
    """
    with torch.no_grad():
        dem = gen_forward(PG, x1, z)
    base = _base(x1)
    if which == "y2":
        real, fake = real2, base + dem                                   # TG:534
    else:
        real, fake = real2 - base, dem                                   # TG:530, 557
    mixed = (ep * real + (1 - ep) * fake).detach().requires_grad_(True)  # TG:536-538 / 556-557
    loss_real = critic_forward(PD, real).mean()                          # TG:540 / 559
    loss_fake = critic_forward(PD, fake).mean()                          # TG:541 / 560
    d_mixed = critic_forward(PD, mixed)
    g = torch.autograd.grad(d_mixed.sum(), mixed, create_graph=True)[0]  # TG:543 / 562
    norm = torch.sqrt((g * g).sum(dim=(1, 2, 3)))                        # TG:544 / 563
    gp = ((norm - 1) ** 2).mean()                                        # TG:545 / 564
    loss = loss_fake - loss_real + delta * gp                            # TG:547 / 566
    return loss, loss_real, loss_fake, gp


def dice_coef(a, b, smooth=1e-7):
    """TG:153-157, batch-global."""
    inter = (a * b).sum()
    return (2.0 * inter + smooth) / (a.sum() + b.sum() + smooth)


def gen_loss(PG, PDy2, PDdem, x1, real2, z, thr, dM1=100.0, dM3=100.0, dM4=1.0):
    """Generator loss TG:573-592.  Returns [loss, loss_fake, loss_fake_dem, M1, M3, M4] (TG:595-598)."""
    dem = gen_forward(PG, x1, z)
    base = _base(x1)
    fake2 = base + dem
    real_dem = real2 - base
    loss_fake = critic_forward(PDy2, fake2).mean()
    loss_fake_dem = critic_forward(PDdem, dem).mean()
    m1 = (dem - real_dem).abs().mean() * dM1                               # TG:576
    # thresholds are compared in float32 in the reference (tf.float32 graph, TG:581-582)
    thr32 = np.float32(thr)
    wr = (real2.detach().to(torch.float32) >= float(thr32)).to(dem.dtype)  # TG:581
    wf = (fake2.detach().to(torch.float32) >= float(thr32)).to(dem.dtype)  # TG:582
    m4 = (1.0 - dice_coef(wr, wf)) * dM4                                   # TG:583
    m3 = ((wr.sum() / 1000.0 - wf.sum() / 1000.0) ** 2) * dM3              # TG:587-589
    loss = (-loss_fake) + (-loss_fake_dem) + m1 + m3 + m4                  # TG:592
    return [loss, loss_fake, loss_fake_dem, m1, m3, m4]


# ----------------------------------------------------------------------------------------------------------
# Keras-form Adam and the four step callables
# ----------------------------------------------------------------------------------------------------------

class KerasAdam:
    """keras.optimizers.Adam.get_updates (Keras 2.x): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    p -= lr_t * m / (sqrt(v) + eps), eps = K.epsilon() = 1e-7.  One instance per network (TG:549,568,594)."""

    def __init__(self, params, lr=1e-4, beta_1=0.0, beta_2=0.9, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.iterations = 0
        self.keys = [k for k in params if is_trainable(k.split("/")[1])]
        self.m = {k: torch.zeros_like(params[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(params[k]) for k in self.keys}

    def step(self, params, grads):
        t = self.iterations + 1
        lr_t = self.lr * (math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))
        with torch.no_grad():
            for k in self.keys:
                g = grads[k]
                self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
                self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
                params[k] -= lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)
        self.iterations = t


class OracleTrainer:
    """State + the four K.function callables of TG:550-552, 569-571, 595-598, with the reference's argument
    orders: critics take [real_2tp, real_1tp, noise, ep]; generator functions take [real_1tp, real_2tp, noise]."""

    def __init__(self, PG, PDy2, PDdem, thr, dtype=torch.float64, lr=1e-4):
        self.dtype = dtype
        self.PG = to_torch(PG, dtype, True)
        self.PDy2 = to_torch(PDy2, dtype, True)
        self.PDdem = to_torch(PDdem, dtype, True)
        self.thr = thr
        self.optG = KerasAdam(self.PG, lr)
        self.optDy2 = KerasAdam(self.PDy2, lr)
        self.optDdem = KerasAdam(self.PDdem, lr)
        self.last_grads = None

    def _t(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self.dtype)

    def _critic_train(self, PD, opt, which, inputs, update=True):
        real2, x1, z, ep = [self._t(a) for a in inputs]
        loss, lr_, lf_, gp = critic_loss(PD, self.PG, real2, x1, z, ep, which)
        keys = opt.keys
        grads = torch.autograd.grad(loss, [PD[k] for k in keys], allow_unused=True)
        gd = {k: (g if g is not None else torch.zeros_like(PD[k])) for k, g in zip(keys, grads)}
        self.last_grads = {k: v.detach().clone() for k, v in gd.items()}
        self.last_gp = float(gp.detach())
        if update:
            opt.step(PD, gd)
        return [float(lr_.detach()), float(lf_.detach())]

    def netD_y2_train(self, inputs, update=True):
        return self._critic_train(self.PDy2, self.optDy2, "y2", inputs, update)

    def netD_dem_train(self, inputs, update=True):
        return self._critic_train(self.PDdem, self.optDdem, "dem", inputs, update)

    def netG_no_update(self, inputs):
        x1, real2, z = [self._t(a) for a in inputs]
        with torch.no_grad():
            out = gen_loss(self.PG, self.PDy2, self.PDdem, x1, real2, z, self.thr)
        return [float(v) for v in out]

    def netG_train(self, inputs, update=True):
        x1, real2, z = [self._t(a) for a in inputs]
        out = gen_loss(self.PG, self.PDy2, self.PDdem, x1, real2, z, self.thr)
        keys = self.optG.keys
        grads = torch.autograd.grad(out[0], [self.PG[k] for k in keys], allow_unused=True)
        gd = {k: (g if g is not None else torch.zeros_like(self.PG[k])) for k, g in zip(keys, grads)}
        self.last_grads = {k: v.detach().clone() for k, v in gd.items()}
        if update:
            self.optG.step(self.PG, gd)
        return [float(v.detach()) for v in out]

    def gen_iteration(self, crit_y2_batches, crit_dem_batches, x1, real2, noises):
        """One steady-state generator iteration TG:796-878: critic updates on the given batches
        (each item = [real_2tp, real_1tp, noise, ep]), then k_noise forward-only evaluations, argmin, update."""
        for b in crit_y2_batches:
            self.netD_y2_train(b)
        for b in crit_dem_batches:
            self.netD_dem_train(b)
        losses = [self.netG_no_update([x1, real2, nz])[0] for nz in noises]      # TG:871-874
        k = int(np.array(losses).argmin(0))                                      # TG:875-876
        return k, losses, self.netG_train([x1, real2, noises[k]])                # TG:877-878


# ----------------------------------------------------------------------------------------------------------
# Inference loop and post-processing (NumPy, dtype-faithful to the reference)
# ----------------------------------------------------------------------------------------------------------

def predict(P, x, z, head="tanh", dtype=torch.float32, batch_size=32):
    """Keras model.predict([x, z]) semantics: float32 out, internally batched by 32 (EG:621)."""
    Pt = P if isinstance(next(iter(P.values())), torch.Tensor) else to_torch(P, dtype)
    outs = []
    with torch.no_grad():
        for i in range(0, x.shape[0], batch_size):
            xb = torch.as_tensor(np.asarray(x[i:i + batch_size]), dtype=dtype)
            zb = torch.as_tensor(np.asarray(z[i:i + batch_size]), dtype=dtype)
            outs.append(gen_forward(Pt, xb, zb, head).to(torch.float32).numpy())
    return np.concatenate(outs, 0)


def inference_mean(preds_f32, mask_f32):
    """EG:616-628 / EU:553-564: float64 accumulator of float32 (prediction * mask), divided by n_repeat.

    preds_f32: list of (Z,H,W) [GAN, squeezed] or (Z,H,W,C) [UResNet] float32 predictions.
    mask_f32: (Z,H,W) float32 for the GAN path; for the UResNet path the reference multiplies a (Z,H,W,4)
    prediction by the mask array as loaded (EU:559) -- the caller passes a broadcast-compatible mask.
    """
    acc = np.zeros(preds_f32[0].shape)                      # float64, np.zeros default (EG:617)
    for p in preds_f32:
        acc = acc + np.multiply(p, mask_f32)                # f32*f32 -> f32, then f64 + f32 -> f64
    return acc / float(len(preds_f32))


def dem_postproc(base_f32, dem_f64, mask_f32, thr):
    """EG:673-686 (volume count) and EG:711-741 (labels).  Returns (count:int, labels:float64 (Z,H,W), fake2)."""
    fake2 = base_f32 + dem_f64                              # f32 + f64 -> f64 (EG:675)
    fake2[fake2 < -1] = -1
    fake2[fake2 > 1] = 1
    wmh = np.zeros(fake2.shape)
    wmh[fake2 > thr] = 1                                    # strict > (EG:679)
    wmh = np.multiply(mask_f32, wmh)
    count = int(np.count_nonzero(wmh))
    labels = np.zeros(fake2.shape)
    b_ge = base_f32 >= thr                                  # f32 array vs python float (NumPy casts thr to f32)
    b_lt = base_f32 < thr
    f_ge = fake2 >= thr
    f_lt = fake2 < thr
    labels[np.all([f_lt, b_ge], axis=0)] = 1                # shrink  EG:723-727
    labels[np.all([f_ge, b_lt], axis=0)] = 2                # grow    EG:730-734
    labels[np.all([f_ge, b_ge], axis=0)] = 3                # stay    EG:737-741
    return count, labels, fake2


def uresnet_labels(prob_mean_f64):
    """EU:570 via convert_from_1hot EU:166-185 (argmax, first max wins, uint8) and EU:597-600 count(label>0)."""
    lab = np.argmax(prob_mean_f64.reshape(-1, prob_mean_f64.shape[-1]), axis=1).astype(np.uint8)
    lab = lab.reshape(prob_mean_f64.shape[:-1])
    return lab, int(np.count_nonzero(lab > 0))


# ----------------------------------------------------------------------------------------------------------
# DEP-UResNet supervised training step (TU:291-428 compile at TU:427, fit at TU:602-606): Keras *training* phase:
# BatchNormalization normalises with the batch statistics (biased variance) and updates its moving statistics
# (momentum 0.99, Bessel-corrected batch variance), Dropout(0.25) `do_gen_1` after conv2d_gen_10 (TU:388), loss =
# keras.losses.categorical_crossentropy on the softmax output, optimizer Adam(lr=1e-4) with Keras defaults
# beta_1=0.9, beta_2=0.999.  The Dropout mask is an explicit input (the reference draws it inside TF, unseeded).
# ----------------------------------------------------------------------------------------------------------
BN_MOMENTUM = 0.99
DROP_RATE = 0.25


def _bn_train(P, name, x, channel_dim, stats):
    g, b = P[name + "/gamma"], P[name + "/beta"]
    dims = [d for d in range(x.dim()) if d != channel_dim]
    mean = x.mean(dim=dims)
    var = x.var(dim=dims, unbiased=False)
    m = x.numel() // x.shape[channel_dim]
    stats[name] = (mean.detach(), (var * m / max(m - 1, 1)).detach())
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    return (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS) * g.view(shape) + b.view(shape)


def film_params_train(P, z, stats):
    h = z @ P["dense_noise_1_add_f0/kernel"] + P["dense_noise_1_add_f0/bias"]
    h = torch.relu(_bn_train(P, "dense_bn_noise_1_add_f0", h, 2, stats))
    h = h @ P["dense_noise_1_add_f1/kernel"] + P["dense_noise_1_add_f1/bias"]
    h = torch.relu(_bn_train(P, "dense_bn_noise_1_add_f1", h, 2, stats))
    h = h.reshape(h.shape[0], -1)
    out = {}
    for suf, _ in FILM_HEADS:
        r = []
        for kind in ("mul", "add"):
            n = "noise_2_%s%s" % (kind, suf)
            v = h @ P["dense_" + n + "/kernel"] + P["dense_" + n + "/bias"]
            r.append(_bn_train(P, "dense_bn_" + n, v, 1, stats))
        out[suf] = (r[0], r[1])
    return out


def gen_forward_train(P, x_nhwc, z, drop_mask, head="softmax"):
    """Gen_UNet2D (DEP-UResNet variant: a single Dropout, TU:388) in Keras training phase.
    drop_mask: (N, H/4, W/4, 96) of {0,1} keep flags for `do_gen_1`; kept values are scaled by 1/(1-rate).
    Returns (output NHWC, stats dict layer -> (batch mean, unbiased batch variance))."""
    stats = {}
    film = film_params_train(P, z, stats)
    x = x_nhwc.permute(0, 3, 1, 2)
    skips = []
    for bi, (c_in, c_noise, c_out, suf, mult) in enumerate(GEN_BLOCKS):
        a = torch.relu(_bn_train(P, "bn_" + c_in, _conv(P, "conv2d_" + c_in, x, 1), 1, stats))
        if c_in == "gen_10":
            a = a * drop_mask.permute(0, 3, 1, 2) / (1.0 - DROP_RATE)
        y = _bn_train(P, "bn_" + c_noise, _conv(P, "conv2d_" + c_noise, a, 1), 1, stats)
        gam, bet = film[suf]
        r = torch.relu(y * gam[:, :, None, None] + bet[:, :, None, None]) + a
        o = torch.relu(_bn_train(P, "bn_" + c_out, _conv(P, "conv2d_" + c_out, r, 1), 1, stats))
        if bi < 3:
            skips.append(o)
            x = F.max_pool2d(o, 2)
        elif bi < 6:
            d = GEN_DECONVS[bi - 3]
            u = torch.relu(_bn_train(P, "bn_" + d, _deconv(P, "deconv2d_" + d, o), 1, stats))
            x = torch.cat([u, skips[5 - bi]], dim=1)
        else:
            x = o
    seg = _conv(P, "gen_segmentation", x, 0)
    out = torch.softmax(seg, dim=1) if head == "softmax" else seg
    return out.permute(0, 2, 3, 1), stats


def categorical_crossentropy(target, output, eps=1e-7):
    """keras.backend.categorical_crossentropy (TF backend, from_logits=False), mean over N,H,W as Keras' fit reports."""
    output = output / output.sum(dim=-1, keepdim=True)
    output = output.clamp(eps, 1.0 - eps)
    return (-(target * torch.log(output)).sum(dim=-1)).mean()


def uresnet_train_step(P, x, z, onehot, drop_mask, opt=None):
    """One Keras train_on_batch of the compiled DEP-UResNet (TU:427, 602).  P: torch leaves (requires_grad on
    trainables).  Returns (loss, grads dict, stats); when `opt` (KerasAdam with beta_1=0.9, beta_2=0.999) is given
    the parameters and the BN moving statistics are updated in place."""
    out, stats = gen_forward_train(P, x, z, drop_mask)
    loss = categorical_crossentropy(onehot, out)
    keys = [k for k in P if P[k].requires_grad]
    grads = torch.autograd.grad(loss, [P[k] for k in keys], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(keys, grads)}
    if opt is not None:
        opt.step(P, gd)
        with torch.no_grad():
            for name, (mean, var_unb) in stats.items():
                P[name + "/moving_mean"].mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * mean)
                P[name + "/moving_variance"].mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * var_unb)
    return float(loss.detach()), {k: v.detach() for k, v in gd.items()}, stats


# ---------------------------------------------------------------------------------------------------------
# Evaluation row (EG:688-807, identical in EU:601-704) -- literal NumPy restatement on the label volumes
# ---------------------------------------------------------------------------------------------------------
def evaluation_row(fake_labels, real_labels, vol_1tp_ml, vol_2tp_ml, vol_out_ml):
    """[true_pred, prog, true_prog, regg, true_regg, vol1, vol2, vol_out, mse, err, d5, d6, avg56, d1, d2, d3, d4,
    avg123] (EG:806-807).  fake_labels / real_labels: arrays with values 0..3 (EG:713-741; brain_code_2tp)."""
    wmh_change_mask_fake = np.squeeze(np.asarray(fake_labels))
    wmh_change_mask_real = np.squeeze(np.asarray(real_labels))
    err_vol = vol_out_ml - vol_2tp_ml
    mse_vol = np.mean((vol_2tp_ml - vol_out_ml) ** 2)
    true_pred = true_prog = true_regg = prog = regg = 0
    if (vol_2tp_ml - vol_1tp_ml) >= 0:
        prog = 1
        if vol_out_ml - vol_1tp_ml >= 0:
            true_pred = 1
            true_prog = 1
    else:
        regg = 1
        if vol_out_ml - vol_1tp_ml < 0:
            true_pred = 1
            true_regg = 1
    smooth = 1e-7

    def dsc(fake, real, k):  # EG:745-794, one expression for all six
        return (np.count_nonzero(fake[real == k] == k) * 2.0 + smooth) / \
               (smooth + np.count_nonzero(real[real == k] == k) + np.count_nonzero(fake[fake == k] == k))

    dice_1 = dsc(wmh_change_mask_fake, wmh_change_mask_real, 1)
    dice_2 = dsc(wmh_change_mask_fake, wmh_change_mask_real, 2)
    dice_3 = dsc(wmh_change_mask_fake, wmh_change_mask_real, 3)
    dice_4 = dsc(wmh_change_mask_fake > 0, wmh_change_mask_real > 0, 1)
    temp_a_fake = (wmh_change_mask_fake == 1).astype(np.int64) + (wmh_change_mask_fake == 2)
    temp_a_real = (wmh_change_mask_real == 1).astype(np.int64) + (wmh_change_mask_real == 2)
    dice_5 = dsc(temp_a_fake > 0, temp_a_real > 0, 1)
    dice_6 = dsc(wmh_change_mask_fake == 3, wmh_change_mask_real == 3, 1)
    avg_all_dice = (dice_1 + dice_2 + dice_3) / 3.0
    avg_dice__56 = (dice_5 + dice_6) / 2.0
    return [true_pred, prog, true_prog, regg, true_regg, vol_1tp_ml, vol_2tp_ml, vol_out_ml, mse_vol, err_vol, dice_5,
            dice_6, avg_dice__56, dice_1, dice_2, dice_3, dice_4, avg_all_dice]
