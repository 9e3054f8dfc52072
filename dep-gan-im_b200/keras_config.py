"""The Keras side of a saved model: `model_config` / `training_config` JSON and Keras' own layer order.

`netG.save(path)` (TG:892) and `my_network.save(path)` (TU:622) write a *full-model* HDF5 file: next to `/model_weights`
there is a `model_config` attribute (the functional-API graph as JSON, what `model.to_json()` of TU:623 returns and what
`keras.models.load_model` rebuilds the network from) and, for the compiled DEP-UResNet, `training_config` plus
`/optimizer_weights`.  This module restates the two network graphs as data (one row per residual block instead of the
reference's line-by-line construction, TG:349-498 / TU:291-428 / TG:316-345), derives

  * the layer list in the order Keras itself uses (`Network._init_graph` of Keras 2.2.x: depth-first traversal from the
    outputs assigns each layer an index, layers are grouped by depth = longest path to an output, deepest first, ties by
    traversal index) -- this is the order of `model.layers`, of the `layer_names` attribute and therefore of the
    position-based `load_weights` of EG:383 / EU:402;
  * the per-layer `config` dictionaries with the Keras 2.2.4 defaults.

No Keras is available here, so the JSON is checked structurally (tests/test_keras_config.py):
names, classes, inbound nodes, weight-bearing layers and their order against the native manifest; loading it with a real
`keras.models.load_model` is untested and INTEGRATION.md says so.
"""
from __future__ import annotations

import json

KERAS_VERSION = "2.2.4"
BACKEND = "tensorflow"

# one row per residual block: (in conv, noise conv, out conv, FiLM dense suffix, tag of the mul/add/relu layers,
# channel multiplier, DEP-GAN dropout names (after the in conv, after the noise conv))
_BLOCKS = [
    ("gen_0", "gen_noise_m1", "gen_1", "_m1", "m1", 1, ("do_gen_a3", "do_gen_b3")),
    ("gen_2", "gen_noise_m2", "gen_3", "_m2", "m2", 2, ("do_gen_a2", "do_gen_b2")),
    ("gen_4", "gen_noise_m3", "gen_5", "_m3", "m3", 3, ("do_gen_a1", "do_gen_b1")),
    ("gen_8", "gen_noise_p4", "gen_9", "", "p4", 4, ("do_gen_0a", "do_gen_0b")),
    ("gen_10", "gen_noise_p3", "gen_11", "_p3", "p3", 3, ("do_gen_1a", "do_gen_1b")),
    ("gen_14", "gen_noise_p2", "gen_15", "_p2", "p2", 2, ("do_gen_2a", "do_gen_2b")),
    ("gen_16", "gen_noise_p1", "gen_17", "_p1", "p1", 1, ("do_gen_3a", "do_gen_3b")),
]
_DECONV = {3: ("de_gen_9", "concat_gen_0", 2), 4: ("de_gen_11", "concat_gen_1", 1), 5: ("de_gen_15", "concat_gen_3", 0)}
_FILM_ORDER = ["_m3", "_m2", "_m1", "", "_p3", "_p2", "_p1"]  # creation order of the FiLM heads (TG:361-395)

_GLOROT = {"class_name": "VarianceScaling", "config": {"scale": 1.0, "mode": "fan_avg", "distribution": "uniform", "seed": None}}
_HE = {"class_name": "VarianceScaling", "config": {"scale": 2.0, "mode": "fan_in", "distribution": "normal", "seed": None}}
_ZEROS, _ONES = {"class_name": "Zeros", "config": {}}, {"class_name": "Ones", "config": {}}


class _Graph:
    """Layers in creation order, each called exactly once (so a layer and its single node coincide)."""

    def __init__(self, name):
        self.name, self.layers, self.by_name = name, [], {}

    def add(self, class_name, name, config, inbound=(), weights=()):
        cfg = {"name": name}
        if class_name != "InputLayer":
            cfg["trainable"] = True
        cfg.update(config)
        layer = {"name": name, "class_name": class_name, "config": cfg, "inbound": list(inbound), "weights": list(weights)}
        self.layers.append(layer)
        self.by_name[name] = layer
        return name

    # ---- layer constructors with the Keras 2.2.4 default configs ----
    def input(self, name, shape):
        return self.add("InputLayer", name, {"batch_input_shape": [None] + list(shape), "dtype": "float32", "sparse": False})

    def _conv_cfg(self, filters, k, padding, init, strides=(1, 1)):
        return {"filters": filters, "kernel_size": [k, k], "strides": list(strides), "padding": padding,
                "data_format": "channels_last", "dilation_rate": [1, 1], "activation": "linear", "use_bias": True,
                "kernel_initializer": init, "bias_initializer": _ZEROS, "kernel_regularizer": None,
                "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None, "bias_constraint": None}

    def conv(self, name, x, filters, k, padding="same", init=_GLOROT):
        return self.add("Conv2D", name, self._conv_cfg(filters, k, padding, init), [x], ["kernel", "bias"])

    def deconv(self, name, x, filters, k):
        cfg = self._conv_cfg(filters, k, "valid", _GLOROT, strides=(2, 2))
        del cfg["dilation_rate"]
        cfg["output_padding"] = None
        return self.add("Conv2DTranspose", name, cfg, [x], ["kernel", "bias"])

    def dense(self, name, x, units):
        return self.add("Dense", name, {"units": units, "activation": "linear", "use_bias": True, "kernel_initializer": _HE,
                                        "bias_initializer": _ZEROS, "kernel_regularizer": None, "bias_regularizer": None,
                                        "activity_regularizer": None, "kernel_constraint": None, "bias_constraint": None},
                        [x], ["kernel", "bias"])

    def bn(self, name, x):
        return self.add("BatchNormalization", name,
                        {"axis": -1, "momentum": 0.99, "epsilon": 0.001, "center": True, "scale": True,
                         "beta_initializer": _ZEROS, "gamma_initializer": _ONES, "moving_mean_initializer": _ZEROS,
                         "moving_variance_initializer": _ONES, "beta_regularizer": None, "gamma_regularizer": None,
                         "beta_constraint": None, "gamma_constraint": None},
                        [x], ["gamma", "beta", "moving_mean", "moving_variance"])

    def act(self, name, x, fn):
        return self.add("Activation", name, {"activation": fn}, [x])

    def dropout(self, name, x, rate=0.25):
        return self.add("Dropout", name, {"rate": rate, "noise_shape": None, "seed": None}, [x])

    def pool(self, name, x):
        return self.add("MaxPooling2D", name, {"pool_size": [2, 2], "padding": "valid", "strides": [2, 2],
                                               "data_format": "channels_last"}, [x])

    def flatten(self, name, x):
        return self.add("Flatten", name, {"data_format": "channels_last"}, [x])

    def merge(self, cls, name, xs, **extra):
        return self.add(cls, name, dict(extra), xs)

    # ---- Keras' ordering (keras/engine/network.py, Network._init_graph) ----
    def keras_order(self, outputs):
        index, finished, post = {}, set(), []

        def visit(name):
            # iterative depth-first traversal in the order of the inbound tensors (the graphs are ~170 layers deep)
            if name in finished:
                return
            index.setdefault(name, len(index))
            stack = [(name, iter(self.by_name[name]["inbound"]))]
            while stack:
                cur, it = stack[-1]
                nxt = next(it, None)
                if nxt is None:
                    finished.add(cur)
                    post.append(cur)
                    stack.pop()
                elif nxt not in finished and nxt not in index:
                    index[nxt] = len(index)
                    stack.append((nxt, iter(self.by_name[nxt]["inbound"])))
                # a layer already indexed but unfinished would be a cycle; already finished: nothing to do
        for o in outputs:
            visit(o)
        depth = {}
        for name in reversed(post):  # consumers before producers
            d = depth.setdefault(name, 0)
            for src in self.by_name[name]["inbound"]:
                depth[src] = max(depth.get(src, 0), d + 1)
        return sorted(post, key=lambda n: (-depth[n], index[n]))

    def model_config(self, inputs, outputs):
        order = self.keras_order(outputs)
        layers = [{"name": n, "class_name": self.by_name[n]["class_name"], "config": self.by_name[n]["config"],
                   "inbound_nodes": [[[s, 0, 0, {}] for s in self.by_name[n]["inbound"]]] if self.by_name[n]["inbound"] else []}
                  for n in order]
        return {"class_name": "Model", "config": {"name": self.name, "layers": layers,
                                                  "input_layers": [[n, 0, 0] for n in inputs],
                                                  "output_layers": [[n, 0, 0] for n in outputs]}}


def generator_graph(input_shape=(256, 256, 1), noise_len=32, nc_out=1, first_fm=32, variant=None, flatten_index=None):
    """Gen_UNet2D as a graph.  variant 'gan' (TG:349-498: Dropout after every in / noise conv, tanh head) or 'uresnet'
    (TU:291-428: the single `do_gen_1`, softmax head); default by nc_out as the reference scripts use them.
    The one auto-named layer is the Flatten of the noise path: `flatten_3` in the DEP-GAN script's first fold (the two
    critics, built first at TG:513-516, take flatten_1 / flatten_2), `flatten_1` in the DEP-UResNet script."""
    variant = variant or ("gan" if nc_out == 1 else "uresnet")
    flatten_index = flatten_index or (3 if variant == "gan" else 1)
    f = first_fm
    g = _Graph("Gen_UNet2D")
    x = g.input("input_gen_chn_0", input_shape)
    z = g.input("input_gen_noiseZ_0", (noise_len, 1))
    h = z
    for tag in ("noise_1_add_f0", "noise_1_add_f1"):
        h = g.dense("dense_" + tag, h, f)
        h = g.bn("dense_bn_" + tag, h)
        h = g.act("dense_relu_" + tag, h, "relu")
    flat = g.flatten("flatten_%d" % flatten_index, h)
    mult = {suf: m for _, _, _, suf, _, m, _ in _BLOCKS}
    film = {}
    for suf in _FILM_ORDER:
        for kind in ("add", "mul"):
            d = g.dense("dense_noise_2_%s%s" % (kind, suf), flat, f * mult[suf])
            film[(kind, suf)] = g.bn("dense_bn_noise_2_%s%s" % (kind, suf), d)
    skips = []
    for bi, (c_in, c_noise, c_out, suf, tag, m, drops) in enumerate(_BLOCKS):
        w = f * m
        a = g.conv("conv2d_" + c_in, x, w, 3)
        a = g.bn("bn_" + c_in, a)
        a = g.act("relu_" + c_in, a, "relu")
        if variant == "gan":
            a = g.dropout(drops[0], a)
        elif c_in == "gen_10":
            a = g.dropout("do_gen_1", a)
        y = g.conv("conv2d_" + c_noise, a, w, 3)
        y = g.bn("bn_" + c_noise, y)
        if variant == "gan":
            y = g.dropout(drops[1], y)
        y = g.merge("Multiply", "mul_noiseZ_" + tag, [y, film[("mul", suf)]])
        y = g.merge("Add", "add_noiseZ_" + tag, [y, film[("add", suf)]])
        y = g.act("relu_noise_" + tag, y, "relu")
        r = g.merge("Add", "add_noiseZres_" + tag, [y, a])
        o = g.conv("conv2d_" + c_out, r, w, 3)
        o = g.bn("bn_" + c_out, o)
        o = g.act("relu_" + c_out, o, "relu")
        if bi < 3:
            skips.append(o)
            x = g.pool("maxpool2d_gen_%d" % bi, o)
        elif bi in _DECONV:
            dname, cname, skip = _DECONV[bi]
            u = g.deconv("deconv2d_" + dname, o, w, 2)
            u = g.bn("bn_" + dname, u)
            u = g.act("relu_" + dname, u, "relu")
            x = g.merge("Concatenate", cname, [u, skips[skip]], axis=-1)
        else:
            x = o
    seg = g.conv("gen_segmentation", x, nc_out, 1)
    out = g.act("non_lin_segment", seg, "tanh" if variant == "gan" else "softmax")
    return g, ["input_gen_chn_0", "input_gen_noiseZ_0"], [out]


def critic_graph(input_shape=(256, 256, 1), dense_index=1):
    """Dis_C2D_FCN1 (TG:316-345).  The final Flatten / Dense are auto-named by Keras (`flatten_<k>`, `dense_<k>`)."""
    g = _Graph("Dis_C2D_FCN1")
    x = g.input("input_dis", input_shape)
    plan = [("dis_0a", 5, 16), ("dis_0b", 5, 16), "maxpool2d_dis_0", ("dis_1a", 5, 32), ("dis_1b", 5, 32), "maxpool2d_dis_1",
            ("dis_2", 3, 64), ("dis_3", 3, 64), "maxpool2d_dis_2", ("dis_4", 3, 128), ("dis_5", 3, 128), "maxpool2d_dis_3",
            ("dis_6", 3, 256), ("dis_7", 3, 256), ("dis_8", 3, 256)]
    for item in plan:
        if isinstance(item, str):
            x = g.pool(item, x)
        else:
            name, k, c = item
            x = g.conv("conv2d_" + name, x, c, k)
            x = g.act("relu_" + name, x, "relu")
    x = g.conv("dis_9", x, 1, 1, init=_HE)
    x = g.flatten("flatten_%d" % dense_index, x)
    out = g.dense("dense_%d" % dense_index, x, 1)
    return g, ["input_dis"], [out]


def describe(model, input_shape, noise_len=32, nc_out=1):
    """-> dict(model_config=JSON text, layer_names=[...Keras order...], weights={layer: [weight names]}) for
    model 'generator' / 'critic'."""
    if model == "generator":
        g, ins, outs = generator_graph(input_shape, noise_len, nc_out)
    elif model == "critic":
        g, ins, outs = critic_graph(input_shape)
    else:
        raise ValueError("model must be 'generator' or 'critic'")
    cfg = g.model_config(ins, outs)
    names = [l["name"] for l in cfg["config"]["layers"]]
    return {"model_config": json.dumps(cfg), "layer_names": names,
            "weights": {n: g.by_name[n]["weights"] for n in names}}


def adam_training_config(lr=1e-4, beta_1=0.9, beta_2=0.999, loss="categorical_crossentropy"):
    """`training_config` of a model compiled as TU:427 (`Adam(lr=1e-4)`, categorical cross-entropy, no metrics)."""
    return json.dumps({"optimizer_config": {"class_name": "Adam",
                                            "config": {"lr": lr, "beta_1": beta_1, "beta_2": beta_2, "decay": 0.0,
                                                       "epsilon": 1e-07, "amsgrad": False}},
                       "loss": loss, "metrics": [], "sample_weight_mode": None, "loss_weights": None})
