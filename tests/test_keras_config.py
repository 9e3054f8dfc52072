"""Keras-side description of the saved models (TG:892 `netG.save`, TU:622-623 `save` + `to_json`): layer graph, Keras'
layer order, full-model HDF5 layout.  No Keras exists here, so these are structural checks against the native manifest,
the oracle's manifest and the reference's layer names."""
import json
import re

import numpy as np
import pytest

from depgan_b200 import h5lite, keras_config as kc, synth
from oracle import depgan_oracle as O


@pytest.mark.parametrize("nicg,nc_out", [(1, 1), (2, 1), (1, 4)])
def test_generator_description_matches_the_manifest(nicg, nc_out):
    d = kc.describe("generator", (256, 256, nicg), 32, nc_out)
    man = O.gen_manifest(nicg, nc_out)
    want = {}
    for layer, w, shape in man:
        want.setdefault(layer, []).append(w)
    have = {n: ws for n, ws in d["weights"].items() if ws}
    assert have == want                                   # same weight-bearing layers, same per-layer tensor order
    assert len(have) == 81
    cfg = json.loads(d["model_config"])["config"]
    assert cfg["name"] == "Gen_UNet2D" and [l["name"] for l in cfg["layers"]] == d["layer_names"]
    assert cfg["input_layers"] == [["input_gen_chn_0", 0, 0], ["input_gen_noiseZ_0", 0, 0]]
    assert cfg["output_layers"] == [["non_lin_segment", 0, 0]]
    by = {l["name"]: l for l in cfg["layers"]}
    assert by["input_gen_chn_0"]["config"]["batch_input_shape"] == [None, 256, 256, nicg]
    assert by["non_lin_segment"]["config"]["activation"] == ("tanh" if nc_out == 1 else "softmax")
    # the decoder concatenations take [deconv output, skip] in that order (TG:450, 465, 479)
    assert [i[0] for i in by["concat_gen_0"]["inbound_nodes"][0]] == ["relu_de_gen_9", "relu_gen_5"]
    assert [i[0] for i in by["concat_gen_3"]["inbound_nodes"][0]] == ["relu_de_gen_15", "relu_gen_1"]
    # FiLM: multiply by the *mul* head, then add the *add* head, ReLU, then the residual add (TG:403-407)
    assert [i[0] for i in by["mul_noiseZ_p4"]["inbound_nodes"][0]][1] == "dense_bn_noise_2_mul"
    assert [i[0] for i in by["add_noiseZ_m2"]["inbound_nodes"][0]][1] == "dense_bn_noise_2_add_m2"
    drops = [n for n in d["layer_names"] if by[n]["class_name"] == "Dropout"]
    assert (len(drops), set(drops) >= {"do_gen_1"}) == ((14, False) if nc_out == 1 else (1, True))  # TG vs TU:388
    shapes = {(l, w): s for l, w, s in man}
    assert by["deconv2d_de_gen_9"]["config"]["strides"] == [2, 2] and shapes[("deconv2d_de_gen_9", "kernel")] == (2, 2, 128, 128)
    for (l, w), s in shapes.items():                      # filters / units of every layer agree with the tensor shapes
        if w == "kernel" and by[l]["class_name"] in ("Conv2D", "Dense"):
            assert by[l]["config"].get("filters", by[l]["config"].get("units")) == s[-1], l


def test_layer_order_is_a_valid_keras_depth_order():
    """Keras sorts model.layers by depth (longest path to the output), deepest first: every layer comes after all its
    inputs, inputs of equal depth keep the traversal order, and the ordering is stable against re-building."""
    g, ins, outs = kc.generator_graph((256, 256, 1), 32, 1)
    order = g.keras_order(outs)
    pos = {n: i for i, n in enumerate(order)}
    assert len(order) == len(g.layers) == len(set(order))
    for l in g.layers:
        for src in l["inbound"]:
            assert pos[src] < pos[l["name"]], (src, l["name"])
    assert order[-1] == "non_lin_segment" and order[-2] == "gen_segmentation"
    assert order == kc.generator_graph((256, 256, 1), 32, 1)[0].keras_order(outs)
    # the noise input is the deepest layer (its path runs through the whole FiLM MLP and all seven blocks)
    assert order[0] == "input_gen_noiseZ_0"


def test_critic_description():
    d = kc.describe("critic", (256, 256, 1))
    have = [n for n in d["layer_names"] if d["weights"][n]]
    assert have == [l for l, w, _ in O.critic_manifest(256, 256) if w == "kernel"]
    cfg = json.loads(d["model_config"])["config"]
    assert cfg["name"] == "Dis_C2D_FCN1" and cfg["layers"][0]["name"] == "input_dis"
    assert [l["class_name"] for l in cfg["layers"]].count("MaxPooling2D") == 4


def test_full_model_file_roundtrip(tmp_path):
    """save_keras_model -> our reader: Keras' layer order incl. weight-less layers, the JSON attributes as variable-length
    strings in a global heap, optimizer weights in nested groups with an int64 iteration counter."""
    d = kc.describe("generator", (64, 64, 1), 32, 4)
    man = O.gen_manifest(1, 4)
    P = synth.init_weights(man, seed=3, trained_like=True)
    opt = [("Adam/iterations:0", np.array(17, np.int64)),
           ("training/Adam/Variable:0", np.arange(6, dtype=np.float32).reshape(2, 3)),
           ("training/Adam/Variable_1:0", np.ones((1,), np.float32))]
    p = tmp_path / "full.h5"
    h5lite.save_keras_model(str(p), P, d["layer_names"], d["weights"], d["model_config"], kc.adam_training_config(), opt,
                            tf_scope_suffix="_2")
    layers, content = h5lite.read_keras_file(str(p))
    assert layers == d["layer_names"]
    assert [w for w, _ in content["bn_gen_4"]] == ["bn_gen_4_2/%s:0" % w for w in ("gamma", "beta", "moving_mean", "moving_variance")]
    got = h5lite.load_keras_weights(str(p), [("%s/%s" % (l, w), s) for l, w, s in man])
    assert all(np.array_equal(got[k], P[k]) for k in P)
    mc, tc = h5lite.read_keras_configs(str(p))
    assert mc == json.loads(d["model_config"]) and tc["optimizer_config"]["class_name"] == "Adam"
    f = h5lite.File(str(p))
    assert f.attrs()["keras_version"] == b"2.2.4" and f["model_weights"].attrs()["backend"] == b"tensorflow"
    assert int(f["optimizer_weights/Adam/iterations:0"].read()) == 17
    assert np.array_equal(f["optimizer_weights/training/Adam/Variable:0"].read(), opt[1][1])
    assert f["optimizer_weights"].attrs()["weight_names"].tolist() == [n.encode() for n, _ in opt]
    raw = p.read_bytes()
    assert raw.count(b"GCOL") >= 4 and len(d["model_config"]) > 60000  # larger than an object header may hold


def test_long_layer_name_lists_are_split_like_keras(tmp_path):
    names = ["layer_%04d_%s" % (i, "x" * 90) for i in range(800)]  # 800 x 101 bytes > 64 512
    p = tmp_path / "many.h5"
    h5lite.save_keras_model(str(p), {}, names, {}, None)
    f = h5lite.File(str(p))
    keys = sorted(k for k in f["model_weights"].attrs() if k.startswith("layer_names"))
    assert keys == ["layer_names0", "layer_names1"]
    assert h5lite._names_attr(f["model_weights"].attrs(), "layer_names") == names
