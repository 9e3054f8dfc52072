"""HDF5-subset reader/writer used by load_weights / save (EG:383, EU:402, TG:892, TU:622)."""
import hashlib
import struct
from pathlib import Path

import numpy as np
import pytest

from depgan_b200 import h5lite, synth
from oracle import depgan_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def test_golden_file_reads_back():
    layers, content = h5lite.read_keras_file(str(GOLD / "tiny_keras.h5"))
    assert layers == ["conv2d_a", "bn_a", "input_1"]
    assert [w for w, _ in content["conv2d_a"]] == ["conv2d_a_1/kernel:0", "conv2d_a_1/bias:0"]
    assert content["input_1"] == []
    k = dict(content["conv2d_a"])["conv2d_a_1/kernel:0"]
    assert k.dtype == np.float32 and k.shape == (3, 3, 1, 2) and np.array_equal(k.ravel(), np.arange(18))
    got = h5lite.load_keras_weights(str(GOLD / "tiny_keras.h5"), [("conv2d_a/bias", (2,)), ("bn_a/gamma", (2,))])
    assert np.array_equal(got["conv2d_a/bias"], [1, 2]) and np.array_equal(got["bn_a/gamma"], [0.5, 1.5])


def test_writer_is_byte_stable(tmp_path):
    w = {"conv2d_a/kernel": np.arange(18, dtype=np.float32).reshape(3, 3, 1, 2),
         "conv2d_a/bias": np.array([1, 2], np.float32), "bn_a/gamma": np.array([0.5, 1.5], np.float32)}
    p = tmp_path / "t.h5"
    h5lite.save_keras_weights(str(p), w, list(w), extra_layers=["input_1"], tf_scope_suffix="_1")
    assert p.read_bytes() == (GOLD / "tiny_keras.h5").read_bytes()
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    (eof,) = struct.unpack_from("<Q", b, 40)
    assert eof == len(b)


@pytest.mark.parametrize("suffix", ["", "_3"])
def test_full_generator_roundtrip(tmp_path, suffix):
    man = O.gen_manifest(2, 1)
    P = synth.init_weights(man, seed=0, trained_like=True)
    names = ["%s/%s" % (l, w) for l, w, _ in man]
    p = tmp_path / "netG.h5"
    h5lite.save_keras_weights(str(p), P, names, extra_layers=["input_%d" % i for i in range(1, 60)],
                              tf_scope_suffix=suffix)
    got = h5lite.load_keras_weights(str(p), [(n, s) for n, (_, _, s) in zip(names, man)])
    assert set(got) == set(P) and all(np.array_equal(got[k], P[k]) for k in P)
    layers, _ = h5lite.read_keras_file(str(p))
    assert len(layers) == 81 + 59  # 81 weight-bearing layers + the weight-less extras


def test_shape_mismatch_and_missing_weights_raise(tmp_path):
    with pytest.raises(ValueError):
        h5lite.load_keras_weights(str(GOLD / "tiny_keras.h5"), [("conv2d_a/bias", (3,))])
    with pytest.raises(KeyError):
        h5lite.load_keras_weights(str(GOLD / "tiny_keras.h5"), [("conv2d_b/bias", (2,))])
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(bad))


def test_auto_named_critic_dense_alias(tmp_path):
    w = {"dis_9/kernel": np.ones((1, 1, 256, 1), np.float32), "dense_7/kernel": np.full((4, 1), 2, np.float32),
         "dense_7/bias": np.zeros((1,), np.float32)}
    p = tmp_path / "d.h5"
    h5lite.save_keras_weights(str(p), w, list(w))
    got = h5lite.load_keras_weights(str(p), [("dense_1/kernel", (4, 1)), ("dense_1/bias", (1,))])
    assert np.array_equal(got["dense_1/kernel"], w["dense_7/kernel"])


# ---- a fixture h5lite's writer cannot produce: hand-assembled from the HDF5 spec by tests/golden/make_h5py_like.py ----
def _h5py_like():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_h5py_like", GOLD / "make_h5py_like.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_h5py_like_fixture_is_reproducible():
    assert _h5py_like().build() == (GOLD / "h5py_like_keras.h5").read_bytes()


def test_reader_handles_what_libhdf5_writes():
    """Multi-level group B-tree, object-header continuation + NIL messages, layer_names0/1, compact datasets, TF-scope
    suffixes, variable-length strings in a global heap, fill-value / mtime messages, a trailing optimizer group."""
    mod = _h5py_like()
    layers, model_config, training_config, opt = mod.content()
    path = str(GOLD / "h5py_like_keras.h5")
    got_layers, content = h5lite.read_keras_file(path)
    assert got_layers == [n for n, _, _ in layers]                      # order of layer_names0 + layer_names1
    for name, scope, weights in layers:
        assert [w for w, _ in content[name]] == ["%s/%s:0" % (scope, w) for w, _ in weights]
        for (_, got), (_, want) in zip(content[name], weights):
            assert got.dtype == np.float32 and np.array_equal(got, want)
    wanted = [("%s/%s" % (n, w), a.shape) for n, _, ws in layers for w, a in ws if n != "dense_7"]
    wanted += [("dense_1/kernel", (6, 1)), ("dense_1/bias", (1,))]      # the auto-numbered Dense resolves by position
    loaded = h5lite.load_keras_weights(path, wanted)
    assert np.array_equal(loaded["conv2d_gen_2/kernel"], dict(layers[9][2])["kernel"])
    assert np.array_equal(loaded["dense_1/bias"], [0.25])
    mc, tc = h5lite.read_keras_configs(path)
    import json
    assert mc == json.loads(model_config) and tc == json.loads(training_config)
    f = h5lite.File(path)
    assert f["optimizer_weights"].attrs()["weight_names"].tolist() == [n.encode() for n, _ in opt]
    it = f["optimizer_weights/training/Adam/iterations:0"].read()
    assert it.dtype == np.int64 and it.shape == () and int(it) == 1234
    assert np.array_equal(f["optimizer_weights/training/Adam/Variable:0"].read(), opt[1][1])
    # the structure really is what the docstring claims (guards against the fixture silently degenerating)
    b = (GOLD / "h5py_like_keras.h5").read_bytes()
    levels = [b[i + 5] for i in range(0, len(b) - 8, 8) if b[i:i + 4] == b"TREE" and b[i + 4] == 0]
    assert max(levels) >= 1 and levels.count(0) >= 4
    assert b"GCOL" in b and b.count(b"SNOD") > 30
