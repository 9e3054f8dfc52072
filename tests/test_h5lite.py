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
