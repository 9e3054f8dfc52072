"""Regenerates the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

The reference (Keras/TF-1, Python 2) cannot be imported or run in this environment and ships no golden vectors
(the vectors made from its executed source are reference_vectors.npz, make_reference_vectors.py); these fixtures pin (a) the oracle restatement itself against silent drift,
(b) the byte layout of the HDF5 writer, (c) hand-derived known answers of the DEM post-processing.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from depgan_b200 import h5lite, synth  # noqa: E402
from oracle import depgan_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    H = W = 32
    out = {}
    for tag, nicg, nc_out, head in [("gan_im", 1, 1, "tanh"), ("gan_pf", 2, 1, "tanh"), ("uresnet", 1, 4, "softmax")]:
        P = synth.init_weights(O.gen_manifest(nicg, nc_out), seed=21, trained_like=True)
        x, y2, mask = synth.make_im_pair(2, H, W, nicg=nicg, seed=3)
        z = synth.make_noise(2, seed=4)
        with torch.no_grad():
            y = O.gen_forward(O.to_torch(P, torch.float64), torch.as_tensor(x, dtype=torch.float64),
                              torch.as_tensor(z, dtype=torch.float64), head).numpy()
        out[tag + "_out"] = y
    Pc = synth.init_weights(O.critic_manifest(H, W), seed=22, trained_like=True)
    _, y2, _ = synth.make_im_pair(2, H, W, seed=3)
    with torch.no_grad():
        out["critic_out"] = O.critic_forward(O.to_torch(Pc, torch.float64), torch.as_tensor(y2, dtype=torch.float64)).numpy()
    # one full critic loss + generator loss (fp64) on tiny inputs
    PG = synth.init_weights(O.gen_manifest(1, 1), seed=21, trained_like=True)
    x, y2, _ = synth.make_im_pair(2, H, W, seed=3)
    z, ep = synth.make_noise(2, seed=4), synth.make_eps(2, seed=5)
    tr = O.OracleTrainer(PG, Pc, synth.init_weights(O.critic_manifest(H, W), seed=23, trained_like=True), thr=0.178)
    out["critic_y2_losses"] = np.array(tr.netD_y2_train([y2, x, z, ep], update=False) + [tr.last_gp])
    out["critic_dem_losses"] = np.array(tr.netD_dem_train([y2, x, z, ep], update=False) + [tr.last_gp])
    out["gen_losses"] = np.array(tr.netG_no_update([x, y2, z]))
    np.savez_compressed(OUT / "oracle_vectors.npz", **out)
    # tiny Keras-layout HDF5 file, byte-level golden
    w = {"conv2d_a/kernel": np.arange(18, dtype=np.float32).reshape(3, 3, 1, 2), "conv2d_a/bias": np.array([1, 2], np.float32),
         "bn_a/gamma": np.array([0.5, 1.5], np.float32)}
    h5lite.save_keras_weights(str(OUT / "tiny_keras.h5"), w, list(w), extra_layers=["input_1"], tf_scope_suffix="_1")
    print("wrote", [p.name for p in OUT.iterdir()])


if __name__ == "__main__":
    main()
