"""The C ABI is usable from plain C with no Python in the loop: examples/c_caller.c builds against
include/depgan_b200.h + the in-tree library, and on a B200 reproduces the Python surface bit for bit."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _binary():
    import __graft_entry__ as ge
    from depgan_b200 import build as _build
    return ge.build_c_caller(_build.build())


def test_c_caller_builds_and_fails_loudly_without_inputs(tmp_path):
    exe = _binary()
    assert exe.exists()
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    # a missing parameter file is an error, not a silent zero-weight run
    r = subprocess.run([str(exe), str(tmp_path / "none.bin"), "x", "z", "o", "1", "1", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("nicg,nc_out,precision", [(1, 1, "bf16"), (1, 4, "bf16"), (2, 1, "fp32"), (1, 4, "f16")])
def test_c_caller_matches_python_surface(tmp_path, nicg, nc_out, precision):
    import torch
    from depgan_b200 import Gen_UNet2D, synth
    H = W = 64
    n = 3
    g = Gen_UNet2D((H, W, nicg), (32, 1), 32, nc_out, precision=precision, max_batch=n, seed=7)
    man3 = [(nm.split("/")[0], nm.split("/")[1], s) for nm, s, _, _ in g.manifest]
    g.set_weights(synth.init_weights(man3, seed=3, trained_like=True))
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (n, H, W, nicg)).astype(np.float32)
    z = synth.make_noise(n, seed=6)
    want = g.predict([x, z])
    g.params.cpu().numpy().tofile(tmp_path / "params.bin")
    x.tofile(tmp_path / "x.bin")
    z.astype(np.float32).tofile(tmp_path / "z.bin")
    exe = _binary()
    r = subprocess.run([str(exe), str(tmp_path / "params.bin"), str(tmp_path / "x.bin"), str(tmp_path / "z.bin"),
                        str(tmp_path / "out.bin"), str(n), str(nicg), str(nc_out), precision, str(H), str(W)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "kernel launches" in r.stdout
    got = np.fromfile(tmp_path / "out.bin", dtype=np.float32).reshape(want.shape)
    assert np.array_equal(got, want)  # same kernels, same flat parameters: identical bits
