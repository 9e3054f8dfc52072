"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/depgan_b200.h
declares, its weight manifest equals the reference's layer/weight naming, and it fails
loudly without a GPU (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest
import torch

from depgan_b200 import _lib
from depgan_b200.api import manifest
from oracle import depgan_oracle as O


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.header_functions()
    assert len(declared) >= 25
    for fn in declared:
        assert hasattr(L, fn), fn
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert L.depgan_abi_version() == 1


@pytest.mark.parametrize("nicg,nc_out,total", [(1, 1, 2491969), (2, 1, 2492257), (1, 4, 2492068)])
def test_generator_manifest_matches_reference_naming(nicg, nc_out, total):
    cfg = _lib.Cfg(256, 256, nicg, nc_out, 32, 4, _lib.PREC_BF16, 0)
    man, nfl = manifest(_lib.MODEL_GEN, cfg)
    want = O.gen_manifest(nicg, nc_out)
    assert [(n, tuple(s)) for n, s, _, _ in man] == [("%s/%s" % (l, w), tuple(s)) for l, w, s in want]
    assert sum(int(np.prod(s)) for _, s, _, _ in man) == total == O.manifest_count(want)
    assert nfl >= total
    for n, s, off, tr in man:
        assert off % 4 == 0
        assert tr == O.is_trainable(n.split("/")[1])
    offs = sorted((off, int(np.prod(s))) for _, s, off, _ in man)
    assert all(a + c <= b for (a, c), (b, _) in zip(offs, offs[1:]))  # no overlap


def test_critic_manifest_matches_reference_naming():
    cfg = _lib.Cfg(256, 256, 1, 1, 32, 4, _lib.PREC_BF16, 0)
    man, _ = manifest(_lib.MODEL_CRITIC, cfg)
    want = O.critic_manifest(256, 256)
    assert [(n, tuple(s)) for n, s, _, _ in man] == [("%s/%s" % (l, w), tuple(s)) for l, w, s in want]
    assert sum(int(np.prod(s)) for _, s, _, _ in man) == 1798002


def test_bad_configs_are_rejected_with_messages():
    L = _lib.lib()
    bad = _lib.Cfg(250, 256, 1, 1, 32, 4, 1, 0)
    assert L.depgan_manifest_count(0, C.byref(bad)) < 0
    assert b"multiples of 16" in L.depgan_last_error()
    bad = _lib.Cfg(256, 256, 1, 7, 32, 4, 1, 0)
    assert L.depgan_workspace_bytes(0, C.byref(bad)) < 0
    assert L.depgan_manifest_count(5, C.byref(_lib.Cfg(256, 256, 1, 1, 32, 4, 1, 0))) < 0


def test_workspace_grows_with_batch_and_training():
    L = _lib.lib()
    a = L.depgan_workspace_bytes(0, C.byref(_lib.Cfg(256, 256, 1, 1, 32, 4, 1, 0)))
    b = L.depgan_workspace_bytes(0, C.byref(_lib.Cfg(256, 256, 1, 1, 32, 8, 1, 0)))
    c = L.depgan_workspace_bytes(0, C.byref(_lib.Cfg(256, 256, 1, 1, 32, 8, 1, 1)))
    f = L.depgan_workspace_bytes(0, C.byref(_lib.Cfg(256, 256, 1, 1, 32, 8, 0, 0)))
    assert 0 < a < b < c and f > b


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from depgan_b200 import Gen_UNet2D
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Gen_UNet2D((256, 256, 1), (32, 1), 32, 1)
    L = _lib.lib()
    cfg = _lib.Cfg(64, 64, 1, 1, 32, 1, 1, 0)
    buf = (C.c_float * 16)()
    h = L.depgan_net_create(0, C.byref(cfg), C.addressof(buf), None, C.addressof(buf), 64)
    assert not h and b"no CUDA device" in L.depgan_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the shipped package may import or execute it."""
    import pathlib
    import re
    pkg = pathlib.Path(__file__).resolve().parent.parent / "dep-gan-im_b200"
    offenders = [p.name for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*"))
                 if p.is_file() and re.search(r"\boracle\b", p.read_text(errors="ignore"))]
    assert not offenders, offenders
