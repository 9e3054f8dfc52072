"""Host-only checks of the tcgen05 convolution planner (csrc/conv_tc.cu::plan) through depgan_op_conv_plan: the
shared-memory / TMEM / ring invariants the kernel relies on, for every layer shape of the two networks and a sweep of
other shapes.  No GPU is needed: the planner is arithmetic."""
import ctypes as C
import itertools
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from depgan_b200 import _lib  # noqa: E402

FIELDS = ["kc", "ncta", "nchunks", "na", "nb", "b_tps", "b_resident", "acc_stages", "n_issuers", "ch", "n_side",
          "tmem_cols", "smem_bytes", "pool", "stage_out", "nsplit"]
FAKE = 0x10000  # the planner only tests pointers for NULL


def plan(ks, c0, c1, cout, H=256, W=256, N=4, *, film=False, add=False, mask=False, deconv=False, pool=False, head=0,
         pre=False, out=True):
    d = _lib.ConvDesc()
    d.in0, d.C0, d.C1 = FAKE, c0, c1
    if c1:
        d.in1 = FAKE
    d.w_bf16 = FAKE
    d.scale = d.shift = FAKE
    if out:
        d.out = FAKE
    if pre:
        d.out_pre = FAKE
    if film:
        d.film_g = d.film_b = d.res = FAKE
        d.film_stride = cout
    if add:
        d.add_src = FAKE
    if mask:
        d.mask_src = FAKE
    if pool:
        d.pool_out = FAKE
    if head:
        d.head_w = d.head_b = d.head_out = FAKE
        d.head_nc, d.head_act = head, 1
    d.relu, d.deconv = 1, int(deconv)
    d.N, d.H, d.W, d.Cout, d.ks = N, H, W, cout, ks
    d.in_bf16 = d.out_bf16 = d.use_tc = 1
    out16 = (C.c_int * 16)()
    rc = _lib.lib().depgan_op_conv_plan(C.byref(d), out16)
    assert rc in (0, 1), _lib.last_error()
    return dict(zip(FIELDS, out16)) if rc == 1 else None


def check_invariants(p, ks, c0, c1, cout, deconv):
    ncols = 4 * cout if deconv else cout
    kc = p["kc"]
    assert kc in (16, 32, 64)
    if c0 % kc == 0 and c1 % kc == 0:
        assert p["nchunks"] == (c0 + c1) // kc
    else:
        # padded 64-channel chunks (TMA zero-fills the channels past a source's end): only for sources of >= 64 channels
        # and while the padded K stays within 25 % of the real one
        assert kc == 64 and c0 >= 64 and (c1 == 0 or c1 >= 64)
        assert p["nchunks"] == -(-c0 // kc) - (-c1 // kc) and 4 * p["nchunks"] * kc <= 5 * (c0 + c1)
    assert p["ncta"] % 16 == 0 and 16 <= p["ncta"] <= 256 and p["ncta"] * p["nsplit"] == ncols
    assert p["ncta"] % p["ch"] == 0 and p["ch"] in (16, 32, 64)
    # TMEM: two strips of ncta columns per accumulator stage, power of two, at most the SM's 512 columns
    t = p["tmem_cols"]
    assert t & (t - 1) == 0 and 32 <= t <= 512 and t >= 2 * p["ncta"] * p["acc_stages"]
    assert p["acc_stages"] in (1, 2)
    # shared memory: the 227 KB opt-in limit of one CTA, with the alignment slack the kernel takes
    assert p["smem_bytes"] <= 227 * 1024
    assert p["na"] >= 2 and p["nb"] >= 1
    if p["b_resident"]:
        assert p["nb"] == ks * ks * p["nchunks"] and p["nsplit"] == 1
    else:
        assert p["b_tps"] in (ks * ks, ks, 1) and p["nb"] >= 2
    # two MMA issuers share the A ring through parity waits: a stage may be refilled at most once per two items
    assert p["n_issuers"] in (1, 2)
    if p["n_issuers"] == 2:
        assert p["b_resident"] and p["acc_stages"] == 2 and p["na"] >= 2 * p["nchunks"]


# every tcgen05 layer of the generator (TG:398-491) and of the critic (TG:316-345) at 256x256
GEN_W = [32, 64, 96, 128, 96, 64, 32]
GEN_LVL = [0, 1, 2, 3, 2, 1, 0]


def network_layers():
    L = []
    for bi, (w, lvl) in enumerate(zip(GEN_W, GEN_LVL)):
        hw = 256 >> lvl
        if 1 <= bi <= 3:
            L.append(("gen_in%d" % bi, dict(ks=3, c0=GEN_W[bi - 1], c1=0, cout=w, H=hw, W=hw)))
        elif bi >= 4:
            L.append(("gen_in%d" % bi, dict(ks=3, c0=w, c1=GEN_W[6 - bi], cout=w, H=hw, W=hw)))
        L.append(("gen_noise%d" % bi, dict(ks=3, c0=w, c1=0, cout=w, H=hw, W=hw, film=True)))
        L.append(("gen_noise%d_train" % bi, dict(ks=3, c0=w, c1=0, cout=w, H=hw, W=hw, film=True, pre=True)))
        L.append(("gen_out%d" % bi, dict(ks=3, c0=w, c1=0, cout=w, H=hw, W=hw, pool=bi < 3)))
        if 3 <= bi < 6:
            L.append(("gen_dec%d" % bi, dict(ks=1, c0=w, c1=0, cout=GEN_W[bi + 1], H=hw, W=hw, deconv=True)))
    L.append(("gen_head_tanh", dict(ks=3, c0=32, c1=0, cout=32, head=1, out=False)))
    L.append(("gen_head_softmax", dict(ks=3, c0=32, c1=0, cout=32, head=4, out=False)))
    crit = [(5, 16, 16, 0), (5, 16, 32, 1), (5, 32, 32, 1), (3, 32, 64, 2), (3, 64, 64, 2), (3, 64, 128, 3),
            (3, 128, 128, 3), (3, 128, 256, 4), (3, 256, 256, 4)]
    for i, (ks, cin, cout, lvl) in enumerate(crit):
        hw = 256 >> lvl
        L.append(("critic%d" % i, dict(ks=ks, c0=cin, c1=0, cout=cout, H=hw, W=hw)))
        L.append(("critic%d_jvp" % i, dict(ks=ks, c0=cin, c1=0, cout=cout, H=hw, W=hw, mask=True)))
        L.append(("critic%d_dgrad" % i, dict(ks=ks, c0=cout, c1=0, cout=cin, H=hw, W=hw, add=True, mask=True)))
    return L


@pytest.mark.parametrize("name,kw", network_layers(), ids=[n for n, _ in network_layers()])
def test_every_network_layer_has_a_valid_tcgen05_plan(name, kw):
    kw = dict(kw)
    ks, c0, c1, cout = kw.pop("ks"), kw.pop("c0"), kw.pop("c1"), kw.pop("cout")
    p = plan(ks, c0, c1, cout, **kw)
    assert p is not None, "%s falls off the tensor-core path" % name
    check_invariants(p, ks, c0, c1, cout, kw.get("deconv", False))


def test_dominant_layers_keep_their_weights_resident_with_two_issuers():
    for c in (32, 64):  # the 256x256 and 128x128 levels: resident weights, both issuers
        p = plan(3, c, 0, c, H=256 * 32 // c, W=256 * 32 // c)
        assert p["b_resident"] == 1 and p["n_issuers"] == 2 and p["acc_stages"] == 2, p
    # the shape that dead-locked the first two-issuer build (na = nchunks = 3): must fall back to one issuer
    p = plan(3, 96, 0, 64, H=64, W=64)
    assert p["n_issuers"] == 1 or p["na"] >= 2 * p["nchunks"], p


def test_planner_sweep_invariants_and_clean_rejections():
    chans = [16, 32, 48, 64, 96, 128, 160, 192, 256]
    n_ok = 0
    for ks, c0, c1, cout in itertools.product((1, 3, 5), chans, (0, 32, 64), chans + [320, 512]):
        for kind in ("plain", "film", "addmask", "pool", "head"):
            if kind == "film" and (c1 or cout != c0):
                continue
            kw = dict(film=kind == "film", add=kind == "addmask", mask=kind == "addmask", pool=kind == "pool",
                      head=4 if kind == "head" else 0)
            p = plan(ks, c0, c1, cout, H=32, W=32, **kw)
            if p is None:
                continue
            n_ok += 1
            check_invariants(p, ks, c0, c1, cout, False)
            assert (kind != "film") or ks == 3          # FiLM epilogue exists for 3x3 only
            assert (kind != "pool") or ks != 1
    assert n_ok > 1000
    # shapes outside the path are refused, not mis-planned
    assert plan(3, 8, 0, 32) is None and plan(3, 32, 0, 24) is None and plan(7, 32, 0, 32) is None
    assert plan(3, 32, 0, 32, H=40, W=64) is None
