"""Kernel-level parity of the convolution kernels (through the C ABI, depgan_op_conv2d) against torch-CPU fp32
references of the same Keras ops (Conv2D 'same' TG:285-304, Conv2DTranspose k2s2 TG:307-312)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def ref_conv(x, w, x1=None, scale=None, shift=None, relu=False, film=None, res=None, add=None, mask=None):
    """x NHWC fp32 (cpu), w HWIO.  Mirrors the epilogue order documented in include/depgan_b200.h."""
    if x1 is not None:
        x = torch.cat([x, x1], dim=3)
    k = w.shape[0]
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(3, 2, 0, 1).double(), padding=k // 2).permute(0, 2, 3, 1)
    if scale is not None:
        y = y * scale.double()
    if shift is not None:
        y = y + shift.double()
    pre = y.clone()
    if film is not None:
        y = torch.relu(y * film[0][:, None, None, :].double() + film[1][:, None, None, :].double()) + res.double()
    if add is not None:
        y = y + add.double()
    if mask is not None:
        y = torch.where(mask > 0, y, torch.zeros_like(y))
    if relu:
        y = torch.relu(y)
    return y.float(), pre.float()


def ref_deconv(x, w, scale, shift):
    """Keras Conv2DTranspose k2 s2 valid, kernel (2,2,Cout,Cin), then affine + relu."""
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), w.permute(3, 2, 0, 1).double(), stride=2).permute(0, 2, 3, 1)
    return torch.relu(y * scale.double() + shift.double()).float()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


@pytest.mark.parametrize("ks,c0,c1,cout", [(3, 1, 0, 32), (3, 2, 0, 32), (3, 32, 0, 32), (5, 1, 0, 16), (5, 16, 0, 16),
                                           (1, 32, 0, 4), (3, 24, 8, 40), (3, 16, 0, 1), (5, 16, 0, 1)])
def test_simt_conv_matches_fp32_reference(ks, c0, c1, cout):
    from depgan_b200 import conv2d_op
    N, H, W = 2, 32, 48
    x = _rand((N, H, W, c0), 1)
    x1 = _rand((N, H, W, c1), 2) if c1 else None
    w = _rand((ks, ks, c0 + c1, cout), 3, 0.2)
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), x1=None if x1 is None else x1.cuda(), scale=sc, shift=sh, relu=True,
                    use_tc=False).cpu()
    want, _ = ref_conv(x, w, x1, sc, sh, relu=True)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), float((got - want).abs().max())


def test_first_and_last_layer_specialisations():
    """conv2d_dis_0a JVP (1 -> 16, 5x5, activation mask, no bias) and its data gradient (16 -> 1, fp32 out)."""
    from depgan_b200 import conv2d_op
    N, H, W = 2, 32, 48
    x, w = _rand((N, H, W, 1), 1), _rand((5, 5, 1, 16), 2, 0.2)
    mask = _rand((N, H, W, 16), 3)
    got = conv2d_op(x.cuda(), w.cuda(), mask=mask.cuda(), use_tc=False).cpu()
    want, _ = ref_conv(x, w, mask=mask)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)
    b, wd = _rand((N, H, W, 16), 4), _rand((5, 5, 16, 1), 5, 0.2)
    got = conv2d_op(b.cuda(), wd.cuda(), use_tc=False).cpu()
    want, _ = ref_conv(b, wd)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)


def test_simt_conv_film_mask_epilogues():
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 16, 32, 32
    x, w = _rand((N, H, W, c), 1), _rand((3, 3, c, c), 2, 0.1)
    g, b, res = 1 + 0.3 * _rand((N, c), 3), 0.2 * _rand((N, c), 4), _rand((N, H, W, c), 5)
    got, ex = conv2d_op(x.cuda(), w.cuda(), film=(g, b), res=res.cuda(), use_tc=False, want_pre=True)
    want, pre = ref_conv(x, w, film=(g, b), res=res)
    assert torch.allclose(got.cpu(), want, atol=2e-4, rtol=1e-4)
    assert torch.allclose(ex["pre"].cpu(), pre, atol=2e-4, rtol=1e-4)
    mask, add = _rand((N, H, W, c), 6), _rand((N, H, W, c), 7)
    got = conv2d_op(x.cuda(), w.cuda(), add=add.cuda(), mask=mask.cuda(), use_tc=False).cpu()
    want, _ = ref_conv(x, w, add=add, mask=mask)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)


TC_CASES = [  # ks, c0, c1, cout, H, W   (kc = 64 / 32 / 16 chunking, two sources, n-split)
    (3, 64, 0, 64, 32, 32), (3, 32, 0, 32, 32, 48), (3, 96, 0, 96, 16, 32), (3, 16, 0, 16, 32, 32),
    (5, 16, 0, 32, 32, 32), (5, 32, 0, 32, 16, 16), (3, 64, 32, 32, 32, 32), (3, 128, 96, 96, 16, 16),
    (3, 128, 0, 256, 16, 16), (3, 256, 0, 256, 16, 16), (1, 64, 0, 64, 16, 32), (3, 96, 64, 64, 32, 16),
]


@pytest.mark.parametrize("ks,c0,c1,cout,H,W", TC_CASES)
def test_tcgen05_conv_matches_reference(ks, c0, c1, cout, H, W):
    from depgan_b200 import conv2d_op
    N = 3
    x = _bf(_rand((N, H, W, c0), 1))
    x1 = _bf(_rand((N, H, W, c1), 2)) if c1 else None
    w = _bf(_rand((ks, ks, c0 + c1, cout), 3, 1.0 / np.sqrt(ks * ks * (c0 + c1))))
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), x1=None if x1 is None else x1.cuda(), scale=sc, shift=sh, relu=True,
                    use_tc=True).cpu()
    want, _ = ref_conv(x, w, x1, sc, sh, relu=True)
    err = float((got - want).abs().max())
    assert err <= 1e-2 * max(1.0, float(want.abs().max())), err  # bf16 output rounding: 2^-9 relative


def test_tcgen05_conv_film_residual_and_mask():
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 32, 32, 64
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.05))
    g, b, res = 1 + 0.3 * _rand((N, c), 3), 0.2 * _rand((N, c), 4), _bf(_rand((N, H, W, c), 5))
    got, ex = conv2d_op(x.cuda(), w.cuda(), film=(g, b), res=res.cuda(), use_tc=True, want_pre=True)
    want, pre = ref_conv(x, w, film=(g, b), res=res)
    assert float((got.cpu() - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))
    assert float((ex["pre"].cpu() - pre).abs().max()) <= 2e-2 * max(1.0, float(pre.abs().max()))
    mask, add = _bf(_rand((N, H, W, c), 6)), _bf(_rand((N, H, W, c), 7))
    got = conv2d_op(x.cuda(), w.cuda(), add=add.cuda(), mask=mask.cuda(), use_tc=True).cpu()
    want, _ = ref_conv(x, w, add=add, mask=mask)
    assert float((got - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("c", [64, 96, 128])
def test_tcgen05_transposed_conv(c):
    from depgan_b200 import conv2d_op
    N, H, W = 2, 16, 32
    x = _bf(_rand((N, H, W, c), 1))
    w = _bf(_rand((2, 2, c, c), 2, 1.0 / np.sqrt(c)))
    sc, sh = 1 + 0.1 * _rand((c,), 4), 0.1 * _rand((c,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True, deconv=True, use_tc=True).cpu()
    want = ref_deconv(x, w, sc, sh)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("nc,act", [(1, 0), (4, 1)])
def test_tcgen05_conv_fused_head(nc, act):
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 32, 32, 32
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.08))
    hw, hb = _rand((c, nc), 3, 0.3), 0.1 * _rand((nc,), 4)
    got, ex = conv2d_op(x.cuda(), w.cuda(), relu=True, head=(hw, hb, act), use_tc=True)
    y, _ = ref_conv(x, w, relu=True)
    seg = y.double() @ hw.double() + hb.double()
    want = torch.tanh(seg) if act == 0 else torch.softmax(seg, dim=-1)
    assert float((ex["head"].cpu().double() - want).abs().max()) <= 1e-2


def ref_wgrad(x, dy, ks):
    """dw[a,b,ci,co] = sum_{n,h,w} xpad[n,h+a,w+b,ci] * dy[n,h,w,co]  (fp64)."""
    p = ks // 2
    xp = F.pad(x.permute(0, 3, 1, 2).double(), (p, p, p, p))
    d = dy.double()
    H, W = x.shape[1], x.shape[2]
    out = torch.zeros(ks, ks, x.shape[3], dy.shape[3], dtype=torch.float64)
    for a in range(ks):
        for b in range(ks):
            out[a, b] = torch.einsum("nchw,nhwo->co", xp[:, :, a:a + H, b:b + W], d)
    return out.float()


WG_CASES = [  # ks, c0, c1, cout, H, W
    (3, 64, 0, 64, 32, 32), (3, 32, 0, 32, 32, 48), (5, 16, 0, 16, 32, 32), (5, 32, 0, 32, 32, 16),
    (5, 16, 0, 32, 16, 32), (3, 32, 0, 64, 16, 16), (3, 96, 0, 96, 16, 32), (3, 64, 32, 32, 32, 32),
    (3, 128, 96, 96, 16, 16), (3, 128, 0, 256, 16, 16), (3, 256, 0, 256, 16, 16), (1, 64, 0, 256, 16, 32),
    (1, 96, 0, 384, 16, 16), (3, 96, 64, 64, 32, 16),
]


@pytest.mark.parametrize("ks,c0,c1,cout,H,W", WG_CASES)
def test_tcgen05_wgrad_matches_reference(ks, c0, c1, cout, H, W):
    from depgan_b200 import wgrad_op
    N = 3
    x = _bf(_rand((N, H, W, c0), 1))
    x1 = _bf(_rand((N, H, W, c1), 2)) if c1 else None
    dy = _bf(_rand((N, H, W, cout), 3))
    got = wgrad_op(x.cuda(), dy.cuda(), ks, x1=None if x1 is None else x1.cuda(), use_tc=True).cpu()
    xa = x if x1 is None else torch.cat([x, x1], dim=3)
    want = ref_wgrad(xa, dy, ks)
    err = float((got - want).abs().max())
    assert err <= 2e-3 * float(want.abs().max()), (err, float(want.abs().max()))


@pytest.mark.parametrize("ks,c0,cout", [(3, 8, 16), (5, 1, 16), (3, 2, 32), (1, 24, 8)])
def test_simt_wgrad_matches_reference(ks, c0, cout):
    from depgan_b200 import wgrad_op
    x, dy = _rand((2, 32, 16, c0), 1), _rand((2, 32, 16, cout), 2)
    got = wgrad_op(x.cuda(), dy.cuda(), ks, use_tc=False).cpu()
    want = ref_wgrad(x, dy, ks)
    assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())


# ---- persistent-loop coverage: more work items than SMs, so the epilogue's side-input ring crosses item boundaries ----
@pytest.mark.parametrize("c,H,W,N", [(32, 128, 128, 3), (96, 64, 64, 11), (64, 64, 48, 13)])
def test_tcgen05_film_residual_many_items(c, H, W, N):
    from depgan_b200 import conv2d_op
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.06))
    g, b, res = 1 + 0.3 * _rand((N, c), 3), 0.2 * _rand((N, c), 4), _bf(_rand((N, H, W, c), 5))
    got, ex = conv2d_op(x.cuda(), w.cuda(), film=(g, b), res=res.cuda(), use_tc=True, want_pre=True)
    want, pre = ref_conv(x, w, film=(g, b), res=res)
    assert float((got.cpu() - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))
    assert float((ex["pre"].cpu() - pre).abs().max()) <= 2e-2 * max(1.0, float(pre.abs().max()))


@pytest.mark.parametrize("ks,cin,cout,H,W,N,use_add", [(3, 32, 32, 128, 128, 3, True), (5, 32, 16, 64, 64, 12, False),
                                                        (3, 128, 96, 32, 32, 40, True), (1, 64, 256, 32, 32, 40, True),
                                                        (3, 256, 128, 16, 16, 160, False)])
def test_tcgen05_add_mask_many_items(ks, cin, cout, H, W, N, use_add):
    """The dgrad / JVP epilogue (add_src, ReLU-pattern mask) over a persistent loop with 1..16 chunks per item."""
    from depgan_b200 import conv2d_op
    x = _bf(_rand((N, H, W, cin), 1))
    w = _bf(_rand((ks, ks, cin, cout), 2, 1.0 / np.sqrt(ks * ks * cin)))
    mask = _bf(_rand((N, H, W, cout), 6))
    add = _bf(_rand((N, H, W, cout), 7)) if use_add else None
    got = conv2d_op(x.cuda(), w.cuda(), add=None if add is None else add.cuda(), mask=mask.cuda(), use_tc=True).cpu()
    want, _ = ref_conv(x, w, add=add, mask=mask)
    assert float((got - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))


def test_tcgen05_plain_many_items_and_deconv_split():
    from depgan_b200 import conv2d_op
    N, H, W, c = 5, 64, 64, 64
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.05))
    sc, sh = 1 + 0.1 * _rand((c,), 4), 0.1 * _rand((c,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True, use_tc=True).cpu()
    want, _ = ref_conv(x, w, None, sc, sh, relu=True)
    assert float((got - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max()))
    wd = _bf(_rand((2, 2, c, c), 6, 1.0 / np.sqrt(c)))
    got = conv2d_op(x.cuda(), wd.cuda(), scale=sc, shift=sh, relu=True, deconv=True, use_tc=True).cpu()
    want = ref_deconv(x, wd, sc, sh)
    assert float((got - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max()))


# ---- edge-layer kernels: ragged strips (H not a multiple of the 32-row strip, odd H for the row pairs) ----
@pytest.mark.parametrize("ks,c0,cout,H,W", [(3, 1, 32, 80, 64), (3, 2, 32, 48, 40), (5, 1, 16, 35, 24)])
def test_first_layer_ragged(ks, c0, cout, H, W):
    from depgan_b200 import conv2d_op, wgrad_op
    N = 3
    x, w = _rand((N, H, W, c0), 1), _rand((ks, ks, c0, cout), 3, 0.2)
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True, use_tc=False).cpu()
    want, _ = ref_conv(x, w, None, sc, sh, relu=True)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), float((got - want).abs().max())
    dy = _rand((N, H, W, cout), 2)
    got = wgrad_op(x.cuda(), dy.cuda(), ks, use_tc=False).cpu()
    want = ref_wgrad(x, dy, ks)
    assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())


def test_last_layer_ragged():
    from depgan_b200 import conv2d_op
    b, wd = _rand((3, 35, 24, 16), 4), _rand((5, 5, 16, 1), 5, 0.2)
    got = conv2d_op(b.cuda(), wd.cuda(), use_tc=False).cpu()
    want, _ = ref_conv(b, wd)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)


# ---- ring sharing between the two MMA issuers: shapes whose activation ring holds one item (na == nchunks), many items
@pytest.mark.parametrize("cin,cout,H,W,N", [(96, 64, 64, 64, 16), (96, 96, 64, 64, 12), (128, 64, 32, 32, 48)])
def test_tcgen05_ring_one_item_many_items(cin, cout, H, W, N):
    from depgan_b200 import conv2d_op
    x = _bf(_rand((N, H, W, cin), 1))
    w = _bf(_rand((3, 3, cin, cout), 2, 1.0 / np.sqrt(9 * cin)))
    got = conv2d_op(x.cuda(), w.cuda(), use_tc=True).cpu()
    want, _ = ref_conv(x, w)
    assert float((got - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))


# ---- first layer on the tensor cores (conv_first_tc.cu): fp32 image split into two bf16 halves, bf16 weights ----
def _first_layer_bf16_out(x, w, sc, sh, relu, mask=None):
    """fp32 image + fp32 Keras weights -> bf16 activations through depgan_op_conv2d (the route the bf16 networks take
    for conv2d_gen_0 / conv2d_dis_0a)."""
    import ctypes as C
    from depgan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N, H, W, c0 = x.shape
    ks, cout = w.shape[0], w.shape[3]
    keep = [x.to(dev).contiguous(), w.reshape(ks * ks * c0, cout).to(dev).contiguous(), sc.to(dev), sh.to(dev)]
    out = torch.empty((N, H, W, cout), dtype=torch.bfloat16, device=dev)
    d = _lib.ConvDesc()
    d.in0, d.C0, d.C1 = keep[0].data_ptr(), c0, 0
    d.w_f32, d.scale, d.shift, d.out = keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(), out.data_ptr()
    if mask is not None:
        keep.append(mask.to(dev).to(torch.bfloat16).contiguous())
        d.mask_src = keep[-1].data_ptr()
    d.relu, d.deconv = int(relu), 0
    d.N, d.H, d.W, d.Cout, d.ks = N, H, W, cout, ks
    d.in_bf16, d.out_bf16, d.use_tc = 0, 1, 0
    _lib.check(L.depgan_op_conv2d(C.byref(d), st), "op_conv2d")
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("ks,c0,cout,H,W,N,use_mask", [(5, 1, 16, 64, 48, 7, False), (5, 1, 16, 32, 32, 40, True),
                                                       (3, 1, 32, 64, 64, 5, False), (3, 2, 32, 48, 80, 9, False)])
def test_first_layer_tensor_core(ks, c0, cout, H, W, N, use_mask):
    x = _rand((N, H, W, c0), 1)
    w = _rand((ks, ks, c0, cout), 3, 0.2)
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    mask = _rand((N, H, W, cout), 6) if use_mask else None
    got = _first_layer_bf16_out(x, w, sc, sh, relu=not use_mask, mask=mask)
    # the kernel keeps 16 mantissa bits of the image and rounds the weights to bf16 (like every layer of a bf16 net)
    want, _ = ref_conv(x, _bf(w), None, sc, sh, relu=not use_mask, mask=None if mask is None else _bf(mask))
    err = float((got - want).abs().max())
    assert err <= 8e-3 * max(1.0, float(want.abs().max())), err   # bf16 output rounding
    # and against exact weights: weight rounding only
    want32, _ = ref_conv(x, w, None, sc, sh, relu=not use_mask, mask=None if mask is None else _bf(mask))
    assert float((got - want32).abs().max()) <= 2e-2 * max(1.0, float(want32.abs().max()))


@pytest.mark.parametrize("N,H,W", [(3, 32, 48), (20, 64, 64), (5, 256, 256)])
def test_first_layer_wgrad_tensor_core(N, H, W):
    """conv2d_dis_0a weight gradient of a bf16 critic: fp32 image (split into two bf16 halves), bf16 gradient."""
    import ctypes as C
    from depgan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x, dy = _rand((N, H, W, 1), 1), _bf(_rand((N, H, W, 16), 2))
    xd, dyd = x.to(dev).contiguous(), dy.to(dev).to(torch.bfloat16).contiguous()
    dw = torch.zeros(25 * 16, device=dev)
    _lib.check(L.depgan_op_wgrad(xd.data_ptr(), None, 1, 0, dyd.data_ptr(), dw.data_ptr(), N, H, W, 16, 5, 2, st), "wgrad")
    torch.cuda.synchronize()
    want = ref_wgrad(x, dy, 5)
    got = dw.cpu().view(5, 5, 1, 16)
    assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max()) + 1e-4


@pytest.mark.parametrize("ks,cin,cout,H,W,N", [(5, 16, 16, 64, 64, 9), (3, 32, 32, 32, 48, 7), (3, 64, 64, 32, 32, 12),
                                               (3, 96, 96, 16, 16, 20), (5, 32, 16, 32, 32, 5)])
def test_tcgen05_wgrad_with_fused_channel_sums(ks, cin, cout, H, W, N):
    """The bias gradient sum_p dy[p][co] accumulated by the weight-gradient kernel's otherwise idle warps."""
    import ctypes as C
    from depgan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x, dy = _bf(_rand((N, H, W, cin), 1)), _bf(_rand((N, H, W, cout), 2))
    xd, dyd = x.to(dev).to(torch.bfloat16).contiguous(), dy.to(dev).to(torch.bfloat16).contiguous()
    dw = torch.zeros(ks * ks * cin * cout, device=dev)
    cs = torch.zeros(cout, device=dev)
    _lib.check(L.depgan_op_wgrad_csum(xd.data_ptr(), None, cin, 0, dyd.data_ptr(), dw.data_ptr(), cs.data_ptr(), N, H, W,
                                      cout, ks, st), "wgrad_csum")
    torch.cuda.synchronize()
    want = ref_wgrad(x, dy, ks)
    got = dw.cpu().view(ks, ks, cin, cout)
    assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())
    want_cs = dy.double().sum(dim=(0, 1, 2)).float()
    assert float((cs.cpu() - want_cs).abs().max()) <= 1e-4 * float(dy.abs().sum(dim=(0, 1, 2)).max())


@pytest.mark.parametrize("ks,cin,cout,H,W,N", [(3, 32, 32, 64, 64, 5), (3, 64, 64, 32, 48, 11), (3, 96, 96, 32, 32, 9),
                                               (5, 16, 16, 64, 64, 7), (5, 32, 32, 32, 32, 12), (3, 128, 128, 16, 16, 20)])
def test_tcgen05_fused_maxpool(ks, cin, cout, H, W, N):
    """MaxPooling2D(2) fused into the conv + BN + ReLU epilogue: bit-identical to pooling the stored bf16 output."""
    from depgan_b200 import conv2d_op
    x = _bf(_rand((N, H, W, cin), 1))
    w = _bf(_rand((ks, ks, cin, cout), 2, 1.0 / np.sqrt(ks * ks * cin)))
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    got, ex = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True, use_tc=True, want_pool=True)
    want, _ = ref_conv(x, w, None, sc, sh, relu=True)
    assert float((got.cpu() - want).abs().max()) <= 1e-2 * max(1.0, float(want.abs().max()))
    pooled = F.max_pool2d(got.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(ex["pool"], pooled)


def _random_tc_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        ks = int(rng.choice([1, 3, 3, 5]))
        c0 = int(rng.choice([16, 32, 48, 64, 96, 128]))
        c1 = int(rng.choice([0, 0, 16, 32, 64]))
        cout = int(rng.choice([16, 32, 48, 64, 96, 128, 160]))
        H, W = int(rng.choice([16, 32, 48])), int(rng.choice([16, 32, 64]))
        N = int(rng.integers(1, 24))
        epi = str(rng.choice(["plain", "plain", "film", "addmask", "mask", "pool"]))
        if epi == "film":  # the FiLM-residual epilogue exists for the 3x3 layers only (conv_tc_supported)
            ks, c1, cout = 3, 0, c0
        if epi == "pool" and ks == 1:
            ks = 3
        cases.append((ks, c0, c1, cout, H, W, N, epi))
    return cases


@pytest.mark.parametrize("case", _random_tc_cases(36, 1234), ids=lambda c: "-".join(map(str, c)))
def test_tcgen05_random_shapes(case):
    """Seeded random sweep over kernel size, channel counts (incl. 48 / 160 and two-source inputs), tile counts and
    epilogue variants: exercises every ring / issuer / staging plan the planner can produce for small tensors."""
    from depgan_b200 import conv2d_op
    ks, c0, c1, cout, H, W, N, epi = case
    x = _bf(_rand((N, H, W, c0), 1))
    x1 = _bf(_rand((N, H, W, c1), 2)) if c1 else None
    w = _bf(_rand((ks, ks, c0 + c1, cout), 3, 1.0 / np.sqrt(ks * ks * (c0 + c1))))
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    kw, rkw = {}, {}
    if epi == "film":
        g, b, res = 1 + 0.3 * _rand((N, cout), 6), 0.2 * _rand((N, cout), 7), _bf(_rand((N, H, W, cout), 8))
        kw = dict(film=(g, b), res=res.cuda())
        rkw = dict(film=(g, b), res=res)
    elif epi in ("addmask", "mask"):
        mask = _bf(_rand((N, H, W, cout), 9))
        add = _bf(_rand((N, H, W, cout), 10)) if epi == "addmask" else None
        kw = dict(mask=mask.cuda(), add=None if add is None else add.cuda())
        rkw = dict(mask=mask, add=add)
    relu = epi in ("plain", "pool")
    out = conv2d_op(x.cuda(), w.cuda(), x1=None if x1 is None else x1.cuda(), scale=sc, shift=sh, relu=relu, use_tc=True,
                    want_pool=(epi == "pool"), **kw)
    got = out[0] if isinstance(out, tuple) else out
    want, _ = ref_conv(x, w, x1, sc, sh, relu=relu, **rkw)
    assert float((got.cpu() - want).abs().max()) <= 2e-2 * max(1.0, float(want.abs().max()))
    if epi == "pool":
        pooled = F.max_pool2d(got.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        assert torch.equal(out[1]["pool"], pooled)


def _random_wg_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n):
        ks = int(rng.choice([1, 3, 3, 5]))
        c0 = int(rng.choice([16, 32, 48, 64, 96, 128, 256]))
        c1 = int(rng.choice([0, 0, 32, 64]))
        cout = int(rng.choice([16, 32, 64, 96, 128, 256]))
        cases.append((ks, c0, c1, cout, int(rng.choice([16, 32, 48])), int(rng.choice([16, 32, 64])), int(rng.integers(1, 12))))
    return cases


@pytest.mark.parametrize("case", _random_wg_cases(24, 4321), ids=lambda c: "-".join(map(str, c)))
def test_tcgen05_wgrad_random_shapes(case):
    """Seeded random sweep of the weight-gradient kernel's plans (tap stacking, channel groups, batch split).  A shape
    the tcgen05 path does not take must be refused by the library with an error, never computed wrongly."""
    from depgan_b200 import wgrad_op
    ks, c0, c1, cout, H, W, N = case
    x = _bf(_rand((N, H, W, c0), 1))
    x1 = _bf(_rand((N, H, W, c1), 2)) if c1 else None
    dy = _bf(_rand((N, H, W, cout), 3))
    try:
        got = wgrad_op(x.cuda(), dy.cuda(), ks, x1=None if x1 is None else x1.cuda(), use_tc=True).cpu()
    except RuntimeError as e:
        assert "not supported" in str(e) or "unsupported" in str(e), e
        got = wgrad_op(x.cuda(), dy.cuda(), ks, x1=None if x1 is None else x1.cuda(), use_tc=False).cpu()
    xa = x if x1 is None else torch.cat([x, x1], dim=3)
    want = ref_wgrad(xa, dy, ks)
    err = float((got - want).abs().max())
    assert err <= 2e-3 * float(want.abs().max()), (err, float(want.abs().max()))
