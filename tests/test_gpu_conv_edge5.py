"""Parity of the kernels for the critics' one-channel edge layer (conv2d_dis_0a, TG:319, 1 -> 16, 5x5 on the fp32 image;
conv_first_tc.cu / conv_simt.cu): forward (+ ReLU), JVP (activation mask), weight gradient and the 16 -> 1 data gradient
of a bf16 critic, against fp64 references, through the kernel-level C ABI, at widths / heights that are not multiples of
the kernels' tiles."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from depgan_b200 import _lib

pytestmark = pytest.mark.gpu
GEOMS = [(2, 32, 128), (1, 21, 256), (3, 8, 132), (2, 5, 4), (1, 64, 384)]  # N, H, W (W a multiple of 4)


def _bf(t):
    return t.to(torch.bfloat16).float()


def _conv_desc(x, w, out, *, scale=None, shift=None, mask=None, relu=False, in_bf16=False, out_bf16=True, keep=None):
    d = _lib.ConvDesc()
    N, H, W, c0 = x.shape
    d.in0, d.C0, d.C1 = x.data_ptr(), c0, 0
    d.w_f32 = w.data_ptr()
    if scale is not None:
        d.scale = scale.data_ptr()
    if shift is not None:
        d.shift = shift.data_ptr()
    if mask is not None:
        d.mask_src = mask.data_ptr()
    d.out = out.data_ptr()
    d.relu = int(relu)
    d.N, d.H, d.W, d.Cout, d.ks = N, H, W, out.shape[3], 5
    d.in_bf16, d.out_bf16, d.use_tc = int(in_bf16), int(out_bf16), 0
    return d


def _run(d):
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.depgan_op_conv2d(C.byref(d), st), "op_conv2d")
    torch.cuda.synchronize()


@pytest.mark.parametrize("N,H,W", GEOMS)
@pytest.mark.parametrize("mode", ["relu", "mask"])
def test_first_layer_forward_and_jvp(N, H, W, mode):
    g = torch.Generator().manual_seed(N * 1000 + H + W)
    x = torch.randn((N, H, W, 1), generator=g)
    w = torch.randn((5, 5, 1, 16), generator=g) * 0.2
    b = torch.randn((16,), generator=g) * 0.1
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(3, 2, 0, 1).double(), padding=2).permute(0, 2, 3, 1)
    xd, wd = x.cuda().contiguous(), w.reshape(25, 1, 16).cuda().contiguous()
    out = torch.empty((N, H, W, 16), dtype=torch.bfloat16, device="cuda")
    if mode == "relu":
        bd = b.cuda()
        _run(_conv_desc(xd, wd, out, shift=bd, relu=True))
        want = torch.relu(ref + b.double())
    else:
        m = _bf(torch.randn((N, H, W, 16), generator=g))
        md = m.cuda().to(torch.bfloat16).contiguous()
        _run(_conv_desc(xd, wd, out, mask=md))
        want = torch.where(m > 0, ref, torch.zeros_like(ref))
    got = out.float().cpu().double()
    tol = 2.0 ** -8 * max(1.0, float(want.abs().max()))  # one bf16 rounding of the output
    assert float((got - want).abs().max()) <= tol


@pytest.mark.parametrize("N,H,W", GEOMS)
def test_first_layer_weight_gradient(N, H, W):
    g = torch.Generator().manual_seed(7 + N + H + W)
    x = torch.randn((N, H, W, 1), generator=g)
    dy = _bf(torch.randn((N, H, W, 16), generator=g))
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xd, dyd = x.cuda().contiguous(), dy.cuda().to(torch.bfloat16).contiguous()
    dw = torch.zeros((5, 5, 1, 16), dtype=torch.float32, device="cuda")
    _lib.check(L.depgan_op_wgrad(xd.data_ptr(), None, 1, 0, dyd.data_ptr(), dw.data_ptr(), N, H, W, 16, 5, 2, st), "op_wgrad")
    torch.cuda.synchronize()
    xp = F.pad(x[..., 0].double(), (2, 2, 2, 2))
    want = torch.stack([torch.stack([torch.einsum("nhw,nhwc->c", xp[:, a:a + H, b:b + W], dy.double()) for b in range(5)])
                        for a in range(5)])  # [dy][dx][co]
    got = dw.cpu().double()[:, :, 0, :]
    assert float((got - want).abs().max()) <= 2e-4 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("N,H,W", GEOMS)
def test_last_layer_data_gradient(N, H, W):
    """16 -> 1, 5x5, bf16 in, fp32 out, no epilogue: the data gradient of conv2d_dis_0a as the training graphs call it
    (the caller passes the flipped taps; here any 5x5x16x1 kernel)."""
    g = torch.Generator().manual_seed(11 + N + H + W)
    x = _bf(torch.randn((N, H, W, 16), generator=g))
    w = torch.randn((5, 5, 16, 1), generator=g) * 0.1
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(3, 2, 0, 1).double(), padding=2).permute(0, 2, 3, 1)
    xd = x.cuda().to(torch.bfloat16).contiguous()
    wd = w.reshape(25, 16, 1).cuda().contiguous()
    out = torch.empty((N, H, W, 1), dtype=torch.float32, device="cuda")
    _run(_conv_desc(xd, wd, out, in_bf16=True, out_bf16=False))
    got = out.cpu().double()
    assert float((got - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
