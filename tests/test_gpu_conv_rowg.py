"""Parity of the generic row-streaming tcgen05 kernel (conv_rowg.cu) against fp64 F.conv2d references, through
depgan_op_conv2d: the 5x5 layers of Dis_C2D_FCN1 (TG:319-325: 16 -> 16, 16 -> 32, 32 -> 32 and the 32 -> 16 data gradient;
plain + ReLU, mask epilogue of the data-gradient / JVP passes, fused MaxPooling2D) and the 64-output-channel 3x3 layers
of Gen_UNet2D at half resolution (TG:411-421: 32 -> 64, 64 -> 64; plain, FiLM residual, add / mask).  Widths 128 / 256 /
384, odd heights, one-row images, many rows per CTA, and bit-identity where the arithmetic is exact (integer data)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_conv import _bf, _rand, ref_conv

pytestmark = pytest.mark.gpu


def _tol(want):
    return 1e-2 * max(1.0, float(want.abs().max()))


def _pool(t):  # NHWC 2x2 stride-2 max-pool
    return torch.nn.functional.max_pool2d(t.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)


SHAPES = [  # ks, cin, cout
    (5, 16, 16), (5, 16, 32), (5, 32, 32), (5, 32, 16), (3, 32, 64), (3, 64, 64),
]
GEOMS = [(2, 32, 256), (1, 16, 128), (3, 33, 128), (2, 7, 384), (5, 1, 128), (1, 2, 256), (2, 64, 128)]  # N, H, W


@pytest.mark.parametrize("ks,cin,cout", SHAPES)
@pytest.mark.parametrize("N,H,W", GEOMS)
def test_rowg_plain(ks, cin, cout, N, H, W):
    from depgan_b200 import conv2d_op
    x = _bf(_rand((N, H, W, cin), 1))
    w = _bf(_rand((ks, ks, cin, cout), 3, 1.0 / np.sqrt(ks * ks * cin)))
    sc, sh = 1 + 0.1 * _rand((cout,), 4), 0.1 * _rand((cout,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True).cpu()
    want, _ = ref_conv(x, w, None, sc, sh, relu=True)
    err = float((got - want).abs().max())
    assert err <= _tol(want), err


@pytest.mark.parametrize("ks,cin,cout", SHAPES)
def test_rowg_is_exact_on_integer_data(ks, cin, cout):
    """Small-integer activations and weights: every product and sum is exact in bf16 x bf16 -> fp32, so the result equals
    the reference bit for bit -- catches any tap / row / column / channel-chunk mix-up a tolerance could hide."""
    from depgan_b200 import conv2d_op
    N, H, W = 2, 21, 256
    g = torch.Generator().manual_seed(7)
    x = torch.randint(-2, 3, (N, H, W, cin), generator=g).float()
    w = torch.randint(-1, 2, (ks, ks, cin, cout), generator=g).float()
    got = conv2d_op(x.cuda(), w.cuda()).cpu()
    want, _ = ref_conv(x, w)
    assert float(want.abs().max()) <= 256 or torch.equal(_bf(want), want)
    assert torch.equal(got, _bf(want))


@pytest.mark.parametrize("ks,cin,cout", [(5, 16, 16), (5, 16, 32), (5, 32, 32), (5, 32, 16), (3, 32, 64), (3, 64, 64)])
def test_rowg_add_and_mask(ks, cin, cout):
    from depgan_b200 import conv2d_op
    N, H, W = 2, 48, 128  # (a multiple of 16: 64 channels with BOTH side inputs exceed this kernel's shared memory and
    x, w = _bf(_rand((N, H, W, cin), 1)), _bf(_rand((ks, ks, cin, cout), 2, 0.08))  # go to the tile kernel)
    add, mask = _bf(_rand((N, H, W, cout), 6)), _bf(_rand((N, H, W, cout), 7))
    for kw in ({"mask": mask}, {"add": add}, {"add": add, "mask": mask}):
        got = conv2d_op(x.cuda(), w.cuda(), **{k: v.cuda() for k, v in kw.items()}).cpu()
        want, _ = ref_conv(x, w, **kw)
        err = float((got - want).abs().max())
        assert err <= _tol(want), (list(kw), err)


@pytest.mark.parametrize("cin", [32, 64])
@pytest.mark.parametrize("N,H,W", [(2, 32, 128), (3, 17, 256)])
def test_rowg_film_residual(cin, N, H, W):
    from depgan_b200 import conv2d_op
    c = 64
    x, w = _bf(_rand((N, H, W, cin), 1)), _bf(_rand((3, 3, cin, c), 2, 0.06))
    sc, sh = 1 + 0.1 * _rand((c,), 6), 0.1 * _rand((c,), 7)
    g, b = 1 + 0.3 * _rand((N, c), 3), 0.2 * _rand((N, c), 4)
    res = _bf(_rand((N, H, W, c), 5))
    got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, film=(g, b), res=res.cuda()).cpu()
    want, _ = ref_conv(x, w, scale=sc, shift=sh, film=(g, b), res=res)
    err = float((got - want).abs().max())
    assert err <= _tol(want), err


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32)])
@pytest.mark.parametrize("N,H,W", [(2, 32, 256), (3, 6, 128), (1, 64, 128)])
def test_rowg_fused_maxpool(cin, cout, N, H, W):
    """conv + bias + ReLU + MaxPooling2D (TG:321, 325): the pooled tensor equals the pool of the stored tensor bit for
    bit (both are taken from the same rounded values)."""
    from depgan_b200 import conv2d_op
    x, w = _bf(_rand((N, H, W, cin), 1)), _bf(_rand((5, 5, cin, cout), 2, 0.05))
    sh = 0.1 * _rand((cout,), 5)
    got, ex = conv2d_op(x.cuda(), w.cuda(), shift=sh, relu=True, want_pool=True)
    want, _ = ref_conv(x, w, None, None, sh, relu=True)
    assert float((got.cpu() - want).abs().max()) <= _tol(want)
    assert torch.equal(ex["pool"].cpu(), _pool(got.cpu()))


def test_rowg_many_rows_per_cta_match_single_slice():
    """96 slices x 256 rows x 2 column blocks: every CTA walks ~330 rows across image borders, wrapping every ring many
    times; slices are independent, so slice k of the batch equals slice k computed alone (another split over CTAs)."""
    from depgan_b200 import conv2d_op
    H, W, c = 256, 256, 16
    x = _bf(_rand((96, H, W, c), 1))
    w = _bf(_rand((5, 5, c, c), 2, 0.05))
    full = conv2d_op(x.cuda(), w.cuda(), relu=True).cpu()
    for k in (0, 41, 95):
        one = conv2d_op(x[k:k + 1].cuda(), w.cuda(), relu=True).cpu()
        assert torch.equal(full[k:k + 1], one), k
    want, _ = ref_conv(x[:1], w, relu=True)
    assert float((full[:1] - want).abs().max()) <= _tol(want)
