"""Parity of the row-streaming tcgen05 kernel (conv_row.cu: 3x3, 32 output channels, width a multiple of 128 -- the
full-resolution layers of Gen_UNet2D, TG:398-409, 482-495) against fp64 F.conv2d references, through depgan_op_conv2d.
Covers every epilogue (plain, FiLM residual, add / mask, fused 1x1 head with and without the bf16 output), one / two
input sources, band heights 16 and 32, several bands per CTA, and bit-identity with the tile kernel's contract where it
is exact (all-integer inputs)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_conv import _bf, _rand, ref_conv

pytestmark = pytest.mark.gpu

ROW_CASES = [  # N, H, W, c0, c1   (W a multiple of 256; odd heights and tiny segments exercise the edge paths)
    (2, 32, 256, 32, 0), (1, 16, 512, 32, 0), (3, 64, 256, 32, 0), (2, 48, 256, 32, 0), (5, 7, 256, 32, 0),
    (2, 32, 256, 64, 32), (1, 33, 512, 64, 32), (2, 32, 256, 32, 32), (1, 16, 256, 96, 0), (3, 1, 256, 32, 0),
    (2, 2, 256, 32, 0), (1, 3, 256, 64, 32),
]


def _tol(want):
    return 1e-2 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("N,H,W,c0,c1", ROW_CASES)
def test_row_kernel_plain(N, H, W, c0, c1):
    from depgan_b200 import conv2d_op
    x = _bf(_rand((N, H, W, c0), 1))
    x1 = _bf(_rand((N, H, W, c1), 2)) if c1 else None
    w = _bf(_rand((3, 3, c0 + c1, 32), 3, 1.0 / np.sqrt(9 * (c0 + c1))))
    sc, sh = 1 + 0.1 * _rand((32,), 4), 0.1 * _rand((32,), 5)
    got = conv2d_op(x.cuda(), w.cuda(), x1=None if x1 is None else x1.cuda(), scale=sc, shift=sh, relu=True).cpu()
    want, _ = ref_conv(x, w, x1, sc, sh, relu=True)
    err = float((got - want).abs().max())
    assert err <= _tol(want), err


def test_row_kernel_is_exact_on_integer_data():
    """Small-integer activations and weights: every product and sum is exact in bf16 x bf16 -> fp32, so the result must
    equal the reference bit for bit -- catches any tap / row / column mix-up that a tolerance could hide."""
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 32, 256, 32
    g = torch.Generator().manual_seed(7)
    x = torch.randint(-3, 4, (N, H, W, c), generator=g).float()
    w = torch.randint(-2, 3, (3, 3, c, 32), generator=g).float()
    got = conv2d_op(x.cuda(), w.cuda()).cpu()
    want, _ = ref_conv(x, w)
    assert float(want.abs().max()) < 256  # bf16 holds these integers exactly
    assert torch.equal(got, want)


@pytest.mark.parametrize("N,H,W", [(2, 32, 256), (1, 32, 512), (3, 17, 256)])
def test_row_kernel_film_residual(N, H, W):
    from depgan_b200 import conv2d_op
    c = 32
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.08))
    sc, sh = 1 + 0.1 * _rand((c,), 6), 0.1 * _rand((c,), 7)
    g, b = 1 + 0.3 * _rand((N, c), 3), 0.2 * _rand((N, c), 4)
    for res in (x, _bf(_rand((N, H, W, c), 5))):       # the nets pass the conv input itself as the residual
        got = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, film=(g, b), res=res.cuda()).cpu()
        want, _ = ref_conv(x, w, scale=sc, shift=sh, film=(g, b), res=res)
        err = float((got - want).abs().max())
        assert err <= _tol(want), err


def test_row_kernel_add_and_mask():
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 31, 256, 32
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.08))
    add, mask = _bf(_rand((N, H, W, c), 6)), _bf(_rand((N, H, W, c), 7))
    for kw in ({"add": add}, {"mask": mask}, {"add": add, "mask": mask}):
        got = conv2d_op(x.cuda(), w.cuda(), **{k: v.cuda() for k, v in kw.items()}).cpu()
        want, _ = ref_conv(x, w, **kw)
        err = float((got - want).abs().max())
        assert err <= _tol(want), (list(kw), err)


@pytest.mark.parametrize("nc,act", [(4, 1), (1, 0), (3, 2)])
def test_row_kernel_fused_head(nc, act):
    """conv2d_gen_17 + gen_segmentation (1x1, tanh / softmax) fused: the head is computed from the fp32 values before
    the bf16 rounding of the stored tensor."""
    from depgan_b200 import conv2d_op
    N, H, W, c = 2, 33, 256, 32
    x, w = _bf(_rand((N, H, W, c), 1)), _bf(_rand((3, 3, c, c), 2, 0.08))
    sc, sh = 1 + 0.1 * _rand((c,), 6), 0.1 * _rand((c,), 7)
    hw, hb = 0.3 * _rand((c, nc), 8), 0.1 * _rand((nc,), 9)
    got, ex = conv2d_op(x.cuda(), w.cuda(), scale=sc, shift=sh, relu=True, head=(hw, hb, act))
    want, _ = ref_conv(x, w, scale=sc, shift=sh, relu=True)
    assert float((got.cpu() - want).abs().max()) <= _tol(want)
    logits = want.double() @ hw.double() + hb.double()
    ref_h = torch.tanh(logits) if act == 0 else torch.softmax(logits, dim=-1) if act == 1 else logits
    assert float((ex["head"].cpu().double() - ref_h).abs().max()) <= 2e-3


def test_row_kernel_many_bands_per_cta_match_first_pass():
    """40 slices x 256 rows: every CTA walks ~70 rows across image borders, wrapping all rings many times; slices are
    independent, so slice k of the batch equals slice k computed alone (a different split of the rows over CTAs)."""
    from depgan_b200 import conv2d_op
    H, W, c = 256, 256, 32
    x = _bf(_rand((40, H, W, c), 1))
    w = _bf(_rand((3, 3, c, c), 2, 0.08))
    full = conv2d_op(x.cuda(), w.cuda(), relu=True).cpu()
    for k in (0, 17, 39):
        one = conv2d_op(x[k:k + 1].cuda(), w.cuda(), relu=True).cpu()
        assert torch.equal(full[k:k + 1], one), k
    want, _ = ref_conv(x[:1], w, relu=True)
    assert float((full[:1] - want).abs().max()) <= _tol(want)
