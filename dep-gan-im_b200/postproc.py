"""Inference repeat loop and DEM post-processing on the GPU (EG:616-628, 673-741; EU:553-570, 597-600).

The float64 accumulation order (repeat 0, 1, ...) and every comparison follow the reference's NumPy code, so the
label maps and WMH voxel counts are bit-exact given the same predictions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import _stream, _torch


class DemAccumulator:
    """acc = zeros(float64); acc += float64(pred_f32 * mask_f32) per repeat; mean = acc / n (EG:617-628)."""

    def __init__(self, shape, device="cuda:0"):
        torch = _torch()
        self.torch, self.device = torch, torch.device(device)
        self.shape = tuple(shape)
        self.acc = torch.zeros(self.shape, dtype=torch.float64, device=self.device)
        self.n = 0

    @classmethod
    def wrap(cls, acc):
        """Accumulator over an existing (already zeroed) float64 CUDA tensor."""
        self = cls.__new__(cls)
        self.torch, self.device, self.shape, self.acc, self.n = _torch(), acc.device, tuple(acc.shape), acc, 0
        return self

    def add(self, pred, mask):
        """pred (Z,H,W[,C]) float32 CUDA tensor, mask (Z,H,W) float32 CUDA tensor."""
        torch = self.torch
        chan = pred.numel() // max(1, mask.numel()) if mask.numel() else 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_dem_accumulate(self.acc.data_ptr(), pred.data_ptr(), mask.data_ptr(),
                                                        pred.numel(), int(chan), _stream(torch)), "dem_accumulate")
        self.n += 1


def dem_postproc_device(x, nicg, acc, n_repeat, mask, thr):
    """Returns (dem f64, fake2 f64, labels u8, count tensor u64[1]) CUDA tensors for a (Z,H,W) volume."""
    torch = _torch()
    dev = acc.device
    npix = acc.numel()
    dem = torch.empty_like(acc)
    fake2 = torch.empty_like(acc)
    labels = torch.empty(acc.shape, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().depgan_dem_postproc(x.data_ptr(), int(nicg), acc.data_ptr(), float(n_repeat),
                                                  mask.data_ptr(), float(thr), dem.data_ptr(), fake2.data_ptr(),
                                                  labels.data_ptr(), count.data_ptr(), npix, _stream(torch)),
                   "dem_postproc")
    return dem, fake2, labels, count


def dem_pipeline(x1, preds, mask, thr, nicg=1, device="cuda:0"):
    """NumPy convenience wrapper: x1 (Z,H,W,nicg) f32, preds list of (Z,H,W) f32, mask (Z,H,W) f32.
    Returns (dem f64, fake2 f64, labels uint8, count int) as the reference computes them."""
    torch = _torch()
    dev = torch.device(device)
    x = torch.from_numpy(np.ascontiguousarray(x1, np.float32)).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(mask, np.float32)).to(dev)
    acc = DemAccumulator(m.shape, dev)
    for p in preds:
        acc.add(torch.from_numpy(np.ascontiguousarray(p, np.float32)).to(dev), m)
    dem, fake2, labels, count = dem_postproc_device(x, nicg, acc.acc, float(len(preds)), m, thr)
    return dem.cpu().numpy(), fake2.cpu().numpy(), labels.cpu().numpy(), int(count.item())


def uresnet_pipeline(preds, mask, device="cuda:0"):
    """preds list of (Z,H,W,C) f32 softmax maps, mask (Z,H,W).  Returns (mean f64, labels uint8, count)."""
    torch = _torch()
    dev = torch.device(device)
    m = torch.from_numpy(np.ascontiguousarray(mask, np.float32)).to(dev)
    chan = preds[0].shape[-1]
    acc = DemAccumulator(preds[0].shape, dev)
    for p in preds:
        acc.add(torch.from_numpy(np.ascontiguousarray(p, np.float32)).to(dev), m)
    labels = torch.empty(m.shape, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        mean = torch.empty_like(acc.acc)  # IEEE division in our kernel (torch.div by a scalar multiplies by 1/n)
        _lib.check(_lib.lib().depgan_uresnet_labels(acc.acc.data_ptr(), float(len(preds)), int(chan),
                                                    mean.data_ptr(), labels.data_ptr(), count.data_ptr(), m.numel(),
                                                    _stream(torch)), "uresnet_labels")
    return mean.cpu().numpy(), labels.cpu().numpy(), int(count.item())


# ---------------------------------------------------------------------------------------------------------
# Evaluation row of the testing scripts (EG:688-807, EU:601-704)
# ---------------------------------------------------------------------------------------------------------
def label_confusion_device(fake_labels, real_labels):
    """4x4 confusion counts conf[real, fake] (int64 CUDA tensor) of two uint8 label maps with values 0..3."""
    torch = _torch()
    assert fake_labels.dtype == torch.uint8 and real_labels.dtype == torch.uint8
    assert fake_labels.numel() == real_labels.numel()
    conf = torch.zeros(16, dtype=torch.int64, device=fake_labels.device)
    with torch.cuda.device(fake_labels.device):
        _lib.check(_lib.lib().depgan_label_confusion(fake_labels.data_ptr(), real_labels.data_ptr(),
                                                     fake_labels.numel(), conf.data_ptr(), _stream(torch)),
                   "label_confusion")
    return conf.view(4, 4)


def dice_scores_from_confusion(conf):
    """(dice_1 .. dice_6) exactly as EG:745-794 computes them: (2*|A and B| + 1e-7) / (1e-7 + |A| + |B|) on
    1 shrink, 2 grow, 3 stay, 4 any WMH (label > 0), 5 changing (label 1 or 2), 6 stay again.  conf[real, fake]."""
    c = np.asarray(conf, dtype=np.int64).reshape(4, 4)
    smooth = 1e-7

    def dice(sel):
        sel = list(sel)
        inter = int(c[np.ix_(sel, sel)].sum())
        n_real = int(c[sel, :].sum())
        n_fake = int(c[:, sel].sum())
        return (inter * 2.0 + smooth) / (smooth + n_real + n_fake)

    return dice([1]), dice([2]), dice([3]), dice([1, 2, 3]), dice([1, 2]), dice([3])


def evaluation_row(conf, vol_1tp_ml, vol_2tp_ml, vol_out_ml):
    """The 18-entry CSV row of the testing scripts (EG:806-807 / EU:703-704):
    [true_pred, prog, true_prog, regg, true_regg, vol1, vol2, vol_out, mse, err, d5, d6, avg56, d1, d2, d3, d4,
    avg123] from the confusion counts and the three WMH volumes in ml."""
    err_vol = vol_out_ml - vol_2tp_ml                        # EG:688
    mse_vol = float(np.mean((vol_2tp_ml - vol_out_ml) ** 2))  # EG:689
    true_pred = true_prog = true_regg = prog = regg = 0
    if (vol_2tp_ml - vol_1tp_ml) >= 0:                       # EG:697-706
        prog = 1
        if vol_out_ml - vol_1tp_ml >= 0:
            true_pred = true_prog = 1
    else:
        regg = 1
        if vol_out_ml - vol_1tp_ml < 0:
            true_pred = true_regg = 1
    d1, d2, d3, d4, d5, d6 = dice_scores_from_confusion(conf)
    avg_all = (d1 + d2 + d3) / 3.0
    avg_56 = (d5 + d6) / 2.0
    return [true_pred, prog, true_prog, regg, true_regg, vol_1tp_ml, vol_2tp_ml, vol_out_ml, mse_vol, err_vol, d5, d6,
            avg_56, d1, d2, d3, d4, avg_all]


def evaluate_labels(fake_labels, real_labels, vol_1tp_ml, vol_2tp_ml, vol_out_ml, device="cuda:0"):
    """NumPy convenience wrapper: label maps (any integer / float dtype holding 0..3) -> the 18-entry row."""
    torch = _torch()
    dev = torch.device(device)
    f = torch.from_numpy(np.ascontiguousarray(fake_labels).astype(np.uint8)).to(dev)
    r = torch.from_numpy(np.ascontiguousarray(real_labels).astype(np.uint8)).to(dev)
    conf = label_confusion_device(f, r).cpu().numpy()
    return evaluation_row(conf, vol_1tp_ml, vol_2tp_ml, vol_out_ml)
