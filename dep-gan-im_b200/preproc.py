"""Host-side data conventions of the reference scripts around the hot path: slice extraction from a NIfTI volume,
the inverse used before saving, intensity normalisation, and the per-subject masking that produces the network
inputs (TG:105-149, EG:525-613, EU:486-521).  Plain NumPy, as in the reference; volumes come from ``nifti.load``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["data_prep", "data_prep_save", "map_image_to_intensity_range", "prepare_subject_dem",
           "prepare_subject_uresnet", "wmh_volume_ml"]


def data_prep(image):
    """(X, Y, Z) volume -> (Z, X, Y, 1) float32 stack of axial slices (the reference's ``data_prep``, TG:105-119: slice
    z of the stack is ``image[:, :, z]``, with a trailing channel axis).  Accepts the array or an object with an
    ``image`` attribute (``load_data`` / ``NiftiImage``)."""
    vol = np.asarray(getattr(image, "image", image))
    if vol.ndim != 3:
        raise ValueError("data_prep expects an (X, Y, Z) volume, got shape %s" % (vol.shape,))
    return np.ascontiguousarray(np.moveaxis(vol, 2, 0), dtype=np.float32)[..., np.newaxis]


def data_prep_save(image_data):
    """Inverse of :func:`data_prep`, applied before a prediction is saved (the reference's ``data_prep_save``,
    TG:121-128, whose squeeze / swapaxes(0, 2) / rot90 / flip sequence composes to exactly this axis move):
    (Z, X, Y[, 1]) -> (X, Y, Z) with ``out[x, y, z] = in[z, x, y]``."""
    stack = np.squeeze(np.asarray(image_data))
    if stack.ndim != 3:
        raise ValueError("data_prep_save expects a (Z, X, Y[, 1]) stack, got shape %s" % (np.shape(image_data),))
    return np.moveaxis(stack, 0, 2)


def map_image_to_intensity_range(image, min_o, max_o, percentiles=0):
    """Affine intensity normalisation used for the FLAIR channel (the reference's ``map_image_to_intensity_range``,
    TG:131-149): the ``percentiles``-th and ``(100 - percentiles)``-th percentiles of ``image`` are sent to ``min_o`` and
    ``max_o`` and everything outside is clamped.  ``percentiles == 0`` is plain min-max scaling.

    Floating-point images keep their dtype (the reference ran NumPy 1.x, where the float64 percentile scalars do not
    promote a float32 volume); integer images are computed in float64.  As in the reference, unsigned images reject a
    target range they could not hold (AssertionError)."""
    img = np.asarray(image)
    if img.dtype.kind == "u" and img.dtype.itemsize <= 4:
        if min_o < 0:
            raise AssertionError("unsigned image (%s) cannot be mapped to a range starting at %g" % (img.dtype, min_o))
        if img.dtype.itemsize == 1 and max_o > 255:
            raise AssertionError("uint8 image cannot be mapped to a range ending at %g" % max_o)
    work = img.dtype.type if img.dtype.kind == "f" else np.float64
    lo, hi = work(np.percentile(img, percentiles)), work(np.percentile(img, 100 - percentiles))
    with np.errstate(invalid="ignore", divide="ignore"):  # a constant image divides by zero, as in the reference
        out = (img.astype(work, copy=True) - lo) / (hi - lo)
    out *= work(max_o - min_o)
    out += work(min_o)
    np.clip(out, work(min_o), work(max_o), out=out)  # NaNs stay NaN, like the reference's two masked assignments
    return out


def _sq(v):
    return np.squeeze(data_prep(v))


def prepare_subject_dem(im_or_pm_1tp, icv_1tp, icv_2tp, flair_1tp=None, sl_1tp=None, sl_2tp=None, nicg=1):
    """Network input and masks of one subject for the DEP-GAN testing path (EG:525-613).

    Arguments are (X, Y, Z) volumes (arrays or ``NiftiImage``): the baseline irregularity / probability map, the two
    intracranial-volume masks, optionally the baseline FLAIR (``nicg == 2``) and the stroke-lesion masks.
    Returns ``x (Z, X, Y, nicg)`` float32 and ``icv_and_sl_mask_1tp``, ``icv_and_sl_mask_2tp`` ``(Z, X, Y)``:
    non-brain and stroke-lesion voxels zeroed (EG:533-566), negative map values clamped to 0 (EG:582-584), FLAIR
    min-max normalised to [0, 1] after masking (EG:574-578), channels concatenated map-first (EG:602-611)."""
    i1, i2 = _sq(icv_1tp), _sq(icv_2tp)
    base = np.multiply(_sq(im_or_pm_1tp), i1)
    flair = np.multiply(_sq(flair_1tp), i1) if flair_1tp is not None else None
    mask1, mask2 = i1, i2
    if sl_1tp is not None:
        s1 = 1 - _sq(sl_1tp)
        base = np.multiply(base, s1)
        if flair is not None:
            flair = np.multiply(flair, s1)
        mask1 = np.multiply(mask1, s1)
    if sl_2tp is not None:
        mask2 = np.multiply(i2, 1 - _sq(sl_2tp))
    if flair is not None:
        flair = map_image_to_intensity_range(flair, 0, 1, percentiles=0)
    base[base < 0] = 0
    sx, sy, sz = base.shape
    x = np.reshape(base, (sx, sy, sz, 1))
    if nicg == 2:
        if flair is None:
            raise ValueError("nicg == 2 needs the baseline FLAIR volume")
        x = np.concatenate((x, np.reshape(flair, (sx, sy, sz, 1))), axis=-1)
    return x.astype(np.float32), mask1, mask2


def prepare_subject_uresnet(flair_1tp, icv_1tp, icv_2tp, sl_1tp=None, sl_2tp=None):
    """Network input and masks of one subject for the DEP-UResNet testing path (EU:486-521): masked FLAIR,
    z-scored over the whole volume, NaNs (an empty volume) replaced by zero.  Returns ``x (Z, X, Y, 1)`` float32 and
    the two ``(Z, X, Y, 1)`` masks (the UResNet script keeps the channel axis on its masks)."""
    f, i1, i2 = data_prep(flair_1tp), data_prep(icv_1tp), data_prep(icv_2tp)
    brain = np.multiply(f, i1)
    mask1, mask2 = i1, i2
    if sl_1tp is not None:
        s1 = 1 - data_prep(sl_1tp)
        brain = np.multiply(brain, s1)
        mask1 = np.multiply(mask1, s1)
    if sl_2tp is not None:
        mask2 = np.multiply(i2, 1 - data_prep(sl_2tp))
    with np.errstate(invalid="ignore", divide="ignore"):
        brain = (brain - np.mean(brain)) / np.std(brain)
    return np.nan_to_num(brain).astype(np.float32), mask1, mask2


def wmh_volume_ml(mask, pixdim):
    """Voxel count x voxel volume in ml (EG:640-641, 681-684): ``count_nonzero(mask) * prod(pixdim) / 1000``."""
    return np.count_nonzero(mask) * np.prod(pixdim) / 1000
