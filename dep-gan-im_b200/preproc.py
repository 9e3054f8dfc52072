"""Host-side data conventions of the reference scripts around the hot path: slice extraction from a NIfTI volume,
the inverse used before saving, intensity normalisation, and the per-subject masking that produces the network
inputs (TG:105-149, EG:525-613, EU:486-521).  Plain NumPy, as in the reference; volumes come from ``nifti.load``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["data_prep", "data_prep_save", "map_image_to_intensity_range", "prepare_subject_dem",
           "prepare_subject_uresnet", "wmh_volume_ml"]


def data_prep(image):
    """(X, Y, Z) volume -> (Z, X, Y, 1) float32 stack of axial slices (TG:105-119: ``image[:, :, z]`` per slice, channel
    axis appended).  Accepts the array or an object with an ``image`` attribute (``load_data`` / ``NiftiImage``)."""
    image = getattr(image, "image", image)
    images = np.array([image[:, :, z] for z in range(image.shape[2])], dtype="float32")
    return np.expand_dims(images, axis=3)


def data_prep_save(image_data):
    """Inverse of :func:`data_prep` applied before ``nib.save`` (TG:121-128): squeeze, swap axes 0 and 2, rotate by
    90 degrees, flip the first axis -- (Z, X, Y[, 1]) back to (X, Y, Z)."""
    image_data = np.squeeze(image_data)
    output_img = np.swapaxes(image_data, 0, 2)
    output_img = np.rot90(output_img)
    return output_img[::-1, ...]


def map_image_to_intensity_range(image, min_o, max_o, percentiles=0):
    """Linear map of [percentile(p), percentile(100 - p)] onto [min_o, max_o], clipped (TG:131-149)."""
    if image.dtype in [np.uint8, np.uint16, np.uint32]:
        assert min_o >= 0, 'Input image type is uintXX but you selected a negative min_o: %f' % min_o
    if image.dtype == np.uint8:
        assert max_o <= 255, 'Input image type is uint8 but you selected a max_o > 255: %f' % max_o
    min_i = np.percentile(image, 0 + percentiles)
    max_i = np.percentile(image, 100 - percentiles)
    image = (np.divide((image - min_i), max_i - min_i) * (max_o - min_o) + min_o).copy()
    image[image > max_o] = max_o
    image[image < min_o] = min_o
    return image


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
