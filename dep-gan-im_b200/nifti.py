"""Library-free NIfTI-1 reader / writer (``.nii`` and ``.nii.gz``) for the volumes the reference scripts move with
nibabel: ``nib.load(name).get_data()`` / ``.affine`` / ``.header['pixdim']`` (TG:93-102, EG:75-83) and
``nib.save(nib.Nifti1Image(arr.astype('float32'), affine), name)`` (EG:818-832).

Only what those calls touch is implemented: single-file NIfTI-1 (magic ``n+1``), little- or big-endian, the scalar
datatypes, ``scl_slope`` / ``scl_inter`` scaling as nibabel's ``get_data`` applies it, the sform / qform affine.  Data
are stored in the file in Fortran order (x fastest), exactly as nibabel writes them.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

__all__ = ["NiftiImage", "load", "save"]

# NIfTI-1 datatype code -> NumPy dtype
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


class NiftiImage:
    """What ``load_data`` keeps of a nibabel image: ``image`` (array), ``affine`` (4x4), ``pixdim`` (voxel sizes of the
    three spatial axes) and ``dt`` (``pixdim[4]``)  -- TG:93-102, EG:75-83."""

    def __init__(self, image, affine, pixdim=None, dt=0.0):
        self.image = image
        self.affine = np.asarray(affine, dtype=np.float64)
        if pixdim is None:  # nibabel derives the zooms from the affine's column norms
            pixdim = np.sqrt((self.affine[:3, :3] ** 2).sum(axis=0))
        self.pixdim = np.asarray(pixdim, dtype=np.float32)
        self.dt = float(dt)


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def _quaternion_affine(b, c, d, qfac, pixdim, offset):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    z = np.array([pixdim[0], pixdim[1], pixdim[2] * (-1.0 if qfac < 0 else 1.0)])
    A = np.eye(4)
    A[:3, :3] = R * z
    A[:3, 3] = offset
    return A


def load(path):
    """Reads ``path`` and returns a :class:`NiftiImage` whose ``image`` equals nibabel's ``get_data()``: the stored
    dtype when no scaling is set, else float64 ``raw * scl_slope + scl_inter``."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError("%s: not a NIfTI-1 file (shorter than the 348-byte header)" % path)
    end = "<"
    if struct.unpack("<i", raw[0:4])[0] != 348:
        end = ">"
        if struct.unpack(">i", raw[0:4])[0] != 348:
            raise ValueError("%s: sizeof_hdr is not 348" % path)
    magic = raw[344:348]
    if magic[:3] != b"n+1":
        raise ValueError("%s: only single-file NIfTI-1 (magic n+1) is supported, found %r" % (path, magic))
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(end + "fff", raw[108:120])
    qform_code, sform_code = struct.unpack(end + "hh", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(end + "6f", raw[256:280])
    srow = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype code %d" % (path, datatype))
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise ValueError("%s: bad dim[0] = %d" % (path, ndim))
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    count = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    data = np.frombuffer(raw, dtype=dt, count=count, offset=off).reshape(shape, order="F")
    data = data.astype(dt.newbyteorder("="), copy=True)
    if slope != 0.0 and not np.isnan(slope) and not (slope == 1.0 and inter == 0.0):
        data = data.astype(np.float64) * float(slope) + float(inter)
    if sform_code > 0:
        affine = np.vstack([srow, [0.0, 0.0, 0.0, 1.0]])
    elif qform_code > 0:
        affine = _quaternion_affine(qb, qc, qd, pixdim[0], pixdim[1:4], (qx, qy, qz))
    else:  # nibabel's fallback: scaling by the zooms, origin at the volume centre (not needed by the reference)
        affine = np.diag([pixdim[1], pixdim[2], pixdim[3], 1.0]).astype(np.float64)
    return NiftiImage(data, affine, pixdim[1:4], pixdim[4])


def save(image, affine, path):
    """``nib.save(nib.Nifti1Image(image, affine), path)`` for a scalar array of up to 7 dimensions: sform = affine
    (code 2, "aligned"), qform unset, zooms = column norms of the affine, no intensity scaling."""
    arr = np.asarray(image)
    key = arr.dtype.newbyteorder("=").str[1:]
    if key not in _CODES:
        raise ValueError("unsupported dtype %s for NIfTI output" % arr.dtype)
    if not 1 <= arr.ndim <= 7:
        raise ValueError("NIfTI stores 1 to 7 dimensions")
    A = np.asarray(affine, dtype=np.float64)
    if A.shape != (4, 4):
        raise ValueError("affine must be 4x4")
    zooms = np.sqrt((A[:3, :3] ** 2).sum(axis=0))
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    pixdim = [1.0] + [float(z) for z in zooms[:min(3, arr.ndim)]] + [1.0] * (7 - min(3, arr.ndim))
    h = bytearray(348)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, *dim)
    struct.pack_into("<hh", h, 70, _CODES[key], arr.dtype.itemsize * 8)
    struct.pack_into("<8f", h, 76, *pixdim)
    struct.pack_into("<fff", h, 108, 352.0, 1.0, 0.0)  # vox_offset, scl_slope, scl_inter
    h[123] = 2                                           # xyzt_units: millimetres
    struct.pack_into("<hh", h, 252, 0, 2)               # qform_code, sform_code
    struct.pack_into("<12f", h, 280, *A[:3, :].reshape(-1))
    h[344:348] = b"n+1\x00"
    payload = np.asfortranarray(arr.astype(arr.dtype.newbyteorder("<"), copy=False)).tobytes(order="F")
    with _open(path, "wb") as f:
        f.write(bytes(h))
        f.write(b"\x00\x00\x00\x00")  # no header extensions
        f.write(payload)
