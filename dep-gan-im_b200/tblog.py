"""TensorBoard event files without TensorFlow: the reference's ``Logger`` (TG:167-248: ``log_scalar``, ``log_images``,
``log_histogram`` on a ``tf.summary.FileWriter``) as a small encoder of the on-disk format.

An event file is a sequence of TFRecords, each

    uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)

whose data is a serialized ``tensorflow.Event`` protobuf.  Only the handful of fields the reference writes are needed,
so the protobuf wire format is emitted by hand (varint / 64-bit / length-delimited / 32-bit fields):

    Event   { 1: wall_time double, 2: step int64, 3: file_version string, 5: summary Summary }
    Summary { 1: repeated Value }
    Value   { 1: tag string, 2: simple_value float, 4: image Image, 5: histo HistogramProto }
    Image   { 1: height, 2: width, 3: colorspace, 4: encoded_image_string (PNG) }
    HistogramProto { 1: min, 2: max, 3: num, 4: sum, 5: sum_squares (doubles), 6: packed bucket_limit, 7: packed bucket }

Images are PNG-encoded here as well (zlib + CRC from the standard library).  The reference renders single-channel images
with matplotlib's viridis colour map (TG:199-206, ``dtype=''``); the same look-up is reproduced from the map's published
control points.  tests/test_tblog.py reads the files back with the ``tensorboard`` package's own event reader where that
package is installed.
"""
from __future__ import annotations

import os
import socket
import struct
import time
import zlib

import numpy as np

__all__ = ["TensorBoardLogger", "encode_png", "viridis"]

# ---- CRC-32C (Castagnoli), table driven; TFRecord masks it so that CRCs of CRCs stay well distributed ----
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _record(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", _masked_crc(head)) + data + struct.pack("<I", _masked_crc(data))


# ---- protobuf wire format ----
def _varint(n: int) -> bytes:
    n &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _f_varint(field, v):
    return _varint(field << 3 | 0) + _varint(int(v))


def _f_double(field, v):
    return _varint(field << 3 | 1) + struct.pack("<d", float(v))


def _f_float(field, v):
    return _varint(field << 3 | 5) + struct.pack("<f", float(v))


def _f_bytes(field, b):
    b = b.encode("utf8") if isinstance(b, str) else bytes(b)
    return _varint(field << 3 | 2) + _varint(len(b)) + b


def _event(step=None, summary=None, file_version=None, wall_time=None):
    e = _f_double(1, time.time() if wall_time is None else wall_time)
    if step is not None:
        e += _f_varint(2, step)
    if file_version is not None:
        e += _f_bytes(3, file_version)
    if summary is not None:
        e += _f_bytes(5, summary)
    return e


# ---- PNG ----
def encode_png(img: np.ndarray) -> bytes:
    """uint8 array (H, W) greyscale, (H, W, 3) RGB or (H, W, 4) RGBA -> PNG bytes."""
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    ctype = {1: 0, 3: 2, 4: 6}[c]
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))  # filter type 0 on every scan line

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


# viridis at 11 equally spaced points (0.0, 0.1, ... 1.0) (matplotlib's _cm_listed data, rounded); linear interpolation in between
_VIRIDIS = np.array([[0.267004, 0.004874, 0.329415], [0.282623, 0.140926, 0.457517], [0.253935, 0.265254, 0.529983],
                     [0.206756, 0.371758, 0.553117], [0.163625, 0.471133, 0.558148], [0.127568, 0.566949, 0.550556],
                     [0.134692, 0.658636, 0.517649], [0.266941, 0.748751, 0.440573], [0.477504, 0.821444, 0.318195],
                     [0.741388, 0.873449, 0.149561], [0.993248, 0.906157, 0.143936]])


def viridis(img: np.ndarray) -> np.ndarray:
    """(H, W) float array -> (H, W, 4) uint8 RGBA, min-max normalised like ``plt.imsave(..., cmap='viridis')``."""
    a = np.asarray(img, np.float64)
    lo, hi = float(a.min()), float(a.max())
    t = np.zeros_like(a) if hi <= lo else (a - lo) / (hi - lo)
    x = t * (len(_VIRIDIS) - 1)
    i0 = np.clip(np.floor(x).astype(int), 0, len(_VIRIDIS) - 2)
    f = (x - i0)[..., None]
    rgb = _VIRIDIS[i0] * (1 - f) + _VIRIDIS[i0 + 1] * f
    out = np.empty(a.shape + (4,), np.uint8)
    out[..., :3] = np.clip(np.rint(rgb * 255), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


class TensorBoardLogger:
    """Drop-in for the reference's ``Logger(log_dir)`` (TG:167-248, created at TG:775).  The series are also kept in
    memory (``.series[tag] = [(step, value), ...]``) like ``trainer.ScalarLog``."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._fh = open(self.path, "ab")
        self._fh.write(_record(_event(file_version="brain.Event:2")))
        self._fh.flush()
        self.series = {}

    def _write_summary(self, values, step):
        self._fh.write(_record(_event(step=step, summary=b"".join(_f_bytes(1, v) for v in values))))
        self._fh.flush()

    def log_scalar(self, tag, value, step):
        self.series.setdefault(tag, []).append((int(step), float(value)))
        self._write_summary([_f_bytes(1, tag) + _f_float(2, value)], step)

    def log_images(self, tag, images, step, dtype="RGB", denorm=(0, 255)):
        """images: iterable of (H, W, 3) arrays in [-1, 1] (dtype 'RGB': mapped to 0..255 as TG:195-196) or single-channel
        arrays of any range (anything else: viridis, TG:199-201).  One Value per image, tagged '<tag>/<nr>'."""
        vals = []
        for nr, img in enumerate(images):
            img = np.asarray(img)
            if dtype == "RGB":
                px = ((img + 1) / 2 * denorm[1]).clip(denorm[0], denorm[1]).astype(np.uint8)
            else:
                px = viridis(np.squeeze(img))
            image = (_f_varint(1, px.shape[0]) + _f_varint(2, px.shape[1]) + _f_varint(3, px.shape[2] if px.ndim == 3 else 1) +
                     _f_bytes(4, encode_png(px)))
            vals.append(_f_bytes(1, "%s/%d" % (tag, nr)) + _f_bytes(4, image))
        self.series.setdefault(tag, []).append((int(step), float(len(vals))))
        self._write_summary(vals, step)

    def log_histogram(self, tag, values, step=0, bins=1000):
        values = np.asarray(values, np.float64)
        counts, edges = np.histogram(values, bins=bins)
        h = (_f_double(1, values.min()) + _f_double(2, values.max()) + _f_double(3, values.size) +
             _f_double(4, values.sum()) + _f_double(5, (values ** 2).sum()) +
             _f_bytes(6, struct.pack("<%dd" % (len(edges) - 1), *edges[1:])) +
             _f_bytes(7, struct.pack("<%dd" % len(counts), *counts.astype(np.float64))))
        self._write_summary([_f_bytes(1, tag) + _f_bytes(5, h)], step)

    def flush(self):
        self._fh.flush()

    def close(self):
        self._fh.close()
