"""TensorBoard event files written without TensorFlow (the reference's Logger, TG:167-248) -- read back with the
tensorboard package's own reader, plus format-level checks that do not need it."""
import struct
import zlib

import numpy as np
import pytest

from depgan_b200 import tblog


def test_crc32c_known_answers():
    assert tblog.crc32c(b"") == 0
    assert tblog.crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    assert tblog.crc32c(b"\x00" * 32) == 0x8A9136AA           # RFC 3720 B.4


def test_png_roundtrip():
    rng = np.random.default_rng(0)
    for shape in ((5, 7), (4, 6, 3), (3, 3, 4)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        b = tblog.encode_png(img)
        assert b[:8] == b"\x89PNG\r\n\x1a\n"
        w, h, depth, ctype = struct.unpack(">IIBB", b[16:26])
        c = 1 if img.ndim == 2 else img.shape[2]
        assert (h, w, depth, ctype) == (shape[0], shape[1], 8, {1: 0, 3: 2, 4: 6}[c])
        pos, data = 8, b""
        while pos < len(b):                                   # walk the chunks, check their CRCs, collect IDAT
            (ln,), tag = struct.unpack(">I", b[pos:pos + 4]), b[pos + 4:pos + 8]
            body = b[pos + 8:pos + 8 + ln]
            assert struct.unpack(">I", b[pos + 8 + ln:pos + 12 + ln])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
            data += body if tag == b"IDAT" else b""
            pos += 12 + ln
        raw = zlib.decompress(data)
        rows = [raw[y * (1 + w * c) + 1:(y + 1) * (1 + w * c)] for y in range(h)]
        assert np.array_equal(np.frombuffer(b"".join(rows), np.uint8).reshape(h, w, c), img.reshape(h, w, c))


def test_viridis_end_points_and_monotone_luminance():
    ramp = np.linspace(0.0, 1.0, 256)[None, :]
    rgba = tblog.viridis(ramp)[0]
    assert tuple(rgba[0][:3]) == (68, 1, 84) and tuple(rgba[-1][:3]) == (253, 231, 37)   # matplotlib's end colours
    lum = rgba[:, :3].astype(float) @ np.array([0.2126, 0.7152, 0.0722])
    assert np.all(np.diff(lum) > -0.5)                                                   # perceptually increasing


def test_event_file_reads_back_with_tensorboard(tmp_path):
    loader = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    log = tblog.TensorBoardLogger(str(tmp_path))
    for step in range(3):
        log.log_scalar("errG_losses", 1.5 * step, step)
    imgs = np.random.default_rng(1).standard_normal((2, 8, 6, 1)).astype(np.float32)
    log.log_images("attributed_img_step0", imgs, 0, "")
    log.log_images("rgb", np.zeros((1, 4, 4, 3), np.float32), 7)
    log.log_histogram("w", np.arange(100.0), step=5, bins=10)
    log.close()
    events = list(loader.EventFileLoader(log.path).Load())
    assert events[0].file_version == "brain.Event:2"
    scal = [(e.step, v.tag, v.simple_value if v.HasField("simple_value") else None)
            for e in events[1:] for v in e.summary.value]
    # the loader may migrate scalars into tensors; accept either representation
    got = []
    for e in events[1:]:
        for v in e.summary.value:
            if v.tag == "errG_losses":
                got.append((e.step, v.simple_value if v.HasField("simple_value") else float(v.tensor.float_val[0])))
    assert got == [(0, 0.0), (1, 1.5), (2, 3.0)], scal
    tags = [v.tag for e in events for v in e.summary.value]
    assert "attributed_img_step0/0" in tags and "attributed_img_step0/1" in tags and "rgb/0" in tags and "w" in tags
    assert log.series["errG_losses"] == [(0, 0.0), (1, 1.5), (2, 3.0)]


def test_records_are_framed_as_tfrecords(tmp_path):
    log = tblog.TensorBoardLogger(str(tmp_path))
    log.log_scalar("a", 2.0, 4)
    log.close()
    b = open(log.path, "rb").read()
    pos, n = 0, 0
    while pos < len(b):
        (ln,) = struct.unpack("<Q", b[pos:pos + 8])
        assert struct.unpack("<I", b[pos + 8:pos + 12])[0] == tblog._masked_crc(b[pos:pos + 8])
        data = b[pos + 12:pos + 12 + ln]
        assert struct.unpack("<I", b[pos + 12 + ln:pos + 16 + ln])[0] == tblog._masked_crc(data)
        pos += 16 + ln
        n += 1
    assert n == 2 and pos == len(b)
