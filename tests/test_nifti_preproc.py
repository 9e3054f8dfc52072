"""NIfTI-1 reader/writer and the reference's host-side data conventions (TG:93-149, EG:525-613, EU:486-521)."""
import gzip
import struct

import numpy as np
import pytest

from depgan_b200 import nifti, preproc


@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.int16, np.float64])
def test_nifti_round_trip(tmp_path, ext, dtype):
    rng = np.random.default_rng(0)
    vol = (rng.random((7, 9, 5)) * 100).astype(dtype)
    aff = np.array([[0.0, -0.9375, 0.0, 120.0], [0.9375, 0.0, 0.0, -110.0], [0.0, 0.0, 4.0, -70.0], [0, 0, 0, 1.0]])
    p = tmp_path / ("v" + ext)
    nifti.save(vol, aff, p)
    im = nifti.load(p)
    assert im.image.dtype == np.dtype(dtype) and np.array_equal(im.image, vol)
    assert np.allclose(im.affine, aff, atol=1e-5)
    assert np.allclose(im.pixdim, [0.9375, 0.9375, 4.0], atol=1e-6)


def test_nifti_header_layout_and_fortran_order(tmp_path):
    vol = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    p = tmp_path / "h.nii"
    nifti.save(vol, np.diag([1.0, 2.0, 3.0, 1.0]), p)
    raw = p.read_bytes()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 2, 3, 4)
    assert struct.unpack("<hh", raw[70:74]) == (16, 32)                      # float32
    assert struct.unpack("<f", raw[108:112])[0] == 352.0
    data = np.frombuffer(raw, np.float32, offset=352)
    assert np.array_equal(data, vol.reshape(-1, order="F"))                 # x runs fastest on disk


def test_nifti_scaling_and_big_endian(tmp_path):
    # a big-endian int16 file with scl_slope / scl_inter, written by hand
    vol = np.arange(24, dtype=">i2").reshape(2, 3, 4)
    h = bytearray(348)
    struct.pack_into(">i", h, 0, 348)
    struct.pack_into(">8h", h, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">hh", h, 70, 4, 16)
    struct.pack_into(">8f", h, 76, -1.0, 1.5, 1.5, 3.0, 2.5, 1, 1, 1)
    struct.pack_into(">fff", h, 108, 352.0, 0.5, 10.0)
    struct.pack_into(">hh", h, 252, 1, 0)                                    # qform only
    struct.pack_into(">6f", h, 256, 0.0, 0.0, 0.0, 5.0, 6.0, 7.0)
    h[344:348] = b"n+1\x00"
    p = tmp_path / "be.nii.gz"
    with gzip.open(p, "wb") as f:
        f.write(bytes(h) + b"\0\0\0\0" + vol.tobytes(order="F"))
    im = nifti.load(p)
    assert im.image.dtype == np.float64
    assert np.array_equal(im.image, vol.astype(np.float64) * 0.5 + 10.0)
    assert np.allclose(im.pixdim, [1.5, 1.5, 3.0]) and im.dt == 2.5
    want = np.diag([1.5, 1.5, -3.0, 1.0])                                    # qfac = -1 flips z
    want[:3, 3] = [5.0, 6.0, 7.0]
    assert np.allclose(im.affine, want)


def test_nifti_rejects_garbage(tmp_path):
    p = tmp_path / "bad.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(ValueError):
        nifti.load(p)


def test_data_prep_conventions():
    vol = np.random.default_rng(1).random((6, 5, 4))
    x = preproc.data_prep(vol)
    assert x.shape == (4, 6, 5, 1) and x.dtype == np.float32
    assert all(np.array_equal(x[z, :, :, 0], vol[:, :, z].astype(np.float32)) for z in range(4))
    back = preproc.data_prep_save(x)                                         # TG:121-128 undoes TG:105-119
    assert back.shape == vol.shape and np.array_equal(back, vol.astype(np.float32))


def test_map_image_to_intensity_range():
    img = np.array([[2.0, 4.0], [6.0, 10.0]])
    out = preproc.map_image_to_intensity_range(img, 0, 1)
    assert np.allclose(out, (img - 2.0) / 8.0)
    out = preproc.map_image_to_intensity_range(np.arange(101, dtype=np.float64), 0, 1, percentiles=10)
    assert out.min() == 0.0 and out.max() == 1.0 and np.isclose(out[50], 0.5)


def _ref_save_axes(a):
    """Test-side restatement of the reference's axis sequence before saving (TG:121-128)."""
    a = np.swapaxes(np.squeeze(a), 0, 2)
    return np.rot90(a)[::-1, ...]


def _ref_intensity_map(image, min_o, max_o, p):
    """Test-side restatement of TG:131-149 in the image's own float dtype (NumPy 1.x scalar casting)."""
    t = image.dtype.type if image.dtype.kind == "f" else np.float64
    lo, hi = t(np.percentile(image, p)), t(np.percentile(image, 100 - p))
    out = (image.astype(t) - lo) / (hi - lo) * t(max_o - min_o) + t(min_o)
    out[out > max_o] = max_o
    out[out < min_o] = min_o
    return out


@pytest.mark.parametrize("shape", [(7, 5, 3), (16, 16, 4), (3, 9, 2)])
def test_data_prep_save_equals_reference_axis_sequence(shape):
    rng = np.random.default_rng(sum(shape))
    stack = rng.random((shape[2], shape[0], shape[1], 1)).astype(np.float32)
    got = preproc.data_prep_save(stack)
    assert got.shape == shape and np.array_equal(got, _ref_save_axes(stack))
    assert np.array_equal(preproc.data_prep(got), stack)              # and it inverts data_prep exactly


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16])
@pytest.mark.parametrize("rng_out,p", [((0, 1), 0), ((-1, 1), 0), ((0, 255), 2), ((0.5, 2.5), 10)])
def test_map_image_to_intensity_range_matches_formula(dtype, rng_out, p):
    rng = np.random.default_rng(11)
    img = (rng.random((9, 8, 5)) * 200).astype(dtype)
    if np.dtype(dtype).kind == "u" and rng_out[0] < 0:
        with pytest.raises(AssertionError):
            preproc.map_image_to_intensity_range(img, rng_out[0], rng_out[1], p)
        return
    got = preproc.map_image_to_intensity_range(img, rng_out[0], rng_out[1], p)
    want = _ref_intensity_map(img, rng_out[0], rng_out[1], p)
    assert got.dtype == want.dtype and got.shape == img.shape
    assert np.array_equal(got, want)
    assert got.min() >= rng_out[0] and got.max() <= rng_out[1]
    if p == 0:                                                          # min-max scaling reaches both ends
        assert got.min() == rng_out[0] and np.isclose(got.max(), rng_out[1])
    # monotone: the map never swaps the order of two voxels
    order = np.argsort(img.reshape(-1), kind="stable")
    assert np.all(np.diff(got.reshape(-1)[order]) >= 0)


def test_map_image_to_intensity_range_edge_cases():
    with pytest.raises(AssertionError):
        preproc.map_image_to_intensity_range(np.zeros((2, 2), np.uint8), 0, 300)
    img = np.full((3, 3), 7.0, np.float32)                              # constant image: 0/0, NaN like the reference
    assert np.isnan(preproc.map_image_to_intensity_range(img, 0, 1)).all()
    src = np.arange(6, dtype=np.float32).reshape(2, 3)
    keep = src.copy()
    preproc.map_image_to_intensity_range(src, 0, 1)
    assert np.array_equal(src, keep)                                    # the input is not modified


def test_prepare_subject_dem_and_uresnet():
    rng = np.random.default_rng(2)
    X, Y, Z = 8, 8, 3
    im = rng.random((X, Y, Z)) - 0.2                     # some negative values
    flair = rng.random((X, Y, Z)) * 300
    icv1 = (rng.random((X, Y, Z)) > 0.2).astype(np.float32)
    icv2 = (rng.random((X, Y, Z)) > 0.2).astype(np.float32)
    sl1 = (rng.random((X, Y, Z)) > 0.9).astype(np.float32)
    x, m1, m2 = preproc.prepare_subject_dem(im, icv1, icv2, flair_1tp=flair, sl_1tp=sl1, nicg=2)
    assert x.shape == (Z, X, Y, 2) and x.dtype == np.float32
    keep = preproc.data_prep(icv1 * (1 - sl1))[..., 0]
    assert np.array_equal(m1, keep) and np.array_equal(m2, preproc.data_prep(icv2)[..., 0])
    assert (x[..., 0] >= 0).all() and (x[..., 0][keep == 0] == 0).all()
    assert np.allclose(x[..., 0], np.maximum(preproc.data_prep(im)[..., 0] * keep, 0))
    assert x[..., 1].min() == 0.0 and np.isclose(x[..., 1].max(), 1.0)      # FLAIR min-max normalised after masking
    xu, u1, u2 = preproc.prepare_subject_uresnet(flair, icv1, icv2, sl_1tp=sl1)
    assert xu.shape == (Z, X, Y, 1) and abs(float(xu.mean())) < 1e-5 and abs(float(xu.std()) - 1.0) < 1e-4
    z0, _, _ = preproc.prepare_subject_uresnet(np.zeros((X, Y, Z)), icv1, icv2)   # empty volume: NaN -> 0
    assert not np.isnan(z0).any() and (z0 == 0).all()
    assert preproc.wmh_volume_ml(np.ones((2, 2, 2)), [0.5, 0.5, 4.0]) == 8 * 1.0 / 1000


def test_save_subject_outputs_round_trip(tmp_path):
    from depgan_b200.infer import save_subject_outputs
    rng = np.random.default_rng(3)
    Z, X, Y = 4, 16, 16
    res = {"fake2": rng.random((Z, X, Y)), "dem": rng.random((Z, X, Y)) - 0.5,
           "labels": rng.integers(0, 4, (Z, X, Y)).astype(np.uint8)}
    aff = np.diag([0.9, 0.9, 4.0, 1.0])
    paths = save_subject_outputs(res, aff, tmp_path, "subj01")
    assert [p.split("subj01")[1] for p in paths] == ["_2tp_prob_fake.nii.gz", "_network_output.nii.gz",
                                                     "_2tp_code_fake.nii.gz"]
    for key, path in zip(("fake2", "dem", "labels"), paths):
        im = nifti.load(path)
        assert im.image.dtype == np.float32 and im.image.shape == (X, Y, Z)
        # reading the file back through data_prep gives the slices the network produced
        assert np.array_equal(preproc.data_prep(im)[..., 0], np.asarray(res[key]).astype(np.float32))
        assert np.allclose(im.affine, aff)
