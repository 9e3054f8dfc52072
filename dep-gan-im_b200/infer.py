"""Per-subject inference drivers (EG:616-628 + 673-741, EU:553-570 + 597-603) and subject sharding for the
whole-cohort sweep (BASELINE configs[4]; SURVEY 8e: subjects are independent, so ranks take disjoint subjects and
no collective is involved).

    res = predict_subject_dem(netG, vol_1tp, mask_2tp, thr, n_repeat=10, seed=0)
    res["dem"], res["fake2"], res["labels"], res["wmh_voxels"]

Everything between the host volume and the returned maps stays on the GPU: n_repeat generator forwards with fresh
N(0,1) noise (the reference draws it unseeded; here the caller passes a seed or the noise itself so results are
reproducible), float64 accumulation of mask * prediction, mean, clip, threshold count and label map.
"""
from __future__ import annotations

import zlib

import numpy as np

from . import postproc
from .api import _torch

__all__ = ["shard_range", "shard_items", "predict_subject_dem", "predict_subject_uresnet", "evaluate_subject",
           "save_subject_outputs", "cohort_sweep"]


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items for `rank` of `world`, sizes differing by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_items(items, rank, world):
    lo, hi = shard_range(len(items), rank, world)
    return list(items[lo:hi])


def _noise(rng, noises, rep, z_shape):
    if noises is not None:
        return np.ascontiguousarray(noises[rep], np.float32)
    return rng.standard_normal(z_shape).astype(np.float32)  # np.random.normal(0, 1, (Z, 32, 1))  EG:620


def _forward_volume(net, xd, zd, out):
    """Runs the generator over a (Z,...) volume in chunks of max_batch (Keras predict batching, EG:621)."""
    bs = net.cfg.max_batch
    for i in range(0, xd.shape[0], bs):
        net.forward_device(xd[i:i + bs], zd[i:i + bs], out[i:i + bs])
    return out


def predict_subject_dem(netG, vol_1tp, mask_2tp, thr, n_repeat=10, seed=0, noises=None):
    """vol_1tp (Z,H,W,nicg) f32, mask_2tp (Z,H,W) f32 -> dict(dem f64, fake2 f64, labels u8, wmh_voxels int)."""
    torch = _torch()
    dev = netG.device
    x = torch.from_numpy(np.ascontiguousarray(vol_1tp, np.float32)).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(mask_2tp, np.float32)).to(dev)
    Z = x.shape[0]
    acc = postproc.DemAccumulator((Z, netG.cfg.H, netG.cfg.W), dev)
    out = torch.empty((Z, netG.cfg.H, netG.cfg.W, 1), dtype=torch.float32, device=dev)
    rng = np.random.default_rng(seed)
    for rep in range(n_repeat):
        z = torch.from_numpy(_noise(rng, noises, rep, (Z, netG.cfg.noise_len, 1))).to(dev)
        _forward_volume(netG, x, z, out)
        acc.add(out, m)                                                    # EG:622-624
    dem, fake2, labels, count = postproc.dem_postproc_device(x, netG.cfg.nicg, acc.acc, float(n_repeat), m, thr)
    return {"dem": dem.cpu().numpy(), "fake2": fake2.cpu().numpy(), "labels": labels.cpu().numpy(),
            "wmh_voxels": int(count.item())}


def predict_subject_uresnet(net, vol, mask, n_repeat=10, seed=0, noises=None):
    """vol (Z,H,W,1) z-scored FLAIR, mask (Z,H,W) -> dict(prob_mean f64 (Z,H,W,4), labels u8, wmh_voxels int)."""
    import ctypes as C  # noqa: F401
    from . import _lib
    from .api import _stream
    torch = _torch()
    dev = net.device
    x = torch.from_numpy(np.ascontiguousarray(vol, np.float32)).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(mask, np.float32)).to(dev)
    Z, nc = x.shape[0], net.nc_out
    acc = postproc.DemAccumulator((Z, net.cfg.H, net.cfg.W, nc), dev)
    out = torch.empty((Z, net.cfg.H, net.cfg.W, nc), dtype=torch.float32, device=dev)
    rng = np.random.default_rng(seed)
    for rep in range(n_repeat):
        z = torch.from_numpy(_noise(rng, noises, rep, (Z, net.cfg.noise_len, 1))).to(dev)
        _forward_volume(net, x, z, out)
        acc.add(out, m)                                                    # EU:559-560
    mean = torch.empty_like(acc.acc)
    labels = torch.empty(m.shape, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().depgan_uresnet_labels(acc.acc.data_ptr(), float(n_repeat), nc, mean.data_ptr(),
                                                    labels.data_ptr(), count.data_ptr(), m.numel(), _stream(torch)),
                   "uresnet_labels")
    return {"prob_mean": mean.cpu().numpy(), "labels": labels.cpu().numpy(), "wmh_voxels": int(count.item())}


def evaluate_subject(result, real_labels, vol_1tp_ml, vol_2tp_ml, voxel_mm3, device="cuda:0"):
    """The 18-entry evaluation row of the testing scripts (EG:681-684, 688-807 / EU:597-704) for one subject:
    vol_out_ml = wmh_voxels * prod(pixdim) / 1000, volume-direction flags, six Dice scores from the GPU confusion
    counts of the predicted label map against `real_labels` (the ground-truth code volume, values 0..3)."""
    vol_out_ml = result["wmh_voxels"] * voxel_mm3 / 1000
    return postproc.evaluate_labels(result["labels"], real_labels, vol_1tp_ml, vol_2tp_ml, vol_out_ml, device=device)


def save_subject_outputs(result, affine, out_dir, name):
    """Writes the three volumes the DEP-GAN testing script saves per subject (EG:813-832): the predicted follow-up map
    ``<name>_2tp_prob_fake.nii.gz``, the disease evolution map ``<name>_network_output.nii.gz`` and the label map
    ``<name>_2tp_code_fake.nii.gz`` -- each through ``data_prep_save``, as float32, with the baseline scan's affine.
    Returns the three paths."""
    import os
    from . import nifti, preproc
    paths = []
    for key, suffix in (("fake2", "_2tp_prob_fake"), ("dem", "_network_output"), ("labels", "_2tp_code_fake")):
        vol = preproc.data_prep_save(np.asarray(result[key])).astype("float32")
        path = os.path.join(str(out_dir), str(name) + suffix + ".nii.gz")
        nifti.save(vol, affine, path)
        paths.append(path)
    return paths


def cohort_sweep(net, subjects, thr, rank=0, world=1, n_repeat=10, kind="dem"):
    """subjects: list of (subject_id, vol, mask) or (subject_id, vol, mask, truth) with truth = dict(labels,
    vol_1tp_ml, vol_2tp_ml, voxel_mm3).  Each rank processes its contiguous shard; returns {subject_id: result dict}
    (with "eval_row" when truth is given).  No collective: gather the dictionaries on the host if a global table is
    wanted."""
    res = {}
    for item in shard_items(subjects, rank, world):
        sid, vol, mask = item[0], item[1], item[2]
        seed = zlib.crc32(str(sid).encode()) & 0x7FFFFFFF  # per-subject noise stream, the same in every process
        if kind == "dem":
            res[sid] = predict_subject_dem(net, vol, mask, thr, n_repeat=n_repeat, seed=seed)
        else:
            res[sid] = predict_subject_uresnet(net, vol, mask, n_repeat=n_repeat, seed=seed)
        if len(item) > 3 and item[3] is not None:
            t = item[3]
            res[sid]["eval_row"] = evaluate_subject(res[sid], t["labels"], t["vol_1tp_ml"], t["vol_2tp_ml"],
                                                    t["voxel_mm3"], device=str(net.device))
    return res
