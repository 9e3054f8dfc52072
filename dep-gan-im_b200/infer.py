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

from . import _lib, postproc
from .api import _stream, _torch

__all__ = ["shard_range", "shard_items", "SubjectEngine", "predict_subject_dem", "predict_subject_uresnet", "evaluate_subject",
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


class SubjectEngine:
    """Per-GPU engine behind ``predict_subject_dem`` / ``predict_subject_uresnet`` / ``cohort_sweep``: the 10-repeat
    loop of the testing scripts (EG:616-628, EU:553-564) and the post-processing (EG:673-741, EU:570, 597-600) for a
    stream of subjects, double-buffered so that the upload of subject s+1 and the download of subject s-1 overlap the
    kernels of subject s (three CUDA streams, pinned host staging, no per-subject allocation).

    * Several noise repeats share one generator pass when they fit ``max_batch`` (the volume is replicated on the
      device); the float64 accumulation still adds repeat 0, 1, 2, ... in order, so results are bit-identical to the
      one-repeat-per-pass loop.
    * ``outputs`` selects what travels back to the host: any of ``"dem"``, ``"fake2"`` (float64 (Z,H,W) maps),
      ``"labels"`` (uint8) for DEP-GAN, ``"prob_mean"`` (float64 (Z,H,W,C)), ``"labels"`` for DEP-UResNet; the WMH voxel
      count always does.  The testing scripts save all of them per subject; a sweep that only needs volumes / label
      maps moves 64 KB instead of 1 MB+ per slice.
    """

    def __init__(self, net, kind="dem", z_max=48, n_repeat=10, outputs=None, depth=2):
        torch = _torch()
        self.torch, self.net, self.kind = torch, net, kind
        if kind not in ("dem", "uresnet"):
            raise ValueError("kind must be 'dem' or 'uresnet'")
        all_out = ("dem", "fake2", "labels") if kind == "dem" else ("prob_mean", "labels")
        self.outputs = tuple(all_out if outputs is None else outputs)
        for o in self.outputs:
            if o not in all_out:
                raise ValueError("unknown output %r for kind %r" % (o, kind))
        cfg, dev = net.cfg, net.device
        self.dev, self.Z, self.R, self.depth = dev, int(z_max), int(n_repeat), int(depth)
        H, W, nc = cfg.H, cfg.W, net.nc_out
        self.rpp = max(1, min(self.R, cfg.max_batch // max(1, self.Z)))   # repeats per generator pass
        self.rows = self.rpp * self.Z
        mshape = (self.Z, H, W) if kind == "dem" else (self.Z, H, W, nc)
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
            d = lambda shape, dt: [torch.empty(shape, dtype=dt, device=dev) for _ in range(depth)]  # noqa: E731
            self.x = d((self.rows, H, W, cfg.nicg), torch.float32)
            self.m = d((self.Z, H, W), torch.float32)
            self.z = d((self.R * self.Z, cfg.noise_len, 1), torch.float32)
            self.acc = d(mshape, torch.float64)
            self.o_a = d(mshape, torch.float64)          # dem / prob_mean
            self.o_b = d(mshape, torch.float64) if kind == "dem" else None   # fake2
            self.lab = d((self.Z, H, W), torch.uint8)
            self.cnt = d((1,), torch.int64)
            self.out = torch.empty((self.rows, H, W, nc), dtype=torch.float32, device=dev)
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
            self.ev_entry = torch.cuda.Event()
        p = lambda shape, dt: [torch.empty(shape, dtype=dt).pin_memory() for _ in range(depth)]  # noqa: E731
        self.hx = p((self.Z, H, W, cfg.nicg), torch.float32)
        self.hm = p((self.Z, H, W), torch.float32)
        self.hz = p((self.R * self.Z, cfg.noise_len, 1), torch.float32)
        self.h_a = p(mshape, torch.float64) if (("dem" in self.outputs) or ("prob_mean" in self.outputs)) else None
        self.h_b = p(mshape, torch.float64) if "fake2" in self.outputs else None
        self.h_lab = p((self.Z, H, W), torch.uint8) if "labels" in self.outputs else None
        self.h_cnt = p((1,), torch.int64)
        self.pending = [None] * depth
        self.i = 0
        self.h2d_bytes = self.d2h_bytes = 0

    # ---- one subject in, (possibly) one finished subject out -------------------------------------------------
    def submit(self, tag, vol, mask, thr=None, seed=0, noises=None):
        """Enqueues one subject (vol (Z,H,W,nicg), mask (Z,H,W), Z <= z_max).  Returns the finished result of the
        subject that previously used this slot (``(tag, dict)``) or None."""
        torch = self.torch
        net, cfg = self.net, self.net.cfg
        k = self.i % self.depth
        done = self._collect(k)
        vol = np.asarray(vol)
        Z = int(vol.shape[0])
        if Z > self.Z or Z < 1:
            raise ValueError("subject has %d slices, engine was built for at most %d" % (Z, self.Z))
        if self.i >= self.depth:
            self.ev_in[k].synchronize()                      # the H2D copies that last read these pinned buffers
        self.hx[k][:Z].numpy()[...] = vol
        self.hm[k][:Z].numpy()[...] = mask
        rng = np.random.default_rng(seed)
        hz = self.hz[k].numpy()
        for rep in range(self.R):                            # one fresh N(0,1) draw per repeat (EG:620 / EU:557)
            hz[rep * Z:(rep + 1) * Z] = _noise(rng, noises, rep, (Z, cfg.noise_len, 1))
        rpp = max(1, min(self.R, cfg.max_batch // Z, self.rows // Z))   # repeats sharing one generator pass
        with torch.cuda.device(self.dev):
            if self.i == 0 or not any(self.pending):
                self.ev_entry.record(torch.cuda.current_stream(self.dev))   # after the caller's weight updates
                for st in (self.s_in, self.s_run, self.s_out):
                    st.wait_event(self.ev_entry)
            with torch.cuda.stream(self.s_in):
                if self.i >= self.depth:
                    self.s_in.wait_event(self.ev_run[k])
                self.x[k][:Z].copy_(self.hx[k][:Z], non_blocking=True)
                self.m[k][:Z].copy_(self.hm[k][:Z], non_blocking=True)
                self.z[k][:self.R * Z].copy_(self.hz[k][:self.R * Z], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            self.h2d_bytes += 4 * (self.hx[k][:Z].numel() + self.hm[k][:Z].numel() + self.R * Z * cfg.noise_len)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                if self.i >= self.depth:
                    self.s_run.wait_event(self.ev_out[k])
                self._run(k, Z, rpp, thr)
                self.ev_run[k].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                if self.h_a is not None:
                    self.h_a[k][:Z].copy_(self.o_a[k][:Z], non_blocking=True)
                    self.d2h_bytes += 8 * self.o_a[k][:Z].numel()
                if self.h_b is not None:
                    self.h_b[k][:Z].copy_(self.o_b[k][:Z], non_blocking=True)
                    self.d2h_bytes += 8 * self.o_b[k][:Z].numel()
                if self.h_lab is not None:
                    self.h_lab[k][:Z].copy_(self.lab[k][:Z], non_blocking=True)
                    self.d2h_bytes += self.lab[k][:Z].numel()
                self.h_cnt[k].copy_(self.cnt[k], non_blocking=True)
                self.d2h_bytes += 8
                self.ev_out[k].record(self.s_out)
        self.pending[k] = (tag, Z)
        self.i += 1
        return done

    def _run(self, k, Z, rpp, thr):
        """The kernels of one subject on the current (run) stream."""
        torch = self.torch
        net = self.net
        x, m, acc = self.x[k], self.m[k][:Z], self.acc[k][:Z]
        for r in range(1, rpp):                              # replicate the volume for multi-repeat passes
            x[r * Z:(r + 1) * Z].copy_(x[:Z])
        acc.zero_()
        accu = postproc.DemAccumulator.wrap(acc)
        rep = 0
        while rep < self.R:
            r_now = min(rpp, self.R - rep)
            rows = r_now * Z
            _forward_volume(net, x[:rows], self.z[k][rep * Z:rep * Z + rows], self.out[:rows])
            for j in range(r_now):                           # EG:622-624: repeat order preserved
                accu.add(self.out[j * Z:(j + 1) * Z], m)
            rep += r_now
        self.cnt[k].zero_()
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            if self.kind == "dem":
                _lib.check(L.depgan_dem_postproc(x.data_ptr(), int(net.cfg.nicg), acc.data_ptr(), float(self.R),
                                                 m.data_ptr(), float(thr), self.o_a[k].data_ptr(),
                                                 self.o_b[k].data_ptr(), self.lab[k].data_ptr(),
                                                 self.cnt[k].data_ptr(), acc.numel(), _stream(torch)), "dem_postproc")
            else:
                _lib.check(L.depgan_uresnet_labels(acc.data_ptr(), float(self.R), int(net.nc_out),
                                                   self.o_a[k].data_ptr(), self.lab[k].data_ptr(),
                                                   self.cnt[k].data_ptr(), m.numel(), _stream(torch)), "uresnet_labels")

    def _collect(self, k):
        if self.pending[k] is None:
            return None
        tag, Z = self.pending[k]
        self.ev_out[k].synchronize()
        res = {"wmh_voxels": int(self.h_cnt[k].item())}
        first = "dem" if self.kind == "dem" else "prob_mean"
        if first in self.outputs:
            res[first] = self.h_a[k][:Z].numpy().copy()
        if "fake2" in self.outputs:
            res["fake2"] = self.h_b[k][:Z].numpy().copy()
        if "labels" in self.outputs:
            res["labels"] = self.h_lab[k][:Z].numpy().copy()
        self.pending[k] = None
        return tag, res

    def flush(self):
        """Finishes every subject in flight; returns their ``(tag, dict)`` results in submission order."""
        out = []
        for j in range(self.depth):
            r = self._collect((self.i + j) % self.depth)
            if r is not None:
                out.append(r)
        torch = self.torch
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_stream(self.s_run)
            cur.wait_stream(self.s_out)
        return out


def predict_subject_dem(netG, vol_1tp, mask_2tp, thr, n_repeat=10, seed=0, noises=None, outputs=None):
    """vol_1tp (Z,H,W,nicg) f32, mask_2tp (Z,H,W) f32 -> dict(dem f64, fake2 f64, labels u8, wmh_voxels int)."""
    eng = SubjectEngine(netG, "dem", z_max=int(np.shape(vol_1tp)[0]), n_repeat=n_repeat, outputs=outputs, depth=1)
    eng.submit(0, np.ascontiguousarray(vol_1tp, np.float32), np.ascontiguousarray(mask_2tp, np.float32), thr,
               seed=seed, noises=noises)
    return eng.flush()[0][1]


def predict_subject_uresnet(net, vol, mask, n_repeat=10, seed=0, noises=None, outputs=None):
    """vol (Z,H,W,1) z-scored FLAIR, mask (Z,H,W) -> dict(prob_mean f64 (Z,H,W,4), labels u8, wmh_voxels int)."""
    eng = SubjectEngine(net, "uresnet", z_max=int(np.shape(vol)[0]), n_repeat=n_repeat, outputs=outputs, depth=1)
    eng.submit(0, np.ascontiguousarray(vol, np.float32), np.ascontiguousarray(mask, np.float32), seed=seed,
               noises=noises)
    return eng.flush()[0][1]


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


def cohort_sweep(net, subjects, thr, rank=0, world=1, n_repeat=10, kind="dem", outputs=None, sink=None, engine=None):
    """subjects: list of (subject_id, vol, mask) or (subject_id, vol, mask, truth) with truth = dict(labels,
    vol_1tp_ml, vol_2tp_ml, voxel_mm3).  Each rank processes its contiguous shard through one double-buffered
    :class:`SubjectEngine`; returns {subject_id: result dict} (with "eval_row" when truth is given), or -- when
    ``sink(subject_id, result)`` is given -- hands every finished subject to the callback and returns the number of
    subjects processed (a whole cohort of float64 maps does not have to sit in host memory).  No collective: gather
    the dictionaries on the host if a global table is wanted."""
    mine = shard_items(subjects, rank, world)
    res = {}
    if not mine:
        return res if sink is None else 0
    kind_e = "dem" if kind == "dem" else "uresnet"
    if engine is None:
        z_max = max(int(np.shape(it[1])[0]) for it in mine)
        engine = SubjectEngine(net, kind_e, z_max=z_max, n_repeat=n_repeat, outputs=outputs)
    truths = {}

    def finish(done):
        if done is None:
            return
        sid, r = done
        t = truths.pop(sid, None)
        if t is not None:
            if "labels" not in r:
                raise ValueError("evaluation rows need the 'labels' output")
            r["eval_row"] = evaluate_subject(r, t["labels"], t["vol_1tp_ml"], t["vol_2tp_ml"], t["voxel_mm3"],
                                             device=str(net.device))
        if sink is None:
            res[sid] = r
        else:
            sink(sid, r)

    for item in mine:
        sid, vol, mask = item[0], item[1], item[2]
        seed = zlib.crc32(str(sid).encode()) & 0x7FFFFFFF  # per-subject noise stream, the same in every process
        if len(item) > 3 and item[3] is not None:
            truths[sid] = item[3]
        finish(engine.submit(sid, vol, mask, thr, seed=seed))
    for done in engine.flush():
        finish(done)
    return res if sink is None else len(mine)


def bind_to_gpu_numa_node(device_index):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs) and returns
    {"node": n, "cpus": k} (or None when the topology cannot be read).  One process per GPU moves its inputs and results
    through pinned host buffers (16.8 MB in + 67 MB out per 64-slice step of the Keras ``predict`` contract); pinned pages
    are placed by first touch, so without the binding every rank's buffers may land on one socket and eight ranks share
    that socket's memory controllers and inter-socket links -- the round-1 end-to-end curve (0.39 of linear at 8 GPUs).
    Call it before the first pinned allocation."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)  # not on every torch build
    except Exception:
        bus = None
    try:
        if bus is None:
            import subprocess
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                               capture_output=True, text=True, timeout=20).stdout.strip()
            addr = q.lower()[-12:] if q else None            # 00000000:1B:00.0 -> 0000:1b:00.0
        else:
            addr = "0000:%02x:00.0" % int(bus) if not isinstance(bus, str) else bus.lower()[-12:]
        if not addr:
            return None
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception:
        return None
