"""Host-side mirror of the reference's Keras surface for the DEP-GAN hot path.

    netG = Gen_UNet2D((256, 256, nicg), (32, 1), 32, 1)        # TG:520 / EG:380 / TU:579 / EU:399
    netG.load_weights('./models/netG_depgan_im_noSL_fold1.h5')  # EG:383 / EU:402
    dem = netG.predict([x, z])                                  # EG:621 / EU:558
    netD = Dis_C2D_FCN1((256, 256, 1)); netD.predict(x)         # TG:513,516,846-848

The objects own a flat float32 parameter buffer on the GPU in Keras tensor layouts and drive the CUDA kernels
through the C ABI (include/depgan_b200.h).  PyTorch is used only for device memory, streams and (in dp.py)
torch.distributed; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import synth

__all__ = ["Gen_UNet2D", "Dis_C2D_FCN1", "InferencePipeline", "conv2d_op", "wgrad_op", "launch_count"]


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("depgan_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _warn_once(obj, key, msg):
    import warnings
    seen = obj.__dict__.setdefault("_warned", set())
    if key not in seen:
        seen.add(key)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(_lib.lib().depgan_launch_count())


def manifest(model, cfg):
    """[(name 'layer/weight', shape tuple, float offset, trainable)] from the native library (host only)."""
    L = _lib.lib()
    n = L.depgan_manifest_count(model, C.byref(cfg))
    if n < 0:
        raise RuntimeError("manifest: " + _lib.last_error())
    out = []
    name = C.create_string_buffer(256)
    ndim, tr, off = C.c_int(), C.c_int(), C.c_longlong()
    shape = (C.c_int * 4)()
    for i in range(n):
        _lib.check(L.depgan_manifest_entry(model, C.byref(cfg), i, name, 256, C.byref(ndim), shape, C.byref(off),
                                           C.byref(tr)), "manifest_entry")
        out.append((name.value.decode(), tuple(shape[k] for k in range(ndim.value)), int(off.value), bool(tr.value)))
    return out, int(L.depgan_manifest_floats(model, C.byref(cfg)))


class _Net:
    """One generator or critic: flat parameters + workspace on one GPU and the native handle."""

    def __init__(self, model, cfg, device, seed, share_params_with=None):
        torch = _torch()
        self._torch = torch
        self.model, self.cfg = model, cfg
        self.device = torch.device(device)
        self.manifest, self.n_floats = manifest(model, cfg)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if share_params_with is not None:  # a second view (own workspace / batch size) of another net's weights
                if share_params_with.n_floats != self.n_floats or share_params_with.device != self.device:
                    raise ValueError("share_params_with: the networks differ")
                self.params = share_params_with.params
            else:
                self.params = torch.zeros(self.n_floats, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros(self.n_floats, dtype=torch.float32, device=self.device) if cfg.training else None
            ws = int(L.depgan_workspace_bytes(model, C.byref(cfg)))
            if ws < 0:
                raise RuntimeError("workspace_bytes: " + _lib.last_error())
            self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
            self.handle = L.depgan_net_create(model, C.byref(cfg), self.params.data_ptr(),
                                              self.grads.data_ptr() if self.grads is not None else None,
                                              self.workspace.data_ptr(), ws)
            if not self.handle:
                raise RuntimeError("net_create: " + _lib.last_error())
        self.adam_m = self.adam_v = None
        self.iterations = 0
        if share_params_with is not None:
            self.prepare()  # derived tensors of this handle; call prepare() again after the owner's weights change
            return
        man3 = [(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in self.manifest]
        self.set_weights(synth.init_weights(man3, seed=seed))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().depgan_net_destroy(self.handle)
                self.handle = None
            if getattr(self, "_peer", None):
                _lib.lib().depgan_peer_destroy(self._peer)
                self._peer = None
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------------
    def set_weights(self, weights):
        """weights: dict 'layer/weight' -> array in Keras layout.  Every manifest entry must be present."""
        torch = self._torch
        flat = np.zeros(self.n_floats, np.float32)
        for name, shape, off, _ in self.manifest:
            if name not in weights:
                raise KeyError("missing weight %s" % name)
            w = np.asarray(weights[name], np.float32)
            if tuple(w.shape) != tuple(shape):
                raise ValueError("weight %s has shape %s, expected %s" % (name, w.shape, shape))
            flat[off:off + w.size] = w.reshape(-1)
        with torch.cuda.device(self.device):
            self.params.copy_(torch.from_numpy(flat))
            self.prepare()

    def get_weights(self):
        flat = self.params.detach().cpu().numpy()
        return {name: flat[off:off + int(np.prod(shape))].reshape(shape).copy()
                for name, shape, off, _ in self.manifest}

    def get_grads(self):
        flat = self.grads.detach().cpu().numpy()
        return {name: flat[off:off + int(np.prod(shape))].reshape(shape).copy()
                for name, shape, off, _ in self.manifest}

    def count_params(self):
        return int(sum(int(np.prod(s)) for _, s, _, _ in self.manifest))

    def prepare(self):
        torch = self._torch
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_net_prepare(self.handle, _stream(torch)), "net_prepare")

    # ---- Keras-form Adam on the flat buffer (TG:549,568,594) ---------------------------------------------
    def adam_step(self, lr=1e-4, beta_1=0.0, beta_2=0.9, eps=1e-7, grad_scale=1.0):
        torch = self._torch
        with torch.cuda.device(self.device):
            if self.adam_m is None:
                self.adam_m = torch.zeros_like(self.params)
                self.adam_v = torch.zeros_like(self.params)
            self.iterations += 1
            _lib.check(_lib.lib().depgan_adam_step(self.params.data_ptr(), self.grads.data_ptr(),
                                                   self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.n_floats,
                                                   self.iterations, lr, beta_1, beta_2, eps, grad_scale,
                                                   _stream(torch)), "adam_step")
            self.prepare()

    # ---- data-parallel update through the C ABI (dp.cu): sum of the buckets over the ranks + Adam + re-pack -------
    def attach_collective(self, kind, *, comm=None, world=1, rank=0, exchange=None):
        """kind 'peer': a CUDA-IPC mailbox per rank and the fused reduce + Adam kernel (one node); `exchange(bytes) ->
        [bytes of every rank, in rank order]` is the caller's transport for the 64-byte handles.  kind 'nccl': attaches
        the communicator `comm` (an ncclComm_t as an integer, from depgan_nccl_init)."""
        L = _lib.lib()
        torch = self._torch
        with torch.cuda.device(self.device):
            if kind == "peer":
                peer = L.depgan_peer_create(self.n_floats, int(world), int(rank))
                if not peer:
                    raise RuntimeError("peer_create: " + _lib.last_error())
                buf = C.create_string_buffer(64)
                _lib.check(L.depgan_peer_handle(peer, buf), "peer_handle")
                handles = b"".join(exchange(buf.raw))
                if len(handles) != 64 * int(world):
                    raise RuntimeError("peer handle exchange returned %d bytes for %d ranks" % (len(handles), world))
                _lib.check(L.depgan_peer_connect(peer, handles), "peer_connect")
                _lib.check(L.depgan_peer_attach(self.handle, peer), "peer_attach")
                self._peer = peer
            elif kind == "nccl":
                _lib.check(L.depgan_allreduce_attach(self.handle, C.c_void_p(comm), int(world)), "allreduce_attach")
            else:
                raise ValueError("collective must be 'peer' or 'nccl'")
        self._collective = kind

    def _adam_state(self):
        torch = self._torch
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.params)
            self.adam_v = torch.zeros_like(self.params)

    def dp_update(self, lr=1e-4, beta_1=0.0, beta_2=0.9, eps=1e-7, extra=None, n_extra=0):
        """All-reduce of the gradient bucket over the attached transport + Keras Adam + prepare in one native call;
        `extra` (float64 CUDA tensor): n_extra loss partial sums reduced in the same pass, in place."""
        torch = self._torch
        with torch.cuda.device(self.device):
            self._adam_state()
            self.iterations += 1
            _lib.check(_lib.lib().depgan_dp_update(self.handle, self.adam_m.data_ptr(), self.adam_v.data_ptr(),
                                                   self.iterations, lr, beta_1, beta_2, eps,
                                                   extra.data_ptr() if extra is not None else None, int(n_extra),
                                                   _stream(torch)), "dp_update")

    def dp_allreduce_grads(self, extra=None, n_extra=0):
        torch = self._torch
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_dp_allreduce_grads(self.handle, extra.data_ptr() if extra is not None else None,
                                                            int(n_extra), _stream(torch)), "dp_allreduce_grads")

    def dp_allreduce_f64(self, buf, n):
        torch = self._torch
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_dp_allreduce_f64(self.handle, buf.data_ptr(), int(n), _stream(torch)),
                       "dp_allreduce_f64")

    def debug_activation(self, name, n):
        torch = self._torch
        cap = int(n) * self.cfg.H * self.cfg.W * 256
        buf = torch.empty(cap, dtype=torch.float32, device=self.device)
        cnt = C.c_longlong()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_debug_activation(self.handle, name.encode(), buf.data_ptr(), cap,
                                                          C.byref(cnt), int(n), _stream(torch)), "debug_activation")
        return buf[:cnt.value].cpu().numpy()

    # ---- HDF5 (Keras 2.x full-model files, EG:383 / EU:402 / TG:892 / TU:622) ----------------------------
    def load_weights(self, path):
        from . import h5lite
        self.set_weights(h5lite.load_keras_weights(path, [(n, s) for n, s, _, _ in self.manifest]))

    def _keras_description(self):
        raise NotImplementedError  # Gen_UNet2D / Dis_C2D_FCN1 name their graph

    def save(self, path, include_optimizer=True):
        """`model.save(path)` (TG:892, TU:622): a Keras 2.x full-model HDF5 file -- /model_weights in Keras' own layer order
        (weight-less layers included), the `model_config` JSON, and for a compiled model that has trained (`fit`) the
        `training_config` + /optimizer_weights (iterations, first and second moments, the unused amsgrad slots) in the
        order of `optimizer.weights`.  See keras_config.py for what is and is not verified about the JSON."""
        from . import h5lite, keras_config
        d = self._keras_description()
        w = self.get_weights()
        training_config = opt = None
        if include_optimizer and self.cfg.training == 2 and self.adam_m is not None:
            training_config = keras_config.adam_training_config()
            m, v = self.adam_m.detach().cpu().numpy(), self.adam_v.detach().cpu().numpy()
            offs = {n: (o, s) for n, s, o, tr in self.manifest if tr}
            tw = ["%s/%s" % (ln, wn) for ln in d["layer_names"] for wn in d["weights"][ln] if "%s/%s" % (ln, wn) in offs]
            opt, k = [("Adam/iterations:0", np.array(self.iterations, np.int64))], 0

            def var(arr):
                nonlocal k
                name = "training/Adam/Variable%s:0" % ("" if k == 0 else "_%d" % k)
                k += 1
                opt.append((name, arr))

            for buf in (m, v):
                for n in tw:
                    o, shp = offs[n]
                    var(buf[o:o + int(np.prod(shp))].reshape(shp).copy())
            for _ in tw:  # vhats of amsgrad=False: K.zeros(1) each
                var(np.zeros((1,), np.float32))
        h5lite.save_keras_model(path, w, d["layer_names"], d["weights"], d["model_config"], training_config, opt)

    def to_json(self):
        """`model.to_json()` (TU:623)."""
        return self._keras_description()["model_config"]

    def save_weights(self, path):
        """Weights only, under /model_weights in manifest order (readable by `load_weights` here and, by layer name, by
        Keras' `load_weights(by_name=True)`)."""
        from . import h5lite
        h5lite.save_keras_weights(path, self.get_weights(), [n for n, _, _, _ in self.manifest])


def _prec(precision):
    if precision in ("bf16", _lib.PREC_BF16):
        return _lib.PREC_BF16
    if precision in ("fp32", "f32", _lib.PREC_FP32):
        return _lib.PREC_FP32
    if precision in ("f16", "fp16", _lib.PREC_F16):
        return _lib.PREC_F16
    if precision in ("f16x3", "split", _lib.PREC_F16X3):
        return _lib.PREC_F16X3
    raise ValueError("precision must be 'bf16', 'f16' / 'f16x3' (generator inference) or 'fp32'")


class Gen_UNet2D(_Net):
    """Drop-in for ``Gen_UNet2D(input_shape, noiseZ_shape, first_fm, nc_out)`` (TG:349-498, TU:291-428).

    nc_out == 1 -> tanh head (DEP-GAN generator); nc_out == 4 -> softmax head (DEP-UResNet).
    Extra keyword arguments choose the arithmetic ('bf16' tcgen05 path; 'f16' = the same kernels with IEEE-half
    activations and weights, inference handles only, ~7x smaller DEM error at the same speed; 'f16x3' = the tensor-core
    <= 1e-4 variant: values kept as (hi, lo) half pairs, three tcgen05 products per convolution, inference handles with
    H, W multiples of 128; 'fp32' CUDA-core path),
    the workspace batch, the device and the synthetic-initialisation seed.
    """

    def __init__(self, input_shape=(256, 256, 1), noiseZ_shape=(32, 1), first_fm=32, nc_out=1, *, precision="bf16",
                 max_batch=32, device="cuda:0", training=False, seed=0, share_params_with=None):
        if first_fm != 32:
            raise ValueError("first_fm must be 32 (first_fm_G, TG:36)")
        if tuple(noiseZ_shape)[1:] != (1,):
            raise ValueError("noiseZ_shape must be (L, 1)")
        h, w, nicg = input_shape
        tmode = 2 if training in ("fit", 2) else int(bool(training))
        cfg = _lib.Cfg(int(h), int(w), int(nicg), int(nc_out), int(noiseZ_shape[0]), int(max_batch), _prec(precision),
                       tmode)
        self.input_shape, self.noiseZ_shape, self.nc_out = tuple(input_shape), tuple(noiseZ_shape), int(nc_out)
        self.precision = precision
        super().__init__(_lib.MODEL_GEN, cfg, device, seed, share_params_with)

    def _keras_description(self):
        from . import keras_config
        return keras_config.describe("generator", self.input_shape, self.noiseZ_shape[0], self.nc_out)

    def forward_device(self, x, z, out=None):
        """x (n,H,W,nicg), z (n,L,1) float32 CUDA tensors -> (n,H,W,nc_out) float32 CUDA tensor (async)."""
        torch = self._torch
        n = int(x.shape[0])
        if out is None:
            out = torch.empty((n, self.cfg.H, self.cfg.W, self.nc_out), dtype=torch.float32, device=self.device)
        if not (x.is_contiguous() and z.is_contiguous() and x.dtype == torch.float32 and z.dtype == torch.float32):
            raise ValueError("forward_device needs contiguous float32 tensors")
        if tuple(x.shape[1:]) != self.input_shape or int(z.shape[0]) != n or tuple(z.shape[1:]) != self.noiseZ_shape:
            raise ValueError("bad input shapes %s %s" % (tuple(x.shape), tuple(z.shape)))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_gen_forward(self.handle, x.data_ptr(), z.data_ptr(), out.data_ptr(), n,
                                                     _stream(torch)), "gen_forward")
        return out

    def forward_graph(self, x, z, out):
        """``forward_device`` replayed from a CUDA graph: the whole forward (28 launches with their tensor maps and
        programmatic-dependent-launch edges) is captured once per (batch size, buffer addresses) and launched as one
        graph afterwards.  Same kernels, same results bit for bit; what it removes is the per-launch host work
        (descriptor encoding, ``cudaLaunchKernelEx``), which matters for small batches where the GPU outruns the host.
        Weight updates need no re-capture (the derived weight buffers keep their addresses)."""
        torch = self._torch
        key = (int(x.shape[0]), x.data_ptr(), z.data_ptr(), out.data_ptr())
        cache = self.__dict__.setdefault("_graphs", {})
        g = cache.get(key)
        if g is None:
            if len(cache) >= 16:   # callers that keep allocating new buffers would grow the cache without bound
                cache.clear()
            with torch.cuda.device(self.device):
                cur = torch.cuda.current_stream(self.device)
                side = torch.cuda.Stream(self.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self.forward_device(x, z, out)   # warm-up outside the capture (one-time set-up paths)
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self.forward_device(x, z, out)
                cur.wait_stream(side)
            cache[key] = g
        g.replay()
        return out

    # ---- DEP-UResNet supervised training: my_network.fit(...) TU:602-606 (model compiled at TU:427) ----------
    def enable_data_parallel(self):
        """Data-parallel ``fit`` under torchrun (one process per GPU, ``torch.distributed`` initialised with NCCL):
        BatchNorm statistics become batch-global (synchronised BN: the C library calls back into an all-reduce for every
        BN layer, forward and backward), the loss is scaled by the global pixel count, and ``train_on_batch_device``
        sums the gradient bucket and the loss over the ranks, so every rank applies the update a single GPU would apply
        to the concatenated batch."""
        torch = self._torch
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            raise RuntimeError("enable_data_parallel needs an initialised torch.distributed group with world size >= 2")
        ws = self.workspace
        base, nbytes = ws.data_ptr(), ws.numel()

        def hook(user, ptr, count, is_f64, stream):
            try:
                off, size = ptr - base, count * (8 if is_f64 else 4)
                if off < 0 or off + size > nbytes:
                    return -1
                dist.all_reduce(ws[off:off + size].view(torch.float64 if is_f64 else torch.float32))
                return 0
            except Exception:  # never let an exception cross the C boundary
                return -1

        self._sync_cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p)(hook)
        self._dist = dist
        _lib.check(_lib.lib().depgan_set_sync_hook(C.cast(self._sync_cb, C.c_void_p), None, dist.get_world_size()),
                   "set_sync_hook")

    def train_on_batch_device(self, x, z, onehot, keep):
        """One Keras training-phase step on CUDA tensors (x (n,H,W,1) f32, z (n,L,1) f32, onehot (n,H,W,nc) f32,
        keep (n,H/4,W/4,96) uint8 Dropout keep mask).  Gradients + Adam(1e-4, 0.9, 0.999); returns the loss tensor."""
        torch = self._torch
        if self.cfg.training != 2:
            raise ValueError("create the model with training='fit' to use fit / train_on_batch")
        n = int(x.shape[0])
        if not hasattr(self, "_loss"):
            self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_uresnet_grads(self.handle, x.data_ptr(), z.data_ptr(), onehot.data_ptr(),
                                                       keep.data_ptr(), self._loss.data_ptr(), n, _stream(torch)),
                       "uresnet_grads")
        if getattr(self, "_dist", None) is not None:  # partial sums of the global-batch mean
            self._dist.all_reduce(self.grads)
            self._dist.all_reduce(self._loss)
        loss = self._loss.clone()
        self.adam_step(1e-4, 0.9, 0.999)  # keras.optimizers.Adam(lr=1e-4) defaults (TU:427)
        return loss

    def evaluate_loss(self, x, z, onehot, batch_size=16):
        """Mean categorical cross-entropy in inference mode (Keras validation_data handling, TU:606)."""
        torch = self._torch
        n = x.shape[0]
        tot = torch.zeros(1, dtype=torch.float32, device=self.device)
        npix_total = float(n * self.cfg.H * self.cfg.W)
        for i in range(0, n, batch_size):
            xb = torch.from_numpy(np.ascontiguousarray(x[i:i + batch_size], np.float32)).to(self.device)
            zb = torch.from_numpy(np.ascontiguousarray(z[i:i + batch_size], np.float32)).to(self.device)
            tb = torch.from_numpy(np.ascontiguousarray(onehot[i:i + batch_size], np.float32)).to(self.device)
            prob = self.forward_device(xb, zb)
            scratch = torch.empty_like(prob)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().depgan_cce_loss(prob.data_ptr(), tb.data_ptr(), scratch.data_ptr(),
                                                      tot.data_ptr(), prob.numel() // self.nc_out, self.nc_out,
                                                      1.0 / npix_total, _stream(torch)), "cce_loss")
        return float(tot.item())

    def fit(self, inputs, onehot, epochs=1, batch_size=16, shuffle=True, validation_data=None, seed=0, verbose=0):
        """Keras ``Model.fit([x, z], y, epochs, batch_size, shuffle, validation_data)`` (TU:602-606).  Returns an
        object with ``.history = {"loss": [...], "val_loss": [...]}`` (one entry per epoch; the epoch loss is the
        sample-weighted mean of the batch losses, as Keras reports).  The shuffle order and the Dropout keep masks
        come from ``numpy.random.default_rng(seed)`` (the reference draws both unseeded)."""
        torch = self._torch
        x, z = inputs
        x = np.ascontiguousarray(x, np.float32)
        z = np.ascontiguousarray(z, np.float32)
        y = np.ascontiguousarray(onehot, np.float32)
        n = x.shape[0]
        rng = np.random.default_rng(seed)
        hist = {"loss": [], "val_loss": []}
        dp = getattr(self, "_dist", None)
        rank, world = (dp.get_rank(), dp.get_world_size()) if dp is not None else (0, 1)
        bs = int(batch_size) if dp is not None else min(int(batch_size), self.cfg.max_batch)
        for _ in range(int(epochs)):
            order = rng.permutation(n) if shuffle else np.arange(n)
            tot, cnt = 0.0, 0
            for i in range(0, n, bs):
                idx = order[i:i + bs]
                if dp is not None and len(idx) % world:
                    # remainder batch under data parallelism: every rank needs the same number of rows for the
                    # synchronised statistics, so the tail that does not divide evenly is left out (at most
                    # world - 1 samples per epoch); Keras on one device trains on all of them
                    cut = len(idx) - len(idx) % world
                    _warn_once(self, "fit_dp_tail", "fit: %d sample(s) of the last batch dropped per epoch (batch not "
                               "divisible by the %d data-parallel ranks)" % (len(idx) - cut, world))
                    idx = idx[:cut]
                if len(idx) < 2 * world:
                    # training-phase BatchNorm cannot normalise a single sample per rank (its variance is 0): the one
                    # deviation from Keras Model.fit, which would train on it
                    if len(idx):
                        _warn_once(self, "fit_single", "fit: a final batch of %d sample(s) is skipped (batch "
                                   "statistics need two samples per rank)" % len(idx))
                    continue
                keep = (rng.uniform(size=(len(idx), self.cfg.H // 4, self.cfg.W // 4, 96)) >= 0.25).astype(np.uint8)
                if dp is not None:  # every rank draws the same global batch and mask and takes its contiguous shard
                    per = len(idx) // world
                    if per > self.cfg.max_batch:
                        raise ValueError("batch_size / world exceeds max_batch")
                    idx, keep = idx[rank * per:(rank + 1) * per], keep[rank * per:(rank + 1) * per]
                loss = self.train_on_batch_device(torch.from_numpy(x[idx]).to(self.device),
                                                  torch.from_numpy(z[idx]).to(self.device),
                                                  torch.from_numpy(y[idx]).to(self.device),
                                                  torch.from_numpy(keep).to(self.device))
                tot += float(loss.item()) * len(idx) * world
                cnt += len(idx) * world
            hist["loss"].append(tot / max(cnt, 1))
            if validation_data is not None:
                (xv, zv), yv = validation_data
                hist["val_loss"].append(self.evaluate_loss(xv, zv, yv, bs))

        class History:
            pass
        h = History()
        h.history = hist
        return h

    def cast_output(self, src, dst):
        """float32 CUDA tensor -> float16 / bfloat16 CUDA tensor of the same shape on the current stream (our cast
        kernel; the opt-in narrow ``predict`` output)."""
        torch = self._torch
        if src.dtype != torch.float32 or src.numel() != dst.numel() or not (src.is_contiguous() and dst.is_contiguous()):
            raise ValueError("cast_output needs contiguous tensors of equal size, float32 source")
        L = _lib.lib()
        fn = {torch.bfloat16: L.depgan_op_f32_to_bf16, torch.float16: L.depgan_op_f32_to_f16}.get(dst.dtype)
        if fn is None:
            raise ValueError("cast_output: destination must be float16 or bfloat16")
        with torch.cuda.device(self.device):
            _lib.check(fn(src.data_ptr(), dst.data_ptr(), src.numel(), _stream(torch)), "cast_output")
        return dst

    def predict(self, inputs, batch_size=32, verbose=0, out_dtype=None):
        """Keras ``model.predict([x, z])`` (EG:621, EU:558): numpy in, numpy float32 out, inference mode.

        Batches of ``batch_size`` (clamped to ``max_batch``) run through a persistent :class:`InferencePipeline` with
        pinned staging buffers owned by the model: no per-call pinning, copies overlap the kernels, and results do not
        depend on the batching (slices are independent in inference mode).  ``out_dtype=np.float16`` is an opt-in
        extension: the maps cross PCIe as float16 (half the device->host bytes) and are returned as float16."""
        torch = self._torch
        x, z = inputs
        x = np.ascontiguousarray(x, np.float32)
        z = np.ascontiguousarray(z, np.float32)
        if x.shape[0] != z.shape[0]:
            raise ValueError("x and z must have the same number of samples")
        if tuple(x.shape[1:]) != self.input_shape or tuple(z.shape[1:]) != self.noiseZ_shape:
            raise ValueError("bad input shapes %s %s" % (x.shape, z.shape))
        n = x.shape[0]
        bs = max(1, min(int(batch_size), self.cfg.max_batch))
        np_dt = np.dtype(out_dtype or np.float32)
        if np_dt not in (np.dtype(np.float32), np.dtype(np.float16)):
            raise ValueError("out_dtype must be float32 or float16")
        t_dt = torch.float32 if np_dt == np.float32 else torch.float16
        pipes = self.__dict__.setdefault("_pipes", {})
        pipe = pipes.get(t_dt)
        if pipe is None:
            pipe = pipes[t_dt] = InferencePipeline(self, depth=2, out_dtype=t_dt, staging=True)
        # (Measured and dropped: allocating the result in page-locked memory so that the device->host copies land in it
        # directly -- 5.5 k slices/s against 7.5 k with the staged copy below; the 268 MB pinned allocation per call costs
        # more than the host copy it saves.)
        out = np.empty((n, self.cfg.H, self.cfg.W, self.nc_out), np_dt)
        for i in range(0, n, bs):
            pipe.submit_numpy(x[i:i + bs], z[i:i + bs], out[i:i + bs])
        pipe.flush()
        return out


_copy_pool = None


def _host_copy(dst, src):
    """dst[...] = src for large host arrays, split along axis 0 over a few threads (NumPy releases the GIL while it
    copies).  One thread moves ~5 GB/s, and predict() on pageable arrays moves 84 MB per 64-slice batch through the
    pinned staging buffers: the single-threaded copies, not the GPU, bounded it (3.8 k slices/s)."""
    global _copy_pool
    n = dst.shape[0]
    if dst.nbytes < (8 << 20) or n < 4:
        np.copyto(dst, src, casting="same_kind")
        return
    if _copy_pool is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 4) // 2)))
    k = _copy_pool._max_workers
    step = (n + k - 1) // k
    futs = [_copy_pool.submit(np.copyto, dst[i:i + step], src[i:i + step], "same_kind") for i in range(0, n, step)]
    for f in futs:
        f.result()


class InferencePipeline:
    """Double-buffered host<->device pipeline around ``Gen_UNet2D.forward_device`` for streams of batches that
    live in pinned host memory: the H2D copy of batch i+1 and the D2H copy of batch i-1 overlap the kernels of
    batch i (three CUDA streams, events instead of host synchronisation).

        pipe = InferencePipeline(netG)
        for xh, zh, oh in batches:            # pinned host tensors; oh receives the prediction
            pipe.submit(xh, zh, oh)
        pipe.flush()                          # all outputs are complete in host memory

    The pipeline's streams are ordered after whatever the caller's current stream had enqueued at the first
    ``submit`` of a cycle (``load_weights`` / ``set_weights`` / ``adam_step`` re-pack the weights there), and
    ``flush`` makes the caller's stream wait for the last copy.

    ``out_dtype``: ``torch.float32`` (the Keras ``predict`` contract) or ``torch.float16`` / ``torch.bfloat16`` -- an
    opt-in narrower device->host transfer (the network output is converted on the device by our cast kernel; half the
    PCIe bytes per slice).  ``staging=True`` adds persistent pinned host buffers so that pageable NumPy arrays can be
    fed with ``submit_numpy`` without a ``pin_memory()`` allocation per batch."""

    def __init__(self, net, depth=2, out_dtype=None, staging=False):
        torch = net._torch
        self.net, self.torch, self.depth = net, torch, depth
        dev, cfg = net.device, net.cfg
        B = cfg.max_batch
        self.out_dtype = out_dtype or torch.float32
        if self.out_dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise ValueError("out_dtype must be float32, float16 or bfloat16")
        with torch.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
            self.x = [torch.empty((B, cfg.H, cfg.W, cfg.nicg), dtype=torch.float32, device=dev) for _ in range(depth)]
            self.z = [torch.empty((B, cfg.noise_len, 1), dtype=torch.float32, device=dev) for _ in range(depth)]
            self.o = [torch.empty((B, cfg.H, cfg.W, net.nc_out), dtype=torch.float32, device=dev) for _ in range(depth)]
            self.o_lp = None
            if self.out_dtype != torch.float32:
                self.o_lp = [torch.empty((B, cfg.H, cfg.W, net.nc_out), dtype=self.out_dtype, device=dev)
                             for _ in range(depth)]
            self.ev_in = [torch.cuda.Event() for _ in range(depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(depth)]
            self.ev_out = [torch.cuda.Event() for _ in range(depth)]
            self.ev_entry = torch.cuda.Event()
        self.i = 0
        self._open = False  # a submit/flush cycle is in progress
        self.hx = self.hz = self.ho = None
        if staging:
            self.hx = [torch.empty((B, cfg.H, cfg.W, cfg.nicg), dtype=torch.float32).pin_memory() for _ in range(depth)]
            self.hz = [torch.empty((B, cfg.noise_len, 1), dtype=torch.float32).pin_memory() for _ in range(depth)]
            self.ho = [torch.empty((B, cfg.H, cfg.W, net.nc_out), dtype=self.out_dtype).pin_memory()
                       for _ in range(depth)]
            self._pending = [None] * depth  # (destination numpy array, n) whose D2H copy into ho[k] is in flight

    def _enter(self):
        # first submit of a cycle: everything the caller's stream has enqueued so far (weight fold / pack kernels of
        # depgan_net_prepare in particular) happens before the pipeline touches the network
        if not self._open:
            torch = self.torch
            cur = torch.cuda.current_stream(self.net.device)
            self.ev_entry.record(cur)
            for s in (self.s_in, self.s_run, self.s_out):
                s.wait_event(self.ev_entry)
            self._open = True

    def submit(self, xh, zh, oh):
        torch = self.torch
        k = self.i % self.depth
        n = int(xh.shape[0])
        if oh.dtype != self.out_dtype:
            raise ValueError("output buffer dtype %s does not match the pipeline's %s" % (oh.dtype, self.out_dtype))
        with torch.cuda.device(self.net.device):
            self._enter()
            with torch.cuda.stream(self.s_in):
                if self.i >= self.depth:
                    self.s_in.wait_event(self.ev_run[k])   # the kernels that last read these input buffers
                self.x[k][:n].copy_(xh, non_blocking=True)
                self.z[k][:n].copy_(zh, non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                if self.i >= self.depth:
                    self.s_run.wait_event(self.ev_out[k])  # the D2H copy that last read this output buffer
                self.net.forward_device(self.x[k][:n], self.z[k][:n], self.o[k][:n])
                src = self.o[k]
                if self.o_lp is not None:                  # opt-in narrow output: converted on the device
                    self.net.cast_output(self.o[k][:n], self.o_lp[k][:n])
                    src = self.o_lp[k]
                self.ev_run[k].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                oh.copy_(src[:n], non_blocking=True)
                self.ev_out[k].record(self.s_out)
        self.i += 1

    def _drain(self, k):
        """Persistent-staging mode: wait for the D2H copy into ho[k] and move it to its NumPy destination."""
        if self._pending[k] is not None:
            dst, n = self._pending[k]
            self.ev_out[k].synchronize()
            if self.out_dtype == self.torch.bfloat16:
                dst[...] = self.ho[k][:n].float().numpy()
            else:
                _host_copy(dst, self.ho[k][:n].numpy())
            self._pending[k] = None

    def submit_numpy(self, x, z, out):
        """Pageable NumPy in / out through the persistent pinned staging buffers (``staging=True``): ``x``/``z`` are
        copied into pinned slot k (host memcpy), ``out`` (a NumPy view) is filled once the slot's D2H copy has
        completed -- at the latest in ``flush``."""
        if self.hx is None:
            raise RuntimeError("create the pipeline with staging=True to use submit_numpy")
        k = self.i % self.depth
        n = int(x.shape[0])
        self._drain(k)                    # the slot's previous result must have left ho[k]
        if self.i >= self.depth:
            self.ev_in[k].synchronize()   # the H2D copy that last read hx[k] / hz[k]
        _host_copy(self.hx[k][:n].numpy(), x)
        self.hz[k][:n].numpy()[...] = z
        if isinstance(out, self.torch.Tensor):   # a page-locked destination: the D2H copy goes straight into it
            self.submit(self.hx[k][:n], self.hz[k][:n], out)
        else:
            self.submit(self.hx[k][:n], self.hz[k][:n], self.ho[k][:n])
            self._pending[k] = (out, n)

    def flush(self):
        torch = self.torch
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()
        if self.hx is not None:
            for k in range(self.depth):
                self._drain(k)
        # later work on the caller's stream (e.g. a weight update) is ordered after the pipeline's last kernel / copy
        with torch.cuda.device(self.net.device):
            cur = torch.cuda.current_stream(self.net.device)
            cur.wait_stream(self.s_run)
            cur.wait_stream(self.s_out)
        self._open = False


class Dis_C2D_FCN1(_Net):
    """Drop-in for ``Dis_C2D_FCN1(input_shape)`` (TG:316-345): the WGAN-GP critic, (N,H,W,1) -> (N,1)."""

    def __init__(self, input_shape=(256, 256, 1), *, precision="bf16", max_batch=32, device="cuda:0", training=False,
                 seed=1, share_params_with=None):
        h, w, c = input_shape
        if c != 1:
            raise ValueError("the critic takes one channel (TG:513,516)")
        cfg = _lib.Cfg(int(h), int(w), 1, 1, 32, int(max_batch), _prec(precision), int(bool(training)))
        self.input_shape = tuple(input_shape)
        self.precision = precision
        super().__init__(_lib.MODEL_CRITIC, cfg, device, seed, share_params_with)

    def _keras_description(self):
        from . import keras_config
        return keras_config.describe("critic", self.input_shape)

    def forward_device(self, x, out=None):
        torch = self._torch
        n = int(x.shape[0])
        if out is None:
            out = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        if not (x.is_contiguous() and x.dtype == torch.float32) or tuple(x.shape[1:]) != self.input_shape:
            raise ValueError("forward_device needs a contiguous float32 (n,H,W,1) tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().depgan_critic_forward(self.handle, x.data_ptr(), out.data_ptr(), n, _stream(torch)),
                       "critic_forward")
        return out

    def predict(self, x, batch_size=32, verbose=0):
        torch = self._torch
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[0]
        bs = max(1, min(int(batch_size), self.cfg.max_batch))
        out = np.empty((n, 1), np.float32)
        for i in range(0, n, bs):
            out[i:i + bs] = self.forward_device(torch.from_numpy(x[i:i + bs]).to(self.device)).cpu().numpy()
        return out


# ---- kernel-level op (tests / micro-benchmarks) -----------------------------------------------------------
def conv2d_op(x, w, *, x1=None, scale=None, shift=None, relu=False, film=None, res=None, add=None, mask=None,
              deconv=False, head=None, use_tc=True, want_pre=False, want_pool=False):
    """One fused convolution through depgan_op_conv2d.

    x (N,H,W,C0) [, x1 (N,H,W,C1)] float32 CUDA tensors; w Keras HWIO (k,k,Cin,Cout) or, for deconv, Keras
    Conv2DTranspose layout (2,2,Cout,Cin).  use_tc=True runs the tcgen05 path on bf16 copies, else the fp32
    CUDA-core path.  film = (gamma (N,C), beta (N,C)); head = (w (C,nc), b (nc), act).  Returns float32 tensors:
    out, or (out, extras dict) when head / want_pre are used.
    """
    torch = _torch()
    L = _lib.lib()
    st = _stream(torch)
    dev = x.device
    N, H, W, C0 = x.shape
    C1 = 0 if x1 is None else x1.shape[3]
    if deconv:
        cout, cin, ks, taps = w.shape[2], w.shape[3], 1, 1
        ncols = 4 * cout
        oh, ow = 2 * H, 2 * W
    else:
        ks, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        taps, ncols = ks * ks, cout
        oh, ow = H, W
    assert cin == C0 + C1
    keep = []

    def dev_f32(t):
        t = t.to(dev, torch.float32).contiguous()
        keep.append(t)
        return t

    def to_act(t):
        t = dev_f32(t)
        if not use_tc:
            return t
        b = torch.empty(t.numel(), dtype=torch.bfloat16, device=dev)
        _lib.check(L.depgan_op_f32_to_bf16(t.data_ptr(), b.data_ptr(), t.numel(), st), "f32_to_bf16")
        keep.append(b)
        return b

    d = _lib.ConvDesc()
    d.in0 = to_act(x).data_ptr()
    d.C0, d.C1 = C0, C1
    if x1 is not None:
        d.in1 = to_act(x1).data_ptr()
    wf = dev_f32(w)
    if use_tc:
        wb = torch.empty(wf.numel(), dtype=torch.bfloat16, device=dev)
        if deconv:  # (2,2,Cout,Cin) is already [4*Cout][Cin]
            _lib.check(L.depgan_op_f32_to_bf16(wf.data_ptr(), wb.data_ptr(), wf.numel(), st), "f32_to_bf16")
        else:
            _lib.check(L.depgan_op_pack_weights(wf.data_ptr(), wb.data_ptr(), taps, cin, cout, st), "pack_weights")
        keep.append(wb)
        d.w_bf16 = wb.data_ptr()
    else:
        assert not deconv, "the fp32 path runs the transposed conv through its own kernel"
        d.w_f32 = wf.data_ptr()
    if scale is not None:
        d.scale = dev_f32(scale).data_ptr()
    if shift is not None:
        d.shift = dev_f32(shift).data_ptr()
    odt = torch.bfloat16 if use_tc else torch.float32
    out = torch.empty((N, oh, ow, cout), dtype=odt, device=dev)
    d.out = out.data_ptr()
    pre = None
    if want_pre:
        pre = torch.empty((N, oh, ow, cout), dtype=odt, device=dev)
        d.out_pre = pre.data_ptr()
    if film is not None:
        g, b = dev_f32(film[0]), dev_f32(film[1])
        d.film_g, d.film_b, d.film_stride = g.data_ptr(), b.data_ptr(), g.shape[1]
        d.res = to_act(res).data_ptr()
    if add is not None:
        d.add_src = to_act(add).data_ptr()
    if mask is not None:
        d.mask_src = to_act(mask).data_ptr()
    d.relu, d.deconv = int(relu), int(deconv)
    hout = None
    if head is not None:
        hw, hb, act = head
        hw, hb = dev_f32(hw), dev_f32(hb)
        hout = torch.empty((N, H, W, hw.shape[1]), dtype=torch.float32, device=dev)
        d.head_w, d.head_b, d.head_out = hw.data_ptr(), hb.data_ptr(), hout.data_ptr()
        d.head_nc, d.head_act = hw.shape[1], int(act)
    d.N, d.H, d.W, d.Cout, d.ks = N, H, W, cout, ks
    d.in_bf16 = d.out_bf16 = int(use_tc)
    d.use_tc = int(use_tc)
    pool = None
    if want_pool:
        pool = torch.empty((N, oh // 2, ow // 2, cout), dtype=odt, device=dev)
        d.pool_out = pool.data_ptr()
    _lib.check(L.depgan_op_conv2d(C.byref(d), st), "op_conv2d")
    torch.cuda.synchronize(dev)
    outf = out.float()
    if head is None and not want_pre and not want_pool:
        return outf
    return outf, {"head": hout, "pre": None if pre is None else pre.float(),
                  "pool": None if pool is None else pool.float()}


def wgrad_op(x, dy, ks, *, x1=None, use_tc=True):
    """Weight gradient of a 'same' stride-1 convolution through depgan_op_wgrad.
    x (N,H,W,C0) [, x1 (N,H,W,C1)], dy (N,H,W,Cout) float32 CUDA tensors -> dw (ks,ks,C0+C1,Cout) float32."""
    torch = _torch()
    L = _lib.lib()
    st = _stream(torch)
    dev = x.device
    N, H, W, C0 = x.shape
    C1 = 0 if x1 is None else x1.shape[3]
    cout = dy.shape[3]
    keep = []

    def conv(t):
        t = t.to(dev, torch.float32).contiguous()
        keep.append(t)
        if not use_tc:
            return t
        b = torch.empty(t.numel(), dtype=torch.bfloat16, device=dev)
        _lib.check(L.depgan_op_f32_to_bf16(t.data_ptr(), b.data_ptr(), t.numel(), st), "f32_to_bf16")
        keep.append(b)
        return b

    xa, da = conv(x), conv(dy)
    x1a = conv(x1) if x1 is not None else None
    dw = torch.zeros((ks, ks, C0 + C1, cout), dtype=torch.float32, device=dev)
    _lib.check(L.depgan_op_wgrad(xa.data_ptr(), None if x1a is None else x1a.data_ptr(), C0, C1, da.data_ptr(),
                                 dw.data_ptr(), N, H, W, cout, ks, int(use_tc), st), "op_wgrad")
    torch.cuda.synchronize(dev)
    return dw
