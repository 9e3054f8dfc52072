"""The four ``K.function`` callables of the DEP-GAN training script (TG:550-552, 569-571, 595-598) and the step
schedule around them (TG:780-878), driving the CUDA train-step graphs through the C ABI.

    tr = DepGanTrainer(netG, netD_y2, netD_dem, IM_TRSH)
    loss_real, loss_fake = tr.netD_y2_train([real_2tp, real_1tp, noise, ep])        # TG:809
    loss_real, loss_fake = tr.netD_dem_train([real_2tp, real_1tp, noise, ep])       # TG:824
    loss, lf, lfd, M1, M3, M4 = tr.netG_no_update([real_1tp, real_2tp, noise])      # TG:873
    loss, lf, lfd, M1, M3, M4 = tr.netG_train([real_1tp, real_2tp, noise])          # TG:878

Argument orders follow the reference (critics: [real_2tp, real_1tp, noise, ep]; generator: [real_1tp, real_2tp,
noise]).  Returned values are computed from the pre-update weights; each ``*_train`` call performs exactly one
Keras-form Adam step (lr 1e-4, beta_1 0, beta_2 0.9; TG:549,568,594) on exactly one network.

Data parallel: when ``torch.distributed`` is initialised with world size W, every rank passes its
shard of the global batch; the flat gradient buckets and the loss partial sums (incl. the batch-global dice / volume
terms) are summed over the ranks by the native library (``depgan_dp_update``: a fused reduce + Adam kernel over
CUDA-IPC peer memory on one node, or NCCL), so every rank reports the global-batch values and applies the identical
update; torch.distributed only carries the 64-byte handles / the NCCL id at start-up.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .api import _stream, _torch

__all__ = ["DepGanTrainer", "ScalarLog", "epoch_schedule"]


def _find_nccl():
    """Path of the NCCL library PyTorch ships (the wheel's nvidia/nccl/lib), or None to let dlopen search."""
    try:
        import nvidia.nccl as _n
        from pathlib import Path
        for p in Path(_n.__path__[0]).glob("lib/libnccl.so*"):
            return str(p).encode()
    except Exception:
        pass
    return None


def epoch_schedule(batches, gen_iterations, Diters=5):
    """The order in which one epoch of the reference loop (TG:790-894) touches its mini-batches, as a generator of
    ``("y2", i)``, ``("dem", ii)`` and ``("gen", b, gen_iterations)`` events: per generator iteration up to ``_Diters``
    Y2-critic updates on batches i, i+1, ... (``_Diters`` = 100 while ``gen_iterations < 25`` or every 500th iteration,
    else ``Diters``; TG:792-797), up to ``_Diters`` DEM-critic updates on its own cursor ii (TG:814-829), then the noise
    search and ONE generator update on the batch the critics saw last (index ``b``; TG:867-878)."""
    i = ii = 0
    last = None
    while i < batches:
        d = 100 if (gen_iterations < 25 or gen_iterations % 500 == 0) else Diters
        j = jj = 0
        while j < d and i < batches:
            j += 1
            last = i
            yield ("y2", i)
            i += 1
        while jj < d and ii < batches:
            jj += 1
            last = ii
            yield ("dem", ii)
            ii += 1
        yield ("gen", last, gen_iterations)
        gen_iterations += 1


class ScalarLog:
    """In-memory / CSV variant of the reference's TensorBoard ``Logger`` (TG:167-248): ``log_scalar(tag, value, step)``
    keeps the series in memory (``.series[tag] = [(step, value), ...]``) and, if a path is given, appends CSV lines.
    ``tblog.TensorBoardLogger`` has the same interface and writes real event files (scalars, images, histograms)."""

    def __init__(self, csv_path=None):
        self.series = {}
        self.csv_path = csv_path

    def log_scalar(self, tag, value, step):
        self.series.setdefault(tag, []).append((int(step), float(value)))
        if self.csv_path:
            with open(self.csv_path, "a") as f:
                f.write("%s,%d,%.9g\n" % (tag, int(step), float(value)))

    def log_images(self, tag, images, step, *args):  # image summaries are not kept
        self.series.setdefault(tag, []).append((int(step), float(np.asarray(images).shape[0])))


class DepGanTrainer:
    def __init__(self, netG, netD_y2, netD_dem, thr, lrD=1e-4, lrG=1e-4, beta_1=0.0, beta_2=0.9, distributed=None,
                 collective=None):
        torch = _torch()
        self.torch = torch
        self.G, self.Dy2, self.Ddem = netG, netD_y2, netD_dem
        for net in (netG, netD_y2, netD_dem):
            if not net.cfg.training:
                raise ValueError("networks must be created with training=True")
        self.thr = float(np.float32(thr))
        self.lrD, self.lrG, self.b1, self.b2 = lrD, lrG, beta_1, beta_2
        self.device = netG.device
        dist = torch.distributed
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.dist = dist if distributed else None
        self.world = dist.get_world_size() if distributed else 1
        # Transport of the gradient sum: 'peer' = the native fused reduce + Adam kernel over CUDA-IPC
        # mailboxes (one node, the default), 'nccl' = ncclAllReduce issued by the native library on its own communicator,
        # 'torch' = torch.distributed.all_reduce + separate Adam (the round-1 path, kept for A/B measurements).
        self.collective = None
        if distributed:
            self.collective = collective or os.environ.get("DEPGAN_COLLECTIVE", "peer")
            self._attach_collective()
        self.batched_eval = False
        self.ex64 = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.out4 = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.out6 = torch.zeros(6, dtype=torch.float32, device=self.device)
        self.sums = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.gen_iterations = 0

    def _attach_collective(self):
        dist, rank = self.dist, self.dist.get_rank()
        nets = (self.G, self.Dy2, self.Ddem)
        if self.collective == "peer":
            def exchange(mine):
                got = [None] * self.world
                dist.all_gather_object(got, mine)
                return got
            for net in nets:
                net.attach_collective("peer", world=self.world, rank=rank, exchange=exchange)
            dist.barrier()
        elif self.collective == "nccl":
            L = _lib.lib()
            _lib.check(L.depgan_nccl_load(_find_nccl()), "nccl_load")
            ident = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                _lib.check(L.depgan_nccl_unique_id(buf), "nccl_unique_id")
                ident = [buf.raw]
            dist.broadcast_object_list(ident, src=0)
            comm = C.c_void_p()
            with self.torch.cuda.device(self.device):
                _lib.check(L.depgan_nccl_init(C.byref(comm), self.world, rank, ident[0]), "nccl_init")
            self._nccl_comm = comm
            for net in nets:
                net.attach_collective("nccl", comm=comm.value, world=self.world)
        elif self.collective != "torch":
            raise ValueError("collective must be 'peer', 'nccl' or 'torch'")

    def _update(self, net, lr, extra=None, n_extra=0):
        """One optimizer step of `net` on the global-batch gradient."""
        if self.collective in ("peer", "nccl"):
            net.dp_update(lr, self.b1, self.b2, extra=extra, n_extra=n_extra)
        else:
            net.adam_step(lr, self.b1, self.b2)

    # ---- helpers -------------------------------------------------------------------------------------
    def _dev(self, a, dtype=None):
        torch = self.torch
        if isinstance(a, torch.Tensor):
            t = a.to(self.device, torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(self.device)
        return t.contiguous()

    def _allreduce_grads(self, net):
        if self.dist is not None:
            self.dist.all_reduce(net.grads, op=self.dist.ReduceOp.SUM)

    # ---- critics (TG:523-571) -------------------------------------------------------------------------
    def critic_grads_device(self, which, real2, x1, z, ep, reduce=True, dem=None):
        """Device tensors in; leaves dLoss/dtheta in the critic's flat gradient buffer and returns the float32
        CUDA tensor [loss_real, loss_fake, gradient_penalty, loss] for the GLOBAL batch."""
        torch = self.torch
        D = self.Dy2 if which == 0 else self.Ddem
        n = int(real2.shape[0])
        if 3 * n > D.cfg.max_batch:
            raise ValueError("the critic needs max_batch >= 3 * batch (real | fake | mixed rows)")
        with torch.cuda.device(self.device):
            if dem is not None:  # G(x1, z) computed ahead in a batched pass (gen_iteration_device)
                _lib.check(_lib.lib().depgan_critic_grads_dem(D.handle, self.G.cfg.nicg, which, real2.data_ptr(),
                                                              x1.data_ptr(), dem.data_ptr(), ep.data_ptr(),
                                                              self.out4.data_ptr(), n, n * self.world, _stream(torch)),
                           "critic_grads_dem")
            else:
                _lib.check(_lib.lib().depgan_critic_grads(D.handle, self.G.handle, which, real2.data_ptr(),
                                                          x1.data_ptr(), z.data_ptr(), ep.data_ptr(),
                                                          self.out4.data_ptr(), n, n * self.world, _stream(torch)),
                           "critic_grads")
        if self.collective == "torch":
            self.dist.all_reduce(self.out4, op=self.dist.ReduceOp.SUM)
            self._allreduce_grads(D)
        elif self.dist is not None and reduce:
            self.ex64[:4].copy_(self.out4)
            D.dp_allreduce_grads(self.ex64, 4)
            self.out4.copy_(self.ex64[:4])
        return self.out4

    def _critic_update_native(self, D):
        """Native transports: the loss partial sums ride along with the gradient bucket of the fused update."""
        self.ex64[:4].copy_(self.out4)
        D.dp_update(self.lrD, self.b1, self.b2, extra=self.ex64, n_extra=4)
        self.out4.copy_(self.ex64[:4])

    def _critic_train(self, which, inputs, update=True):
        real2, x1, z, ep = [self._dev(a) for a in inputs]
        ep = ep.reshape(-1).contiguous()
        D = self.Dy2 if which == 0 else self.Ddem
        native = self.collective in ("peer", "nccl")
        out = self.critic_grads_device(which, real2, x1, z, ep, reduce=not (native and update))
        if update:
            if native:
                self._critic_update_native(D)
            else:
                D.adam_step(self.lrD, self.b1, self.b2)
        vals = out.cpu().numpy().copy()
        self.last_gp = float(vals[2])
        return [np.float32(vals[0]), np.float32(vals[1])]

    def netD_y2_train(self, inputs, update=True):
        return self._critic_train(0, inputs, update)

    def netD_dem_train(self, inputs, update=True):
        return self._critic_train(1, inputs, update)

    # ---- generator (TG:573-598) -----------------------------------------------------------------------
    def gen_device(self, x1, real2, z, grads, update=False):
        torch = self.torch
        n = int(x1.shape[0])
        L = _lib.lib()
        fn = L.depgan_gen_grads if grads else L.depgan_gen_eval
        with torch.cuda.device(self.device):
            _lib.check(fn(self.G.handle, self.Dy2.handle, self.Ddem.handle, x1.data_ptr(), real2.data_ptr(),
                          z.data_ptr(), self.thr, self.out6.data_ptr(), self.sums.data_ptr(), n, n * self.world,
                          _stream(torch)), "gen_grads" if grads else "gen_eval")
            if self.dist is not None:  # batch-global dice / volume / means need the summed partials
                if self.collective == "torch":
                    part = self.sums[:6].clone()
                    self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM)
                    self.sums[:6] = part
                    if grads:
                        self._allreduce_grads(self.G)
                elif grads and update:      # the sums ride along with the bucket of the fused update
                    self.G.dp_update(self.lrG, self.b1, self.b2, extra=self.sums, n_extra=6)
                elif grads:
                    self.G.dp_allreduce_grads(self.sums, 6)
                else:
                    self.G.dp_allreduce_f64(self.sums, 6)
                _lib.check(L.depgan_gen_loss_finalize(self.out6.data_ptr(), self.sums.data_ptr(), _stream(torch)),
                           "gen_loss_finalize")
            elif grads and update:
                self.G.adam_step(self.lrG, self.b1, self.b2)
            if self.collective == "torch" and grads and update:
                self.G.adam_step(self.lrG, self.b1, self.b2)
        return self.out6

    # ---- the ten candidate evaluations of a generator iteration in one pass (TG:868-874) ------------------------
    def enable_batched_eval(self, k_noise=10, max_rows=320):
        """Creates inference handles for up to k_noise * batch rows (capped at max_rows: ~0.14 GB of workspace per row at
        256 x 256) on the training networks' parameter buffers (own workspaces; the training handles' backward buffers
        are not multiplied) and switches gen_iteration_device to batched evaluation passes.  The candidates' losses are
        those of the one-by-one evaluation (slices never mix).  Returns False (and changes nothing) when not even two
        batches fit under max_rows -- large batches fill the GPU on their own."""
        from .api import Dis_C2D_FCN1, Gen_UNet2D
        G = self.G
        per = min(int(k_noise), int(max_rows) // G.cfg.max_batch)
        if per < 2:
            return False
        rows = per * G.cfg.max_batch
        H, W = G.cfg.H, G.cfg.W
        self._Ge = Gen_UNet2D(G.input_shape, G.noiseZ_shape, 32, 1, precision=G.precision, max_batch=rows,
                              device=str(self.device), share_params_with=G)
        self._De = [Dis_C2D_FCN1((H, W, 1), precision=D.precision, max_batch=rows, device=str(self.device),
                                 share_params_with=D) for D in (self.Dy2, self.Ddem)]
        torch = self.torch
        self._ek = per
        self._escratch = torch.empty(rows * H * W * (G.cfg.nicg + 1), dtype=torch.float32, device=self.device)
        self._eout6 = torch.zeros((self._ek, 6), dtype=torch.float32, device=self.device)
        self._esums = torch.zeros((self._ek, 8), dtype=torch.float64, device=self.device)
        self.batched_eval = True
        return True

    def gen_eval_multi_device(self, x1, real2, noises):
        """noises (k, n, L, 1) CUDA tensor -> (k, 6) float32 CUDA tensor of [loss, lf, lfd, M1, M3, M4] per candidate."""
        torch = self.torch
        k, n = int(noises.shape[0]), int(x1.shape[0])
        if n > self.G.cfg.max_batch:
            raise ValueError("batch larger than the networks' max_batch")
        if k > self._ek:  # more candidates than one pass holds: several passes
            return torch.cat([self.gen_eval_multi_device(x1, real2, noises[c:c + self._ek]).clone()
                              for c in range(0, k, self._ek)])
        L = _lib.lib()
        with torch.cuda.device(self.device):
            for net in (self._Ge, *self._De):  # derived tensors of the evaluation handles follow the trained weights
                net.prepare()
            _lib.check(L.depgan_gen_eval_multi(self._Ge.handle, self._De[0].handle, self._De[1].handle, x1.data_ptr(),
                                               real2.data_ptr(), noises.data_ptr(), self.thr, self._eout6.data_ptr(),
                                               self._esums.data_ptr(), self._escratch.data_ptr(), n, k, n * self.world,
                                               _stream(torch)), "gen_eval_multi")
            if self.dist is not None:
                if self.collective == "torch":
                    self.dist.all_reduce(self._esums, op=self.dist.ReduceOp.SUM)
                else:
                    self.G.dp_allreduce_f64(self._esums, 8 * k)
                _lib.check(L.depgan_gen_loss_finalize_multi(self._eout6.data_ptr(), self._esums.data_ptr(), k,
                                                            n * self.world, self.G.cfg.H * self.G.cfg.W, _stream(torch)),
                           "gen_loss_finalize_multi")
        return self._eout6[:k]

    def netG_no_update(self, inputs):
        x1, real2, z = [self._dev(a) for a in inputs]
        return [np.float32(v) for v in self.gen_device(x1, real2, z, False).cpu().numpy()]

    def netG_train(self, inputs, update=True):
        x1, real2, z = [self._dev(a) for a in inputs]
        vals = self.gen_device(x1, real2, z, True, update=update).cpu().numpy().copy()
        return [np.float32(v) for v in vals]

    # ---- fully asynchronous device-side variants (no host round trip inside a generator iteration) -------
    def critic_update_device(self, which, real2, x1, z, ep, dem=None):
        """One critic update (grads + all-reduce + Adam) on CUDA tensors; returns the loss tensor (no sync)."""
        D = self.Dy2 if which == 0 else self.Ddem
        native = self.collective in ("peer", "nccl")
        out = self.critic_grads_device(which, real2, x1, z, ep, reduce=not native, dem=dem)
        if native:
            self._critic_update_native(D)
        else:
            D.adam_step(self.lrD, self.b1, self.b2)
        return out

    def gen_iteration_device(self, crit_y2_batches, crit_dem_batches, x1, real2, noises):
        """TG:796-878 on CUDA tensors.  crit_*_batches: lists of (real2, x1, z, ep); noises: (k,N,L,1) tensor.
        The argmin over the k candidate losses and the gather of the selected noise stay on the device."""
        torch = self.torch
        todo = [(0, b) for b in crit_y2_batches] + [(1, b) for b in crit_dem_batches]
        if self.batched_eval and todo:
            # the generator is frozen until the end of the iteration: its forwards for all critic updates run as batched
            # passes (up to k_noise batches at a time) on the evaluation handle
            self._Ge.prepare()
            n = int(todo[0][1][1].shape[0])
            per = max(1, self._Ge.cfg.max_batch // n)
            for c0 in range(0, len(todo), per):
                chunk = todo[c0:c0 + per]
                xcat = torch.cat([b[1] for _, b in chunk]) if len(chunk) > 1 else chunk[0][1][1]
                zcat = torch.cat([b[2] for _, b in chunk]) if len(chunk) > 1 else chunk[0][1][2]
                dems = self._Ge.forward_device(xcat.contiguous(), zcat.contiguous())
                for j, (which, b) in enumerate(chunk):
                    self.critic_update_device(which, *b, dem=dems[j * n:(j + 1) * n])
        else:
            for which, b in todo:
                self.critic_update_device(which, *b)
        if self.batched_eval:
            losses = self.gen_eval_multi_device(x1, real2, noises)[:, 0].contiguous()
        else:
            losses = torch.empty(noises.shape[0], dtype=torch.float32, device=self.device)
            for k in range(noises.shape[0]):
                losses[k] = self.gen_device(x1, real2, noises[k], False)[0]
        sel = torch.argmin(losses)                        # TG:875-876
        z = noises.index_select(0, sel.reshape(1))[0].contiguous()
        out = self.gen_device(x1, real2, z, True, update=True)
        self.gen_iterations += 1
        return losses, out

    # ---- one generator iteration of the reference schedule (TG:796-878) --------------------------------
    def diters(self, Diters=5):
        """Critic iterations for the current generator iteration (TG:792-795)."""
        return 100 if (self.gen_iterations < 25 or self.gen_iterations % 500 == 0) else Diters

    def gen_iteration(self, crit_y2_batches, crit_dem_batches, x1, real2, noises):
        """Critic updates on the given batches (each [real_2tp, real_1tp, noise, ep]), then k_noise forward-only
        evaluations, argmin over the total loss, one generator update with the selected noise (TG:867-878)."""
        for b in crit_y2_batches:
            self.netD_y2_train(b)
        for b in crit_dem_batches:
            self.netD_dem_train(b)
        x1d, r2d = self._dev(x1), self._dev(real2)
        losses = []
        for nz in noises:
            losses.append(float(self.gen_device(x1d, r2d, self._dev(nz), False)[0].item()))
        k = int(np.array(losses).argmin(0))
        out = self.netG_train([x1d, r2d, self._dev(noises[k])])
        self.gen_iterations += 1
        return k, losses, out


    # ---- the reference's training loop (TG:780-894) ------------------------------------------------------
    def fit(self, x1_train, y2_train, niter=1, batchSize=16, Diters=5, noiseSize=32, k_noise=10, val=None,
            fixed_noise=None, logger=None, log_dir=None, save_path=None, save_every=1, seed=None, shuffle=True, verbose=False):
        """``niter`` epochs of the DEP-GAN loop on host arrays ``x1_train (M,H,W,nicg)`` / ``y2_train (M,H,W,1)``:
        per epoch a shuffle (TG:785-788), then :func:`epoch_schedule`; noise ~ N(0,1), ep ~ U(0,1) per critic update
        (TG:807-808, 822-823), ``k_noise`` candidate noises for the generator and the argmin rule (TG:869-878); the
        scalars of TG:810-840 and 879-885 go to ``logger.log_scalar``; every 10th generator iteration the three
        validation means of TG:845-855 on ``val = (x1_val, y2_val)`` with ``fixed_noise``; ``netG.save(save_path)`` every
        ``save_every`` generator iterations (the reference saves after each, TG:892).  The reference draws from the
        global NumPy RNG; ``seed`` makes the run reproducible.  Returns the logger."""
        rng = np.random.default_rng(seed)
        if logger is None:  # TG:775: Logger('./logs/...') -> TensorBoard event files when a directory is given
            from .tblog import TensorBoardLogger
            logger = TensorBoardLogger(log_dir) if log_dir else ScalarLog()
        x1_train = np.ascontiguousarray(x1_train, np.float32)
        y2_train = np.ascontiguousarray(y2_train, np.float32)
        crit_it = crit_dem_it = 0
        errD = errD_real = errD_fake = errD_dem = errD_real_dem = errD_fake_dem = 0.0
        for epoch in range(niter):
            idx = np.arange(x1_train.shape[0])
            if shuffle:
                rng.shuffle(idx)
            x1_e, y2_e = x1_train[idx], y2_train[idx]
            batches = x1_e.shape[0] // batchSize
            for ev in epoch_schedule(batches, self.gen_iterations, Diters):
                if ev[0] in ("y2", "dem"):
                    b = ev[1]
                    r1, r2 = x1_e[b * batchSize:(b + 1) * batchSize], y2_e[b * batchSize:(b + 1) * batchSize]
                    noise = rng.standard_normal((batchSize, noiseSize, 1))
                    ep = rng.uniform(size=(batchSize, 1, 1, 1))
                    if ev[0] == "y2":
                        errD_real, errD_fake = self.netD_y2_train([r2, r1, noise, ep])
                        errD = errD_real - errD_fake
                        logger.log_scalar("errCrit_aaLosses", errD, crit_it)
                        logger.log_scalar("errCrit_aReal_losses", errD_real, crit_it)
                        logger.log_scalar("errCrit_aFake_losses", errD_fake, crit_it)
                        crit_it += 1
                    else:
                        errD_real_dem, errD_fake_dem = self.netD_dem_train([r2, r1, noise, ep])
                        errD_dem = errD_real_dem - errD_fake_dem
                        logger.log_scalar("errCrit_DEM_aaLosses", errD_dem, crit_dem_it)
                        logger.log_scalar("errCrit_DEM_aReal_losses", errD_real_dem, crit_dem_it)
                        logger.log_scalar("errCrit_DEM_aFake_losses", errD_fake_dem, crit_dem_it)
                        crit_dem_it += 1
                    continue
                b, g_it = ev[1], self.gen_iterations
                for tag, v in (("errDC_aaLosses", errD), ("errDC_aReal_losses", errD_real), ("errDC_aFake_losses", errD_fake),
                               ("errDC_DEM_aaLosses", errD_dem), ("errDC_DEM_aReal_losses", errD_real_dem),
                               ("errDC_DEM_aFake_losses", errD_fake_dem)):
                    logger.log_scalar(tag, v, g_it)
                if val is not None and g_it % 10 == 0:                         # TG:842-855
                    x1_v, y2_v = val
                    fz = fixed_noise if fixed_noise is not None else np.zeros((x1_v.shape[0], noiseSize, 1), np.float32)
                    v_fake = float(np.mean(self.Dy2.predict(np.ascontiguousarray(x1_v[..., :1], np.float32))))
                    v_real = float(np.mean(self.Dy2.predict(np.ascontiguousarray(y2_v, np.float32))))
                    v_gen = float(np.mean(self.Dy2.predict(self.G.predict([x1_v, fz]))))
                    logger.log_scalar("val_D_fake_loss", v_fake, g_it)
                    logger.log_scalar("val_D_real_loss", v_real, g_it)
                    logger.log_scalar("val_D_real_generated_loss", v_gen, g_it)
                    if g_it % 500 == 0:                                        # TG:857-865: image summaries
                        attributed = self.G.predict([x1_v, fz])
                        fake = np.ascontiguousarray(x1_v[..., :1], np.float32) + attributed
                        logger.log_images("attributed_img_step%d" % g_it, attributed[:50], g_it, "")
                        logger.log_images("fake_img_step%d" % g_it, fake[:50], g_it, "")
                r1, r2 = x1_e[b * batchSize:(b + 1) * batchSize], y2_e[b * batchSize:(b + 1) * batchSize]
                noises = rng.standard_normal((k_noise, batchSize, noiseSize, 1)).astype("float32")
                _, _, out = self.gen_iteration([], [], r1, r2, noises)        # TG:867-878
                for tag, v in zip(("errG_losses", "errG_CY2_losses", "errG_DEM_losses", "errG_MSE_losses",
                                   "errG_VOL_losses", "errG_WMH_losses"), out):
                    logger.log_scalar(tag, float(v), g_it)
                if verbose:
                    print("GEN ERR [%d/%d][%d] errG: %f" % (epoch, niter, g_it, float(out[0])))
                if save_path and save_every and (g_it + 1) % save_every == 0:
                    self.G.save(save_path)                                    # TG:892
        return logger
