"""Multi-rank host logic on CPU (gloo, world_size 2): subject sharding for the inference sweep, and the data-parallel
reductions the trainer relies on -- summed loss partials and summed shard gradients reproduce the 1-rank global-batch
numbers.  The per-shard terms come from the oracle; the test pins the *decomposition* the CUDA path uses:
means scaled by 1/global_n, batch-global dice / volume from summed counts, gradient = sum of shard gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depgan_b200 import synth
from depgan_b200.infer import shard_items, shard_range
from oracle import depgan_oracle as O

THR = 0.178


def test_shard_range_partitions():
    for n in (0, 1, 7, 39, 156):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert shard_items(list("abcde"), 1, 2) == ["d", "e"]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(PG, PD1, PD2, x1, y2, z, global_n):
    """Loss partial sums of one shard exactly as depgan_gen_eval leaves them (sums[0..5])."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    with torch.no_grad():
        dem = O.gen_forward(PG, t(x1), t(z))
        base = t(x1)[..., 0:1]
        fake2 = base + dem
        s0 = O.critic_forward(PD1, fake2).sum()
        s1 = O.critic_forward(PD2, dem).sum()
        s2 = (dem - (t(y2) - base)).abs().sum()
        wr = (t(y2).float() >= np.float32(THR)).double()
        wf = (fake2.float() >= np.float32(THR)).double()
    return torch.stack([s0, s1, s2, wr.sum(), wf.sum(), (wr * wf).sum()])


def _finalize(s, n, hw):
    lf, lfd = s[0] / n, s[1] / n
    m1 = 100.0 * s[2] / (n * hw)
    m4 = 1.0 - (2.0 * s[5] + 1e-7) / (s[3] + s[4] + 1e-7)
    m3 = 100.0 * (s[3] / 1000.0 - s[4] / 1000.0) ** 2
    return [float(-lf - lfd + m1 + m3 + m4), float(lf), float(lfd), float(m1), float(m3), float(m4)]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    H, N = 16, 4
    PG = O.to_torch(synth.init_weights(O.gen_manifest(1, 1), seed=1, trained_like=True), requires_grad=True)
    PD1 = O.to_torch(synth.init_weights(O.critic_manifest(H, H), seed=2, trained_like=True))
    PD2 = O.to_torch(synth.init_weights(O.critic_manifest(H, H), seed=3, trained_like=True))
    x1, y2, _ = synth.make_im_pair(N, H, H, thr=THR, seed=4)
    z = synth.make_noise(N, seed=5)
    lo, hi = shard_range(N, rank, world)
    # loss partials: all-reduce SUM of sums[0..5], then the same finalisation on every rank
    part = _partials(PG, PD1, PD2, x1[lo:hi], y2[lo:hi], z[lo:hi], N)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    got = _finalize(part, N, H * H)
    # gradient: each shard differentiates (its sums scaled by 1/global_n); the all-reduced SUM is the global gradient
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    dem = O.gen_forward(PG, t(x1[lo:hi]), t(z[lo:hi]))
    base = t(x1[lo:hi])[..., 0:1]
    shard_loss = (-O.critic_forward(PD1, base + dem).sum() - O.critic_forward(PD2, dem).sum()) / N \
        + 100.0 * (dem - (t(y2[lo:hi]) - base)).abs().sum() / (N * H * H)
    keys = [k for k in PG if PG[k].requires_grad]
    grads = torch.autograd.grad(shard_loss, [PG[k] for k in keys], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(PG[k])).reshape(-1) for k, g in zip(keys, grads)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        PG1 = O.to_torch(synth.init_weights(O.gen_manifest(1, 1), seed=1, trained_like=True), requires_grad=True)
        want = O.gen_loss(PG1, PD1, PD2, t(x1), t(y2), t(z), THR)
        gw = torch.autograd.grad(want[0], [PG1[k] for k in keys], allow_unused=True)
        flat_w = torch.cat([(g if g is not None else torch.zeros_like(PG1[k])).reshape(-1) for k, g in zip(keys, gw)])
        out["loss_err"] = float(max(abs(a - float(b)) for a, b in zip(got, want)))
        out["grad_err"] = float((flat - flat_w).abs().max() / flat_w.abs().max())
    dist.destroy_process_group()


def test_data_parallel_decomposition_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["loss_err"] < 1e-9, out["loss_err"]
    assert out["grad_err"] < 1e-9, out["grad_err"]
