"""Inference driver (EG:616-741 / EU:553-603) and data-parallel training on real GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from depgan_b200 import synth
from oracle import depgan_oracle as O
from tests import util

pytestmark = pytest.mark.gpu
THR = 0.178


def test_subject_driver_is_bit_exact_given_the_predictions():
    from depgan_b200 import Gen_UNet2D
    from depgan_b200.infer import predict_subject_dem
    H, Z, R = 64, 6, 4
    P = util.gen_weights(1, 1, seed=3)
    vol, _, mask = synth.make_im_pair(Z, H, H, thr=THR, seed=1)
    noises = [synth.make_noise(Z, seed=50 + r) for r in range(R)]
    g = Gen_UNet2D((H, H, 1), precision="bf16", max_batch=4)  # Z > max_batch exercises the chunking
    g.set_weights(P)
    res = predict_subject_dem(g, vol, mask, THR, n_repeat=R, noises=noises)
    preds = [g.predict([vol, nz])[..., 0] for nz in noises]          # the same GPU predictions, via predict()
    dem_o = O.inference_mean(preds, mask)
    count_o, labels_o, fake2_o = O.dem_postproc(vol[..., 0], dem_o, mask, THR)
    assert np.array_equal(res["dem"], dem_o) and np.array_equal(res["fake2"], fake2_o)
    assert np.array_equal(res["labels"], labels_o.astype(np.uint8)) and res["wmh_voxels"] == count_o
    # and the DEM itself is within the bf16 tolerance of the oracle network
    want = np.mean([util.oracle_gen(P, vol, nz)[..., 0] * mask for nz in noises], axis=0)
    assert np.abs(res["dem"] - want).max() <= 1e-2


def test_subject_driver_uresnet():
    from depgan_b200 import Gen_UNet2D
    from depgan_b200.infer import predict_subject_uresnet
    H, Z, R = 32, 3, 3
    vol, mask = synth.make_flair(Z, H, H, seed=2)
    noises = [synth.make_noise(Z, seed=60 + r) for r in range(R)]
    g = Gen_UNet2D((H, H, 1), (32, 1), 32, 4, precision="bf16", max_batch=4)
    res = predict_subject_uresnet(g, vol, mask, n_repeat=R, noises=noises)
    preds = [g.predict([vol, nz]) for nz in noises]
    mean_o = O.inference_mean(preds, mask[..., None])
    lab_o, cnt_o = O.uresnet_labels(mean_o)
    assert np.array_equal(res["prob_mean"], mean_o) and np.array_equal(res["labels"], lab_o)
    assert res["wmh_voxels"] == cnt_o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from depgan_b200 import Dis_C2D_FCN1, Gen_UNet2D
    from depgan_b200.infer import shard_range
    from depgan_b200.trainer import DepGanTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    H, N = 32, 4
    dev = "cuda:%d" % rank
    PG, PD1, PD2 = util.gen_weights(1, 1, seed=1), util.critic_weights(H, H, seed=2), util.critic_weights(H, H, seed=3)
    x1, y2, _ = synth.make_im_pair(N, H, H, thr=THR, seed=4)
    z, ep = synth.make_noise(N, seed=5), synth.make_eps(N, seed=6)
    lo, hi = shard_range(N, rank, world)
    n = hi - lo
    G = Gen_UNet2D((H, H, 1), precision="fp32", max_batch=n, device=dev, training=True)
    D1 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * n, device=dev, training=True)
    D2 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * n, device=dev, training=True)
    G.set_weights(PG), D1.set_weights(PD1), D2.set_weights(PD2)
    tr = DepGanTrainer(G, D1, D2, THR)
    assert tr.world == world
    r = {}
    r["d"] = [float(v) for v in tr.netD_y2_train([y2[lo:hi], x1[lo:hi], z[lo:hi], ep[lo:hi]], update=False)]
    r["gp"] = tr.last_gp
    r["d_grads"] = D1.get_grads()
    r["g"] = [float(v) for v in tr.netG_train([x1[lo:hi], y2[lo:hi], z[lo:hi]], update=False)]
    r["g_grads"] = G.get_grads()
    if rank == 0:
        out.update(r)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_two_gpus_reproduce_global_batch_step():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    H, N = 32, 4
    PG, PD1, PD2 = util.gen_weights(1, 1, seed=1), util.critic_weights(H, H, seed=2), util.critic_weights(H, H, seed=3)
    x1, y2, _ = synth.make_im_pair(N, H, H, thr=THR, seed=4)
    z, ep = synth.make_noise(N, seed=5), synth.make_eps(N, seed=6)
    ora = O.OracleTrainer(PG, PD1, PD2, THR)
    want = ora.netD_y2_train([y2, x1, z, ep], update=False)
    assert np.allclose(out["d"], want, rtol=1e-4, atol=1e-5), (out["d"], want)
    assert abs(out["gp"] - ora.last_gp) <= 1e-4 * max(1.0, abs(ora.last_gp))
    for k, w in ora.last_grads.items():
        w = w.numpy()
        if np.linalg.norm(w) > 1e-9:
            assert np.linalg.norm(out["d_grads"][k] - w) / np.linalg.norm(w) < 2e-3, k
    want = ora.netG_train([x1, y2, z], update=False)
    assert np.allclose(out["g"], want, rtol=1e-4, atol=1e-5), (out["g"], want)
    for k, w in ora.last_grads.items():
        w = w.numpy()
        if np.linalg.norm(w) > 1e-9:
            assert np.linalg.norm(out["g_grads"][k] - w) / np.linalg.norm(w) < 2e-3, k


# ---- data-parallel DEP-UResNet fit: synchronised BatchNorm over two GPUs == one GPU on the concatenated batch ----
def _fit_data(H, N, seed=0):
    x, _ = synth.make_flair(N, H, H, seed=seed + 2)
    z = synth.make_noise(N, seed=seed + 3)
    rng = np.random.default_rng(seed)
    onehot = np.eye(4, dtype=np.float32)[rng.integers(0, 4, (N, H, H))]
    keep = (rng.uniform(size=(N, H // 4, H // 4, 96)) >= 0.25).astype(np.uint8)
    return x, z, onehot, keep


def _fit_dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from depgan_b200 import Gen_UNet2D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    H, N = 32, 8
    dev = "cuda:%d" % rank
    per = N // world
    g = Gen_UNet2D((H, H, 1), (32, 1), 32, 4, precision="fp32", max_batch=per, device=dev, training="fit")
    g.set_weights(util.gen_weights(1, 4, seed=1))
    g.enable_data_parallel()
    sl = slice(rank * per, (rank + 1) * per)
    loss = g.train_on_batch_device(*[torch.from_numpy(a[sl]).to(dev) for a in _fit_data(H, N)])
    if rank == 0:
        out["loss"] = float(loss.item())
        out["grads"] = g.get_grads()
        out["weights"] = g.get_weights()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_uresnet_fit_data_parallel_sync_bn_matches_single_gpu():
    import torch.multiprocessing as mp
    from depgan_b200 import Gen_UNet2D
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fit_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    H, N = 32, 8
    g = Gen_UNet2D((H, H, 1), (32, 1), 32, 4, precision="fp32", max_batch=N, training="fit")
    g.set_weights(util.gen_weights(1, 4, seed=1))
    loss = g.train_on_batch_device(*[torch.from_numpy(a).cuda() for a in _fit_data(H, N)])
    assert abs(out["loss"] - float(loss.item())) <= 1e-5 * max(1.0, abs(float(loss.item())))
    want = g.get_grads()
    a = np.concatenate([out["grads"][k].ravel() for k in want])
    b = np.concatenate([want[k].ravel() for k in want])
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4
    for k, w in want.items():
        nrm = np.linalg.norm(w)
        if nrm > 1e-6:
            assert np.linalg.norm(out["grads"][k] - w) / nrm < 5e-3, k
    wts = g.get_weights()  # after the Adam step, moving statistics included
    for k in wts:
        assert np.allclose(out["weights"][k], wts[k], rtol=2e-4, atol=2e-5), k


# ---- the native transports of the gradient sum (dp.cu): fused peer-memory reduce + Adam, NCCL behind the C ABI ----
def _collective_worker(rank, world, port, kind, out):
    import torch.distributed as dist
    from depgan_b200 import Dis_C2D_FCN1, Gen_UNet2D
    from depgan_b200.infer import shard_range
    from depgan_b200.trainer import DepGanTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    H, N = 32, 4
    dev = "cuda:%d" % rank
    PG, PD1, PD2 = util.gen_weights(1, 1, seed=1), util.critic_weights(H, H, seed=2), util.critic_weights(H, H, seed=3)
    x1, y2, _ = synth.make_im_pair(N, H, H, thr=THR, seed=4)
    z, ep = synth.make_noise(N, seed=5), synth.make_eps(N, seed=6)
    lo, hi = shard_range(N, rank, world)
    n = hi - lo
    G = Gen_UNet2D((H, H, 1), precision="fp32", max_batch=n, device=dev, training=True)
    D1 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * n, device=dev, training=True)
    D2 = Dis_C2D_FCN1((H, H, 1), precision="fp32", max_batch=3 * n, device=dev, training=True)
    G.set_weights(PG), D1.set_weights(PD1), D2.set_weights(PD2)
    tr = DepGanTrainer(G, D1, D2, THR, collective=kind)
    r = {"d": [], "g": []}
    for it in range(3):  # three updates of each network: the mailboxes' double buffering wraps
        r["d"].append([float(v) for v in tr.netD_y2_train([y2[lo:hi], x1[lo:hi], z[lo:hi], ep[lo:hi]])])
        r["d"].append([float(v) for v in tr.netD_dem_train([y2[lo:hi], x1[lo:hi], z[lo:hi], ep[lo:hi]])])
        r["g"].append([float(v) for v in tr.netG_no_update([x1[lo:hi], y2[lo:hi], z[lo:hi]])])
        r["g"].append([float(v) for v in tr.netG_train([x1[lo:hi], y2[lo:hi], z[lo:hi]])])
    # gradient inspection without an update (the all-reduce that leaves the sum in the bucket)
    r["d_noupd"] = [float(v) for v in tr.netD_y2_train([y2[lo:hi], x1[lo:hi], z[lo:hi], ep[lo:hi]], update=False)]
    r["d_grads"] = D1.get_grads()
    r["w"] = {"G": G.get_weights(), "D1": D1.get_weights(), "D2": D2.get_weights()}
    out[(kind, rank)] = r
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_native_collectives_reproduce_the_torch_distributed_path():
    """Three full update rounds (both critics, ten... one evaluation, the generator) on two GPUs with each transport:
    'peer' (fused reduce + Adam over CUDA-IPC mailboxes) and 'nccl' (ncclAllReduce issued by the library) give the losses
    and the weights of the torch.distributed path to float rounding; with 'peer' the two replicas are bit-identical
    (same summation order on every rank)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    for kind in ("torch", "peer", "nccl"):
        mp.spawn(_collective_worker, args=(2, _free_port(), kind, out), nprocs=2, join=True)
    ref = out[("torch", 0)]
    for kind in ("peer", "nccl"):
        got = out[(kind, 0)]
        assert np.allclose(got["d"], ref["d"], rtol=2e-4, atol=1e-5), (kind, got["d"], ref["d"])
        assert np.allclose(got["g"], ref["g"], rtol=2e-4, atol=1e-5), (kind, got["g"], ref["g"])
        assert np.allclose(got["d_noupd"], ref["d_noupd"], rtol=2e-4, atol=1e-5)
        for k, w in ref["d_grads"].items():
            if np.linalg.norm(w) > 1e-9:
                assert np.linalg.norm(got["d_grads"][k] - w) / np.linalg.norm(w) < 1e-4, (kind, k)
        for net in ("G", "D1", "D2"):
            for k, w in ref["w"][net].items():
                # Adam with beta_1 = 0 moves every weight by ~lr per step whatever the gradient's size, so weights whose
                # gradient is rounding noise may differ by 2 * lr * steps; everything else agrees far tighter
                assert np.abs(got["w"][net][k] - w).max() <= 6.5e-4, (kind, net, k)
    for net in ("G", "D1", "D2"):
        for k, w in out[("peer", 0)]["w"][net].items():
            assert np.array_equal(out[("peer", 1)]["w"][net][k], w), (net, k)
