#!/usr/bin/env python
"""bench.py -- headline benchmark of the DEP-GAN hot path on B200 (contract: see the task statement / DESIGN.md).

Workloads (BASELINE.json configs):
  uresnet_infer   (default, configs[1]) DEP-UResNet inference, 256x256, batch 64 per GPU, bf16 tcgen05 path
  depgan_infer    (configs[0])          DEP-GAN generator inference, 256x256 IM slices, batch 16
  depgan_train    (configs[2])          DEP-GAN two-critic generator iteration (IM), batch 32 per GPU
  depgan_train_pf (configs[3])          PROB+FLAIR two-critic training, GLOBAL batch 256 split over the GPUs (strong scaling)
  cohort          (configs[4])          whole-cohort 4-fold test sweep, 156 synthetic 48-slice subjects x 10 repeats

The default run prints ONE JSON line for configs[1] and carries the other four configs as sub-records under
"configs" (each with its own value / unit / ms, measured in the same process right after the headline leg).

One "step" = one pass of the hot path over one batch of synthetic slices.  `value` = slices/s with inputs
resident in HBM; `e2e` = the same through the Keras-like public call with pinned host buffers (H2D + D2H inside
the timed region).  N > 1 (torchrun): every rank runs its own batch, no data-path collective (weak scaling).
`--impl reference` times the CPU oracle (torch fp32 restatement of the Keras graph; the reference itself cannot
run here) on the host cores with the same metric/config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

G_FLOP = {(1, 1): 23.513e9, (2, 1): 23.551e9, (1, 4): 23.526e9}  # per slice

WORKLOADS = {
    "uresnet_infer": dict(nicg=1, nc_out=4, batch=64, head="softmax",
                          desc="DEP-UResNet (wNoises) inference, 256x256, batch 64 per GPU (BASELINE configs[1])"),
    "depgan_infer": dict(nicg=1, nc_out=1, batch=16, head="tanh",
                         desc="DEP-GAN generator inference, 256x256 IM slices, batch 16 (BASELINE configs[0])"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(w, n, seed):
    from depgan_b200 import synth
    if w["nc_out"] == 4:
        x, _ = synth.make_flair(n, 256, 256, seed=seed)
    else:
        x, _, _ = synth.make_im_pair(n, 256, 256, nicg=w["nicg"], seed=seed)
    return x, synth.make_noise(n, seed=seed + 1)


def make_weights(w, net=None):
    """Seeded synthetic weights in the Keras layout.  The GPU arm takes names / shapes from the library's own manifest
    (C ABI); only the CPU arms (cpu_baseline leg, --impl reference) read the oracle's."""
    from depgan_b200 import synth
    if net is not None:
        man = [(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in net.manifest]
    else:
        from oracle import depgan_oracle as O
        man = O.gen_manifest(w["nicg"], w["nc_out"])
    return synth.init_weights(man, seed=0, trained_like=True)


# ---------------------------------------------------------------------------------------------------------
# CPU arms (oracle): cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------
def cpu_forward_rate(w, P, sample, reps):
    import torch
    from oracle import depgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    Pt = O.to_torch(P, torch.float32)
    x, z = make_inputs(w, sample, seed=5)
    xt, zt = torch.from_numpy(x), torch.from_numpy(z)
    times = []
    with torch.no_grad():
        O.gen_forward(Pt, xt[:2], zt[:2], w["head"])  # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            O.gen_forward(Pt, xt, zt, w["head"])
            times.append(time.perf_counter() - t0)
    return sample / statistics.median(times), times


def run_reference(args, w, rank):
    if rank != 0:
        return
    P = make_weights(w)
    sample = 8
    import torch
    from oracle import depgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    Pt = O.to_torch(P, torch.float32)
    x, z = make_inputs(w, sample, seed=5)
    xt, zt = torch.from_numpy(x), torch.from_numpy(z)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            O.gen_forward(Pt, xt[:2], zt[:2], w["head"])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.gen_forward(Pt, xt, zt, w["head"])
        dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {"impl": "reference", "metric": "256x256 slices/sec (gen inference)", "value": val, "unit": "slices/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "batch_per_step_sample": sample,
                       "note": "per-slice rate of the same graph on %d-slice batches (a bounded sample of the batch-64 "
                               "workload; the CPU rate does not depend on the batch size beyond 8)" % sample},
            "cpu_baseline": {"value": val, "unit": "slices/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d-slice batches x %d steps of the torch-CPU fp32 oracle "
                                       "(Keras reference not runnable here)" % (sample, args.steps)},
            "e2e": {"value": val, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from depgan_b200 import Gen_UNet2D, _lib, launch_count

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # every rank's pinned staging buffers on the memory of its own GPU's NUMA node (DEPGAN_NO_NUMA=1: off, for A/B)
    from depgan_b200.infer import bind_to_gpu_numa_node
    numa = None if os.environ.get("DEPGAN_NO_NUMA") else bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or w["batch"]
    g = Gen_UNet2D((256, 256, w["nicg"]), (32, 1), 32, w["nc_out"], precision=args.precision, max_batch=B,
                   device=str(dev))
    P = make_weights(w, g)
    # weights travel through the Keras-h5 layout, as load_weights would read the shipped files
    h5 = Path("/tmp/depgan_bench_rank%d.h5" % rank)
    from depgan_b200 import h5lite
    h5lite.save_keras_weights(str(h5), P, [n for n, _, _, _ in g.manifest])
    g.load_weights(str(h5))

    x, z = make_inputs(w, B, seed=100 + rank)
    xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
    out = torch.empty((B, 256, 256, w["nc_out"]), dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        g.forward_device(xd, zd, out)
    barrier()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g.forward_device(xd, zd, out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launch_count() - l0
    # clocks / throttle reasons under this exact load: the timed region can be shorter than nvidia-smi's start-up, so
    # the same step loop is continued for >= 0.8 s while nvidia-smi samples every 50 ms
    clocks = None
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
        t_end = time.perf_counter() + 0.8
        while time.perf_counter() < t_end:
            for _ in range(8):
                g.forward_device(xd, zd, out)
            torch.cuda.synchronize(dev)
        clocks = sampler.stop()
        clocks["window"] = "0.8 s continuation of the timed step loop"
    barrier()

    # ---- end to end through the public pipelined path: pinned host -> device -> kernels -> pinned host, every
    # step copies its own inputs in and its result out (depgan_b200.InferencePipeline, what predict() uses) ----
    from depgan_b200 import InferencePipeline
    xh, zh = torch.from_numpy(x).pin_memory(), torch.from_numpy(z).pin_memory()
    ohs = [torch.empty((B, 256, 256, w["nc_out"]), dtype=torch.float32).pin_memory() for _ in range(2)]
    oh = ohs[0]
    pipe = InferencePipeline(g, depth=int(os.environ.get("DEPGAN_PIPE_DEPTH", "2")))
    for i in range(max(3, args.warmup)):
        pipe.submit(xh, zh, ohs[i % 2])
    pipe.flush()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        pipe.submit(xh, zh, ohs[i % 2])
    pipe.flush()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)  # f1 is recorded after all three pipeline streams were flushed
    # the same pipeline with the opt-in float16 device->host output (half the PCIe bytes per slice; predict(out_dtype=))
    ohs16 = [torch.empty((B, 256, 256, w["nc_out"]), dtype=torch.float16).pin_memory() for _ in range(2)]
    pipe16 = InferencePipeline(g, depth=int(os.environ.get("DEPGAN_PIPE_DEPTH", "2")), out_dtype=torch.float16)
    for i in range(max(3, args.warmup)):
        pipe16.submit(xh, zh, ohs16[i % 2])
    pipe16.flush()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(args.steps):
        pipe16.submit(xh, zh, ohs16[i % 2])
    pipe16.flush()
    h1.record()
    barrier()
    ms_e2e16 = h0.elapsed_time(h1)
    d2h16 = int(ohs16[0].numel() * ohs16[0].element_size())
    del pipe16, ohs16

    t = torch.tensor([ms, ms_e2e, ms_e2e16], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e16 = float(t[0]), float(t[1]), float(t[2])

    # ---- roofline leg: per-launch CUDA events around every convolution of the same step ----
    roof = None
    if rank == 0:
        L = _lib.lib()
        ncls = 7  # 0 tc3x3, 1 tc5x5, 2 tc 1x1 / transposed, 3 simt, 4-5 weight gradients, 6 tc conv + fused max-pool
        arr_ms, arr_fl, arr_by = (C.c_double * ncls)(), (C.c_double * ncls)(), (C.c_double * ncls)()
        arr_n = (C.c_longlong * ncls)()
        L.depgan_profile_begin()
        psteps = 3
        for _ in range(psteps):
            g.forward_device(xd, zd, out)
        _lib.check(L.depgan_profile_end(arr_ms, arr_fl, arr_by, arr_n, ncls), "profile_end")
        pk = peaks()
        k = 0  # class 0 = tcgen05 3x3 convolutions: the dominant kernel (conv_tc_kernel<3>)
        if arr_n[k] > 0 and arr_ms[k] > 0:
            achieved = arr_fl[k] / (arr_ms[k] * 1e-3) / 1e12
            conv_share = arr_ms[k] / max(1e-9, sum(arr_ms))
            roof = {"kernel": "tcgen05 implicit-GEMM 3x3 convolutions: conv_tc_kernel<3> (tile) / conv_row_kernel / "
                              "conv_rowg_kernel (row-streaming), %d launches/step" % (arr_n[k] // psteps),
                    "bound": "tensor", "achieved": achieved, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tf_sust"], "peak_source": pk["src"] + " (sustained bf16)",
                    "traffic": None, "avg_launch_ms": arr_ms[k] / arr_n[k],
                    "flops_per_launch": arr_fl[k] / arr_n[k],
                    "hbm_gbs_same_kernel": arr_by[k] / (arr_ms[k] * 1e-3) / 1e9,
                    # the second ceiling of the same launches: the 32-channel 256x256 layers sit below the ridge
                    # (134 FLOP/B vs 211), so their ceiling is the HBM one -- see DESIGN.md section 6
                    "hbm_frac_same_kernel": arr_by[k] / (arr_ms[k] * 1e-3) / 1e9 / pk["hbm"], "hbm_peak": pk["hbm"],
                    "share_of_conv_time": conv_share,
                    "other_classes_ms_per_step": {"tc5x5": arr_ms[1] / psteps, "tc_deconv": arr_ms[2] / psteps,
                                                  "simt": arr_ms[3] / psteps, "tc3x3": arr_ms[0] / psteps,
                                                  "tc3x3_fused_maxpool": arr_ms[6] / psteps},
                    # every 3x3 launch, the ones with the fused 2x2 max-pool epilogue (conv_tc_kernel<3,*,*,4>) included
                    "all_3x3_launches": {"launches_per_step": (arr_n[0] + arr_n[6]) // psteps,
                                         "achieved": (arr_fl[0] + arr_fl[6]) / ((arr_ms[0] + arr_ms[6]) * 1e-3) / 1e12,
                                         "frac": (arr_fl[0] + arr_fl[6]) / ((arr_ms[0] + arr_ms[6]) * 1e-3) / 1e12
                                                 / pk["tf_sust"]}}
        # DRAM traffic per launch of the same kernel class: from the committed ncu --set full capture of this step
        # (scripts/r2_run4.sh -> scripts/summarize_profiles.py r02); the record names the commit it was taken at
        prof_path = ROOT / "profiles" / "ncu_traffic_r02.json"
        if roof and prof_path.exists():
            try:
                pj = json.loads(prof_path.read_text())
                roof["traffic"] = pj.get("dram_bytes_per_launch")
                roof["traffic_source"] = "profiles/ncu_traffic_r02.json (ncu --set full at commit %s)" % pj.get("commit")
                roof["algorithmic_bytes_per_launch"] = arr_by[k] / arr_n[k]
                roof["ncu_tensor_pipe_pct_time_weighted"] = pj.get("time_weighted_tensor_pipe_pct_batch64_cold")
            except Exception:
                pass

    # ---- the drop-in call itself on NumPy arrays, then the other BASELINE configs as sub-records ----
    extra = {}
    if not args.no_extra:
        extra["predict_numpy"] = measure_predict_numpy(args, w, g, rank, world, dev, B)
    del g, out, xd, zd, pipe
    torch.cuda.empty_cache()
    if not args.no_extra:
        extra["configs[0]"] = measure_config0(args, rank, world, dev)
        torch.cuda.empty_cache()
        extra["precision_variants"] = measure_precision_variants(args, rank, dev)
        torch.cuda.empty_cache()
        extra["configs[4]"] = measure_cohort(args, rank, world, dev)
        torch.cuda.empty_cache()

    # ---- second headline metric: DEP-GAN train steps/s (configs[2] weak scaling at 32 slices per GPU; configs[3]
    # PROB+FLAIR at a fixed global batch of 256; data parallel when world > 1) ----
    train = train_pf = None
    if not args.no_train:
        train = measure_train(args, rank, world, dev, steps=10, warmup=2)
        if not args.no_extra:
            train_pf = measure_train(args, rank, world, dev, steps=max(2, min(8, 2 * world)), warmup=1,
                                     workload="depgan_train_pf", global_batch=args.global_batch)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on a bounded sample ----
    cpu = None
    if not args.no_cpu:
        sample = 8
        rate, times = cpu_forward_rate(w, P, sample, reps=3)
        cpu = {"value": rate, "unit": "slices/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d slices x %d reps (median) of the torch-CPU fp32 oracle of the same graph" % (sample, len(times))}

    slices = B * args.steps * world
    value = slices / (ms * 1e-3)
    flop = G_FLOP[(w["nicg"], w["nc_out"])]
    line = {
        "metric": "256x256 slices/sec (gen inference)", "value": value, "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "f16": "f16", "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_gpu": B, "precision": args.precision,
                   "weights": "synthetic (seeded), round-tripped through the Keras-h5 layout",
                   "l2": "per-step activation working set (~%.1f GB) >> 126 MB L2; no explicit flush" % (B * 0.1137),
                   "parallelism": "slices sharded across %d GPU(s), no collective" % world,
                   "numa_binding_rank0": numa},
        "tflops_effective": value * flop / 1e12,
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": slices / (ms_e2e * 1e-3), "unit": "slices/s",
                "h2d_bytes_per_step": int(xh.numel() * 4 + zh.numel() * 4), "d2h_bytes_per_step": int(oh.numel() * 4)},
        "e2e_float16_output": {"value": slices / (ms_e2e16 * 1e-3), "unit": "slices/s",
                               "h2d_bytes_per_step": int(xh.numel() * 4 + zh.numel() * 4), "d2h_bytes_per_step": d2h16,
                               "what": "the same pipeline with the opt-in float16 device->host output "
                                       "(InferencePipeline(out_dtype=float16) / predict(out_dtype=np.float16))"},
        "roofline": roof, "cpu_baseline": cpu, "train": train,
        "configs": {"configs[0]": extra.get("configs[0]"),
                    "configs[1]": "this line (value / e2e / roofline); predict(numpy) below",
                    "configs[2]": "the 'train' record of this line (IM, 32 slices per GPU, weak scaling)",
                    "configs[3]": train_pf, "configs[4]": extra.get("configs[4]")},
        "predict_numpy": extra.get("predict_numpy"),
        "precision_variants": extra.get("precision_variants"),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _barrier(torch, dist, dev, world):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def measure_config0(args, rank, world, dev):
    """BASELINE configs[0]: DEP-GAN generator (tanh head, IM input) inference at batch 16: device-resident rate and the
    rate of the public ``Gen_UNet2D.predict([x, z])`` call on pageable NumPy arrays (host copies inside the timing)."""
    import torch
    import torch.distributed as dist
    from depgan_b200 import Gen_UNet2D, launch_count
    w = WORKLOADS["depgan_infer"]
    B = w["batch"]
    g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 1, precision=args.precision, max_batch=B, device=str(dev))
    g.set_weights(make_weights(w, g))
    x, z = make_inputs(w, B, seed=300 + rank)
    xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
    out = torch.empty((B, 256, 256, 1), dtype=torch.float32, device=dev)
    steps = max(args.steps, 20)
    for _ in range(5):
        g.forward_device(xd, zd, out)
    _barrier(torch, dist, dev, world)
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.forward_device(xd, zd, out)
    e1.record()
    _barrier(torch, dist, dev, world)
    ms = e0.elapsed_time(e1)
    launches = launch_count() - l0
    # the same forward replayed from a CUDA graph (whole-step capture: one graph launch instead of 28 kernel launches)
    out_g = torch.empty_like(out)
    for _ in range(3):
        g.forward_graph(xd, zd, out_g)
    _barrier(torch, dist, dev, world)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        g.forward_graph(xd, zd, out_g)
    g1.record()
    _barrier(torch, dist, dev, world)
    ms_graph = g0.elapsed_time(g1)
    graph_same = bool(torch.equal(out, out_g))
    # the drop-in call itself: predict() on NumPy arrays, 8 batches of 16 per call
    nb = 8
    xs, zs = np.concatenate([x] * nb), np.concatenate([z] * nb)
    g.predict([xs, zs], batch_size=B)
    g.predict([xs, zs], batch_size=B)
    _barrier(torch, dist, dev, world)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        y = g.predict([xs, zs], batch_size=B)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t = torch.tensor([ms, dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, dt = float(t[0]), float(t[1])
    del g
    if rank != 0:
        return None
    return {"workload": w["desc"], "metric": "256x256 slices/sec (gen inference)", "unit": "slices/s",
            "value": B * steps * world / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "batch_per_gpu": B,
            "n_gpus": world, "gpu_launches": int(launches), "tflops_effective": B * steps * world / (ms * 1e-3) * G_FLOP[(1, 1)] / 1e12,
            "cuda_graph_replay": {"value": B * steps * world / (ms_graph * 1e-3), "ms_per_step": ms_graph / steps,
                                  "bit_identical": graph_same,
                                  "what": "Gen_UNet2D.forward_graph: the same forward captured once and replayed"},
            "e2e": {"value": reps * xs.shape[0] * world / dt, "unit": "slices/s",
                    "what": "Gen_UNet2D.predict([x, z], batch_size=16) on pageable NumPy arrays, %d slices per call, "
                            "host wall clock (max over ranks)" % xs.shape[0],
                    "h2d_bytes_per_step": int(x.nbytes + z.nbytes), "d2h_bytes_per_step": int(y.nbytes // nb)}}


def measure_precision_variants(args, rank, dev):
    """The other arithmetic variants of the same forward (configs[0] shape, batch 16, device-resident, rank 0 only):
    'fp32' = the FP32-storage / FP32-accumulate CUDA-core path (the <= 1e-4 variant of BASELINE.json's north_star) and
    'f16x3' = the tensor-core <= 1e-4 variant (split-half storage, three products per convolution), 'bf16' = the tcgen05
    path with bfloat16 storage (what training runs), next to the IEEE-half default."""
    import torch
    from depgan_b200 import Gen_UNet2D
    if rank != 0:
        return None
    w = WORKLOADS["depgan_infer"]
    B = w["batch"]
    x, z = make_inputs(w, B, seed=300)
    xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
    out = torch.empty((B, 256, 256, 1), dtype=torch.float32, device=dev)
    res, ref = {}, None
    for prec, steps in (("fp32", 3), ("f16x3", 10), ("bf16", 20), ("f16", 20)):
        g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 1, precision=prec, max_batch=B, device=str(dev))
        g.set_weights(make_weights(w, g))
        for _ in range(2):
            g.forward_device(xd, zd, out)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.forward_device(xd, zd, out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        y = out.clone()
        if ref is None:
            ref = y  # the fp32 path (<= 1e-4 of the fp64 oracle, tests/test_gpu_nets.py) is the yardstick here
        res[prec] = {"slices_per_s": B / (ms * 1e-3), "ms_per_step": ms,
                     "tflops_effective": B / (ms * 1e-3) * G_FLOP[(1, 1)] / 1e12,
                     "dem_max_abs_vs_fp32_path": float((y - ref).abs().max())}
        if prec == "f16x3":  # three tensor-core products per convolution: the tensor pipe does 3x the algorithmic work
            tw = 3.0 * res[prec]["tflops_effective"]
            res[prec].update(tensor_work_tflops=tw, frac_of_sustained_bf16_peak=tw / peaks()["tf_sust"])
        del g
        torch.cuda.empty_cache()
    # the <= 1e-4 tensor-core variant at the headline shape (configs[1]: softmax head, batch 64)
    w1 = WORKLOADS["uresnet_infer"]
    B1 = w1["batch"]
    x1, z1 = make_inputs(w1, B1, seed=301)
    x1d, z1d = torch.from_numpy(x1).to(dev), torch.from_numpy(z1).to(dev)
    out1 = torch.empty((B1, 256, 256, 4), dtype=torch.float32, device=dev)
    g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 4, precision="f16x3", max_batch=B1, device=str(dev))
    g.set_weights(make_weights(w1, g))
    for _ in range(3):
        g.forward_device(x1d, z1d, out1)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.forward_device(x1d, z1d, out1)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 10
    tw = 3.0 * B1 / (ms * 1e-3) * G_FLOP[(1, 4)] / 1e12
    res["f16x3_configs1_batch64"] = {"slices_per_s": B1 / (ms * 1e-3), "ms_per_step": ms, "tensor_work_tflops": tw,
                                     "frac_of_sustained_bf16_peak": tw / peaks()["tf_sust"]}
    del g, out1, x1d, z1d
    torch.cuda.empty_cache()
    res["what"] = ("DEP-GAN generator forward, batch 16, device-resident; fp32 = CUDA-core FP32 storage + accumulate, "
                   "f16x3 = the tensor-core <= 1e-4 variant (values as IEEE-half (hi, lo) pairs, three tcgen05 products "
                   "per convolution, fp32 accumulate in TMEM), bf16 / f16 = tcgen05 with bfloat16 / IEEE-half "
                   "activations and weights")
    return res


def measure_predict_numpy(args, w, g, rank, world, dev, B):
    """The drop-in call of configs[1] itself: ``predict([x, z], batch_size=B)`` with NumPy in / NumPy out."""
    import torch
    import torch.distributed as dist
    nb = 4
    x, z = make_inputs(w, B, seed=400 + rank)
    xs, zs = np.concatenate([x] * nb), np.concatenate([z] * nb)
    res = {}
    for name, dt_ in (("float32", None), ("float16_opt_in", np.float16)):
        g.predict([xs, zs], batch_size=B, out_dtype=dt_)
        _barrier(torch, dist, dev, world)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            y = g.predict([xs, zs], batch_size=B, out_dtype=dt_)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"value": reps * xs.shape[0] * world / float(t[0]), "unit": "slices/s",
                     "d2h_bytes_per_batch": int(y.nbytes // nb)}
    res["what"] = "Gen_UNet2D.predict on pageable NumPy arrays (%d slices per call, batch_size %d), persistent pinned " \
                  "staging inside the model, host wall clock, max over ranks" % (xs.shape[0], B)
    return res if rank == 0 else None


def measure_cohort(args, rank, world, dev):
    """BASELINE configs[4]: the whole-cohort 4-fold test sweep (EG:378, 484, 616-741): 4 folds x 39 subjects of 48
    synthetic 256x256 slices, 10 noise repeats per subject, float64 mean, DEM post-processing (labels + WMH voxel count),
    subjects sharded over the ranks with no collective, one weight set per fold loaded through the Keras-h5 layout.
    Wall clock from the first subject to the last result on the host, max over ranks."""
    import torch
    import torch.distributed as dist
    from depgan_b200 import Gen_UNet2D, h5lite, launch_count, synth
    from depgan_b200.infer import SubjectEngine, cohort_sweep
    Z, folds, per_fold, R = 48, 4, 39, 10
    thr = 0.178
    w = WORKLOADS["depgan_infer"]
    g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 1, precision=args.precision, max_batch=2 * Z, device=str(dev))
    man3 = [(n.split("/")[0], n.split("/")[1], s_) for n, s_, _, _ in g.manifest]
    paths = []
    for f in range(folds):  # four seeded weight sets, stored as the reference's per-fold .h5 files would be
        pth = "/tmp/depgan_cohort_fold%d_rank%d.h5" % (f, rank)
        h5lite.save_keras_weights(pth, synth.init_weights(man3, seed=40 + f, trained_like=True),
                                  [n for n, _, _, _ in g.manifest])
        paths.append(pth)
    # a small pool of distinct synthetic volumes stands in for the 39 subjects of a fold (the arithmetic does not depend
    # on the content; every subject still gets its own noise stream, upload, kernels, post-processing and download)
    pool = [synth.make_im_pair(Z, 256, 256, thr=thr, seed=900 + i) for i in range(3)]
    out = {}
    for mode, outputs in (("full_outputs", ("dem", "fake2", "labels")), ("labels_only", ("labels",))):
        eng = SubjectEngine(g, "dem", z_max=Z, n_repeat=R, outputs=outputs)
        voxels = [0]

        def sink(sid, r):
            voxels[0] += r["wmh_voxels"]

        def fold(f, n_subjects):
            g.load_weights(paths[f])
            subs = [("f%d_s%02d" % (f, i), pool[i % len(pool)][0], pool[i % len(pool)][2]) for i in range(n_subjects)]
            return cohort_sweep(g, subs, thr, rank=rank, world=world, n_repeat=R, sink=sink, engine=eng)

        fold(0, 2 * world)  # warm-up: two subjects per rank
        _barrier(torch, dist, dev, world)
        eng.h2d_bytes = eng.d2h_bytes = 0
        l0 = launch_count()
        t0 = time.perf_counter()
        mine = 0
        for f in range(folds):
            mine += fold(f, per_fold)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        launches = launch_count() - l0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        n_sub = folds * per_fold
        out[mode] = {"wall_s": dt, "subjects_per_s": n_sub / dt, "slice_forwards_per_s": n_sub * Z * R / dt,
                     "outputs_to_host": list(outputs) + ["wmh_voxels"], "subjects_this_rank": mine,
                     "h2d_bytes_per_subject": eng.h2d_bytes // max(mine, 1),
                     "d2h_bytes_per_subject": eng.d2h_bytes // max(mine, 1), "gpu_launches_this_rank": int(launches)}
        del eng
    del g
    if rank != 0:
        return None
    return {"workload": "whole-cohort 4-fold test sweep: 4 x 39 synthetic subjects x 48 slices x 10 noise repeats "
                        "(74 880 slice-forwards) + DEM post-processing, subjects sharded over %d GPU(s), no collective "
                        "(BASELINE configs[4])" % world,
            "metric": "256x256 slices/sec (gen inference)", "unit": "slices/s", "n_gpus": world,
            "value": out["full_outputs"]["slice_forwards_per_s"], "scaling": "strong",
            "what": "host wall clock incl. weight loading per fold, uploads, kernels, float64 mean, post-processing and "
                    "the download of every subject's maps (max over ranks); 'full_outputs' returns what the testing "
                    "script saves per subject (DEM, predicted follow-up map, label map), 'labels_only' the label map "
                    "and the voxel count", **out}


def dp_self_check(tr, G, D1, D2, rank, world, dev, nicg, thr, cap):
    """Data-parallel correctness inside the bench run (the 2-GPU pytest cases are skipped on one-GPU boxes): every rank
    feeds its shard of ONE fixed global batch through the data-parallel critic / generator gradient calls (NCCL
    all-reduce of the flat bucket and of the loss partial sums); rank 0 then runs the same batch unsharded through a
    non-distributed trainer on the same weights and reports the relative differences."""
    import torch
    from depgan_b200 import synth
    from depgan_b200.trainer import DepGanTrainer
    Bg = max(world, (min(cap, 32) // world) * world)  # global check batch, divisible by world, fits rank 0 unsharded
    per = Bg // world
    x1, y2, _ = synth.make_im_pair(Bg, 256, 256, nicg=nicg, thr=thr, seed=4242)
    z, ep = synth.make_noise(Bg, seed=4243), synth.make_eps(Bg, seed=4244).reshape(-1)
    full = [torch.from_numpy(a).to(dev) for a in (y2, x1, z, ep)]
    sl = slice(rank * per, (rank + 1) * per)
    mine = [t[sl].contiguous() for t in full]
    res = {}
    # data-parallel pass (all ranks)
    c_dp = tr.critic_grads_device(0, *mine).clone()
    gD_dp = D1.grads.clone()
    g_dp = tr.gen_device(mine[1], mine[0], mine[2], True).clone()
    gG_dp = G.grads.clone()
    torch.cuda.synchronize(dev)
    if rank == 0:
        solo = DepGanTrainer(G, D1, D2, thr, distributed=False)
        c_1 = solo.critic_grads_device(0, *full).clone()
        gD_1 = D1.grads.clone()
        g_1 = solo.gen_device(full[1], full[0], full[2], True).clone()
        gG_1 = G.grads.clone()
        torch.cuda.synchronize(dev)

        def rel(a, b):
            return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
        res = {"global_batch": Bg, "per_rank": per,
               "critic_loss_rel_err": rel(c_dp, c_1), "critic_grad_rel_err": rel(gD_dp, gD_1),
               "gen_loss_rel_err": rel(g_dp, g_1), "gen_grad_rel_err": rel(gG_dp, gG_1),
               "what": "||dp - single|| / ||single|| over the 4 critic loss terms / the flat critic gradient / the 6 "
                       "generator loss terms / the flat generator gradient (bf16 activations: the two runs differ only "
                       "in summation order)"}
    if world > 1:
        torch.distributed.barrier()
    return res


def measure_train(args, rank, world, dev, steps, warmup, workload="depgan_train", global_batch=None):
    """BASELINE configs[2]/[3]: DEP-GAN two-critic training, one step = one generator iteration of the reference
    schedule (5 Y2-critic + 5 DEM-critic updates, 10 noise evaluations, 1 generator update; TG:796-878).
    ``global_batch``: strong scaling -- the global batch is fixed and split evenly over the ranks (configs[3]);
    otherwise every rank runs ``--train-batch`` slices (weak scaling).
    Returns the result dict on rank 0 (None elsewhere).  The process group must already exist for world > 1."""
    import torch
    import torch.distributed as dist
    from depgan_b200 import Dis_C2D_FCN1, Gen_UNet2D, launch_count, synth
    from depgan_b200.trainer import DepGanTrainer

    if global_batch:
        if global_batch % world:
            raise ValueError("global batch %d is not divisible by %d ranks" % (global_batch, world))
        B = global_batch // world
    else:
        B = args.train_batch
    nicg = 2 if workload == "depgan_train_pf" else 1
    thr = 0.5 if nicg == 2 else 0.178
    G = Gen_UNet2D((256, 256, nicg), (32, 1), 32, 1, precision=args.train_precision, max_batch=B, device=str(dev),
                   training=True, seed=0)
    D1 = Dis_C2D_FCN1((256, 256, 1), precision=args.train_precision, max_batch=3 * B, device=str(dev), training=True, seed=1)
    D2 = Dis_C2D_FCN1((256, 256, 1), precision=args.train_precision, max_batch=3 * B, device=str(dev), training=True, seed=2)
    tr = DepGanTrainer(G, D1, D2, thr)
    if not os.environ.get("DEPGAN_NO_BATCHED_EVAL"):  # the ten noise evaluations of an iteration as one 10 x B pass
        tr.enable_batched_eval(10)

    def batch(seed):
        x1, y2, _ = synth.make_im_pair(B, 256, 256, nicg=nicg, thr=thr, seed=seed)
        z, ep = synth.make_noise(B, seed=seed + 1), synth.make_eps(B, seed=seed + 2).reshape(-1)
        return tuple(torch.from_numpy(a).to(dev) for a in (y2, x1, z, ep))

    nb = 3
    batches = [batch(1000 * rank + 10 * i) for i in range(nb)]
    noises = torch.from_numpy(np.stack([synth.make_noise(B, seed=7000 + 100 * rank + k) for k in range(10)])).to(dev)

    def step(i):
        by2 = [batches[(i + j) % nb] for j in range(5)]
        bdem = [batches[(i + j + 1) % nb] for j in range(5)]
        y2, x1, _, _ = bdem[-1]  # the generator trains on the last DEM-critic batch (TG:873-878)
        return tr.gen_iteration_device(by2, bdem, x1, y2, noises)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    dp_check = None
    if world > 1:  # before any update: all ranks hold identical weights (same seeds)
        dp_check = dp_self_check(tr, G, D1, D2, rank, world, dev, nicg, thr, B)
    for i in range(warmup):
        step(i)
    barrier()
    prof = None
    if rank == 0 and os.environ.get("DEPGAN_PROFILE_LOG"):
        from depgan_b200 import _lib
        L = _lib.lib()
        ncls = 7
        a_ms, a_fl, a_by = (C.c_double * ncls)(), (C.c_double * ncls)(), (C.c_double * ncls)()
        a_n = (C.c_longlong * ncls)()
        L.depgan_profile_begin()
        step(0)
        _lib.check(L.depgan_profile_end(a_ms, a_fl, a_by, a_n, ncls), "profile_end")
        names = ["tc3x3", "tc5x5", "tc1x1_deconv", "simt_conv", "simt_wgrad", "tc_wgrad", "tc_conv_fused_maxpool"]
        prof = {nm: {"ms": a_ms[i], "launches": int(a_n[i]),
                     "tflops": (a_fl[i] / (a_ms[i] * 1e-3) / 1e12) if a_ms[i] > 0 else None}
                for i, nm in enumerate(names)}
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses, out = step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    last = [float(v) for v in out.cpu().numpy()]
    del tr, G, D1, D2, batches, noises
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    flop_per_slice = 1013.3e9  # 10 critic updates + 10 evals + 1 G update
    pk = peaks()
    tf = steps * B * world * flop_per_slice / (ms * 1e-3) / 1e12
    return {
        "metric": "DEP-GAN train steps/sec (generator iterations: 5+5 critic updates, 10 noise evals, 1 G update)",
        "value": steps / (ms * 1e-3), "unit": "gen-iterations/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong" if global_batch else "weak", "vs_baseline": None,
        "dtype": "bf16" if args.train_precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "DEP-GAN %s two-critic training, random init, batch %d per GPU, global batch %d "
                               "(BASELINE configs[%s])" % ("PROB+FLAIR (nicg=2, T=0.5)" if nicg == 2 else "IM", B,
                                                           B * world, "3" if nicg == 2 else "2"),
                   "precision": args.train_precision,
                   "parallelism": "data parallel over %d GPU(s): NCCL all-reduce of the flat gradient bucket + "
                                  "loss partial sums" % world,
                   "l2": "per-step working set >> 126 MB L2; no explicit flush"},
        "slices_per_s": steps * B * world / (ms * 1e-3),
        "tflops_effective": tf, "frac_of_sustained_bf16_peak_per_gpu": tf / world / pk["tf_sust"],
        "last_losses": last, "dp_check": dp_check,
        "clocks": clocks, "gpu_launches": int(launches), "conv_time_per_iteration": prof,
    }


def run_train(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gb = args.global_batch if args.workload == "depgan_train_pf" and not args.batch else None
    line = measure_train(args, rank, world, dev, args.steps, args.warmup, args.workload, global_batch=gb)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION prints to stdout) goes
    # to stderr instead
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uresnet_infer",
                    choices=sorted(WORKLOADS) + ["depgan_train", "depgan_train_pf"])
    # inference legs: 'f16' (default) and 'bf16' run the same tcgen05 kernels at the same rate, f16 = IEEE-half storage
    # (DEM error ~7x smaller); the training legs keep bf16 (gradient range) unless fp32 is asked for
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the DEP-GAN train-step leg of the default run")
    ap.add_argument("--train-batch", type=int, default=32, help="per-GPU batch of the train-step leg")
    ap.add_argument("--global-batch", type=int, default=256,
                    help="global batch of the PROB+FLAIR strong-scaling leg (BASELINE configs[3])")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the sub-records of configs[0], [3], [4] and the predict(numpy) leg")
    args = ap.parse_args()
    args.train_precision = "fp32" if args.precision == "fp32" else "bf16"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload.startswith("depgan_train"):
        if args.batch:
            args.train_batch = args.batch
        run_train(args, rank, world, local_rank)
        return
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    run_gpu(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
