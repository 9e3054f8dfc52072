"""What the host can move when N ranks copy at once: every rank repeats the end-to-end step's transfers WITHOUT any kernel
(16.8 MB host->device, 67.1 MB device->host, pinned buffers, two streams) between barriers; rank 0 prints per-rank and
aggregate GB/s and the slices/s ceiling they imply for bench.py's `e2e` (1 MiB of float32 softmax maps per slice).
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
     scripts/pcie_bw_multi.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if os.environ.get("DEPGAN_NO_NUMA") is None:
    try:
        from depgan_b200.infer import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)
    except Exception as e:  # measurement script: report and go on unbound
        print("rank %d: NUMA binding not applied (%s)" % (rank, e), flush=True)
B = 64
n_out, n_in = B * 256 * 256 * 4, B * (256 * 256 + 32)
d_out, d_in = torch.empty(n_out, device=dev), torch.empty(n_in, device=dev)
h_out, h_in = torch.empty(n_out).pin_memory(), torch.empty(n_in).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for mode in ("both", "d2h_only"):
    for _ in range(3):
        step()
    sync()
    reps = 40
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode == "both":
            step()
        else:
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sync()
    if rank == 0:
        sec = float(dt[0])
        d2h = reps * n_out * 4 / sec / 1e9
        h2d = reps * n_in * 4 / sec / 1e9 if mode == "both" else 0.0
        print("%d rank(s), %s: per rank D2H %.1f GB/s, H2D %.1f GB/s; aggregate %.1f GB/s; e2e ceiling from the copies "
              "alone %.0f slices/s (all ranks)" % (world, mode, d2h, h2d, world * (d2h + h2d), world * reps * B / sec), flush=True)
if world > 1:
    dist.destroy_process_group()
