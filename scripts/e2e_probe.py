"""Where does the end-to-end inference path lose time against the device-only loop?  Times 40 steps of
(compute only | +H2D | +D2H | both) with the same three-stream structure as InferencePipeline."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from depgan_b200 import Gen_UNet2D, synth  # noqa: E402

B, steps = 64, 40
dev = torch.device("cuda:0")
g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 4, precision="bf16", max_batch=B, device=str(dev))
g.set_weights(synth.init_weights([(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in g.manifest], seed=0,
                                 trained_like=True))
xh = torch.from_numpy(synth.make_flair(B, 256, 256, seed=1)[0]).pin_memory()
zh = torch.from_numpy(synth.make_noise(B, seed=2)).pin_memory()
ohs = [torch.empty((B, 256, 256, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
s_in, s_run, s_out = (torch.cuda.Stream(dev) for _ in range(3))
s_more = [torch.cuda.Stream(dev) for _ in range(3)]
NSPLIT = 1
x = [torch.empty((B, 256, 256, 1), device=dev) for _ in range(2)]
z = [torch.empty((B, 32, 1), device=dev) for _ in range(2)]
o = [torch.empty((B, 256, 256, 4), device=dev) for _ in range(2)]
for k in range(2):
    x[k].copy_(xh); z[k].copy_(zh)


def run(h2d, d2h, n):
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_run = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    for i in range(n):
        k = i % 2
        if h2d:
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_run[k])
                x[k].copy_(xh, non_blocking=True); z[k].copy_(zh, non_blocking=True)
                ev_in[k].record(s_in)
        with torch.cuda.stream(s_run):
            if h2d:
                s_run.wait_event(ev_in[k])
            if d2h and i >= 2:
                s_run.wait_event(ev_out[k])
            g.forward_device(x[k], z[k], o[k])
            ev_run[k].record(s_run)
        if d2h:
            streams = [s_out] + s_more[:NSPLIT - 1]
            per = B // NSPLIT
            for j, sj in enumerate(streams):
                with torch.cuda.stream(sj):
                    sj.wait_event(ev_run[k])
                    ohs[k][j * per:(j + 1) * per].copy_(o[k][j * per:(j + 1) * per], non_blocking=True)
                    if j > 0:
                        e = torch.cuda.Event(); e.record(sj); s_out.wait_event(e)
            with torch.cuda.stream(s_out):
                ev_out[k].record(s_out)
    for s in [s_in, s_run, s_out] + s_more:
        s.synchronize()


for name, h2d, d2h, ns in (("compute only", 0, 0, 1), ("+D2H x1", 0, 1, 1), ("+D2H x2", 0, 1, 2), ("+D2H x4", 0, 1, 4),
                           ("both x1", 1, 1, 1), ("both x2", 1, 1, 2), ("both x4", 1, 1, 4), ("compute only", 0, 0, 1)):
    NSPLIT = ns
    run(h2d, d2h, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(h2d, d2h, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%-14s %.3f ms/step  %.0f slices/s" % (name, ms, B / ms * 1e3))
