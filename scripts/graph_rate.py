"""forward_device vs forward_graph (CUDA-graph replay of the same forward) at several batch sizes, device-resident.
Usage: python scripts/graph_rate.py [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depgan_b200 import Gen_UNet2D, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
dev = torch.device("cuda:0")
for B in (1, 4, 16, 64):
    x, _ = synth.make_flair(B, 256, 256, seed=1)
    z = synth.make_noise(B, seed=2)
    xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
    out = torch.empty((B, 256, 256, 4), dtype=torch.float32, device=dev)
    out2 = torch.empty_like(out)
    g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 4, precision=prec, max_batch=B)
    man = [(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in g.manifest]
    g.set_weights(synth.init_weights(man, seed=0, trained_like=True))
    res = {}
    for name, fn, o in (("launches", g.forward_device, out), ("graph", g.forward_graph, out2)):
        for _ in range(5):
            fn(xd, zd, o)
        torch.cuda.synchronize()
        steps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(xd, zd, o)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / steps
    same = bool(torch.equal(out, out2))
    print("batch %3d  launches %.3f ms (%.0f slices/s)   graph %.3f ms (%.0f slices/s)   identical %s" %
          (B, res["launches"], B / res["launches"] * 1e3, res["graph"], B / res["graph"] * 1e3, same), flush=True)
    del g
    torch.cuda.empty_cache()
