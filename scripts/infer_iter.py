"""One DEP-UResNet inference forward (BASELINE configs[1]) between cudaProfilerStart/Stop, for ncu
(`--profile-from-start off`).  usage: infer_iter.py [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from depgan_b200 import Gen_UNet2D, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 4, precision="bf16", max_batch=B, device=str(dev))
g.set_weights(synth.init_weights([(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in g.manifest], seed=0,
                                 trained_like=True))
x = torch.from_numpy(synth.make_flair(B, 256, 256, seed=1)[0]).to(dev)
z = torch.from_numpy(synth.make_noise(B, seed=2)).to(dev)
out = torch.empty((B, 256, 256, 4), dtype=torch.float32, device=dev)
for _ in range(2):
    g.forward_device(x, z, out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
g.forward_device(x, z, out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", float(out.sum()))
