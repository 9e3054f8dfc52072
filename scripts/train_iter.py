"""One DEP-GAN generator iteration (BASELINE configs[2], batch 32) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from depgan_b200 import Dis_C2D_FCN1, Gen_UNet2D, synth  # noqa: E402
from depgan_b200.trainer import DepGanTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
G = Gen_UNet2D((256, 256, 1), (32, 1), 32, 1, precision="bf16", max_batch=B, device=str(dev), training=True, seed=0)
D1 = Dis_C2D_FCN1((256, 256, 1), precision="bf16", max_batch=3 * B, device=str(dev), training=True, seed=1)
D2 = Dis_C2D_FCN1((256, 256, 1), precision="bf16", max_batch=3 * B, device=str(dev), training=True, seed=2)
tr = DepGanTrainer(G, D1, D2, 0.178)
import os
if not os.environ.get("DEPGAN_NO_BATCHED_EVAL"):
    tr.enable_batched_eval(10)


def batch(seed):
    x1, y2, _ = synth.make_im_pair(B, 256, 256, nicg=1, thr=0.178, seed=seed)
    z, ep = synth.make_noise(B, seed=seed + 1), synth.make_eps(B, seed=seed + 2).reshape(-1)
    return tuple(torch.from_numpy(a).to(dev) for a in (y2, x1, z, ep))


bs = [batch(10 * i) for i in range(3)]
noises = torch.from_numpy(np.stack([synth.make_noise(B, seed=7000 + k) for k in range(10)])).to(dev)


def step(i):
    by2 = [bs[(i + j) % 3] for j in range(5)]
    bdem = [bs[(i + j + 1) % 3] for j in range(5)]
    y2, x1, _, _ = bdem[-1]
    return tr.gen_iteration_device(by2, bdem, x1, y2, noises)


step(0)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
