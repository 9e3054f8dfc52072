"""Error statistics of the bf16 tcgen05 generator vs the fp32 CPU oracle at 256x256 (run on the GPU box)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from depgan_b200 import Gen_UNet2D, synth
from tests import util
for trained in (False, True):
    for seed in (11, 12):
        P = util.gen_weights(1, 1, seed=seed, trained_like=trained)
        x, _, _ = synth.make_im_pair(4, 256, 256, seed=1)
        z = synth.make_noise(4, seed=2)
        g = Gen_UNet2D((256, 256, 1), precision="bf16", max_batch=4)
        g.set_weights(P)
        got = g.predict([x, z])
        want = util.oracle_gen(P, x, z, dtype=torch.float32)
        e = np.abs(got - want).ravel()
        print("trained_like=%s seed=%d  max=%.4g p99.99=%.4g p99=%.4g mean=%.4g  |dem|max=%.3g" % (
            trained, seed, e.max(), np.quantile(e, 0.9999), np.quantile(e, 0.99), e.mean(), np.abs(want).max()))
