"""Per-tensor gradient agreement of the bf16 train graphs vs the fp64 oracle (run on the GPU box)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.test_gpu_train import _setup, _cos

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
tr, ora, (x1, y2, z, ep) = _setup(H, 2, prec)
for which, name in ((0, "netD_y2_train"), (1, "netD_dem_train")):
    got = getattr(tr, name)([y2, x1, z, ep], update=False)
    want = getattr(ora, name)([y2, x1, z, ep], update=False)
    print(name, got, want, "gp", tr.last_gp, ora.last_gp)
    D = tr.Dy2 if which == 0 else tr.Ddem
    g = D.get_grads()
    for k, v in ora.last_grads.items():
        w = v.numpy()
        print("   %-28s cos=%.4f  |ours|=%.3e |oracle|=%.3e" % (k, _cos(g[k], w), np.linalg.norm(g[k]), np.linalg.norm(w)))
got = tr.netG_train([x1, y2, z], update=False)
want = ora.netG_train([x1, y2, z], update=False)
print("netG_train", got, want)
g = tr.G.get_grads()
for k, v in ora.last_grads.items():
    w = v.numpy()
    if "kernel" in k or "gamma" in k:
        print("   %-34s cos=%.4f  |ours|=%.3e |oracle|=%.3e" % (k, _cos(g[k], w), np.linalg.norm(g[k]), np.linalg.norm(w)))
