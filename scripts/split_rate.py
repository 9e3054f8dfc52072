"""Rate of the split-half (precision='f16x3') generator forward next to the other storage formats, device-resident,
at a given batch (default 64 = the bench's configs[1] shape, softmax head).  Usage: python scripts/split_rate.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depgan_b200 import Gen_UNet2D, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
x, _ = synth.make_flair(B, 256, 256, seed=1)
z = synth.make_noise(B, seed=2)
xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
out = torch.empty((B, 256, 256, 4), dtype=torch.float32, device=dev)
ref = None
only = os.environ.get("DEPGAN_ONLY")  # e.g. DEPGAN_ONLY=f16x3 under ncu
for prec, steps in (("f16x3", 10 if not only else 1), ("f16", 20), ("fp32", 2)):
    if only and prec != only:
        continue
    g = Gen_UNet2D((256, 256, 1), (32, 1), 32, 4, precision=prec, max_batch=B)
    man = [(n.split("/")[0], n.split("/")[1], s) for n, s, _, _ in g.manifest]
    g.set_weights(synth.init_weights(man, seed=0, trained_like=True))
    for _ in range(3):
        g.forward_device(xd, zd, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.forward_device(xd, zd, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    y = out.clone()
    if prec == "fp32":
        print("f16x3 vs fp32 path max abs", float((ref - y).abs().max()))
    if ref is None:
        ref = y
    print(prec, "batch", B, "%.3f ms" % ms, "%.0f slices/s" % (B / ms * 1e3), flush=True)
    if prec == "f16x3" and os.environ.get("DEPGAN_PROFILE_LOG"):
        from depgan_b200 import _lib
        import ctypes as C
        L = _lib.lib()
        L.depgan_profile_begin()
        g.forward_device(xd, zd, out)
        ms_ = (C.c_double * 8)(); fl = (C.c_double * 8)(); by = (C.c_double * 8)(); ln = (C.c_longlong * 8)()
        L.depgan_profile_end(ms_, fl, by, ln, 8)
        print(open(os.environ["DEPGAN_PROFILE_LOG"]).read())
    del g
    torch.cuda.empty_cache()
