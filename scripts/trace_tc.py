"""Per-role event clocks of CTA 0 of one conv_tc_kernel<3,...> launch (needs a DG_DBG_TRACE variant build:
UNITS=conv_tc_k3 scripts/build_variant.sh trace -DDG_DBG_TRACE; DEPGAN_B200_LIB=build_ab/libtrace.so)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kbench  # noqa: E402

L = kbench.L
buf = torch.zeros(3 * 40 * 8, dtype=torch.int64, device="cuda")
L.depgan_dbg_set_trace.argtypes = [C.c_void_p]
for mk in kbench.CASES:
    name, run, flops, nbytes, keep = mk()
    if not any(s in name for s in sys.argv[1:]):
        continue
    run(); run()
    torch.cuda.synchronize()
    assert L.depgan_dbg_set_trace(C.c_void_p(buf.data_ptr())) == 0
    buf.zero_()
    run()
    torch.cuda.synchronize()
    L.depgan_dbg_set_trace(C.c_void_p(0))
    t = buf.cpu().view(3, 40, 8)
    t0 = int(t[0, 0, 0])
    print("== %s (clocks relative to the producer's first item; P: top, emptyA ok, issued | M: top, accEmpty ok, "
          "fullA ok, committed | E: start, accFull ok, item end, first group landed, last group staged, store drained, barrier passed, side inputs of the first chunk ready)" % name)
    for i in range(0, 40):
        if int(t[0, i, 0]) == 0:
            break
        f = lambda r, n: " ".join("%7d" % ((int(v) - t0) & 0xFFFFFFFF) for v in t[r, i, :n])
        print("%3d  P %s | M %s | E %s" % (i, f(0, 3), f(1, 4), f(2, 8)))
