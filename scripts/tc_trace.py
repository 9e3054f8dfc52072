"""Role trace of conv_tc_kernel (CTA 0) for one kbench case, with a library built by
`UNITS="conv_tc_k3 conv_tc" scripts/build_variant.sh trace -DDG_DBG_TRACE` and DEPGAN_B200_LIB=build_ab/libtrace.so.
Prints, per work item of CTA 0: producer (top, first emptyA ok, all issued), issuer 0 (top, accEmpty ok, first fullA ok,
committed), epilogue warp 0 (start, accFull ok, item end, ...), in clocks relative to the producer's first item."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kbench  # noqa: E402

L = kbench.L
buf = torch.zeros(3 * 40 * 8, dtype=torch.int64, device="cuda")
assert L.depgan_dbg_set_trace(C.c_void_p(buf.data_ptr())) == 0
for mk in kbench.CASES:
    name, run, flops, nbytes, keep = mk()
    if not any(s in name for s in sys.argv[1:]):
        del keep
        continue
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf.zero_()
    run()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(3, 40, 8).astype("int64") & 0xFFFFFFFF
    t0 = int(t[0, 0, 0])
    print("== %s" % name)
    for i in range(40):
        rel = lambda v: (int(v) - t0) & 0xFFFFFFFF
        print("%3d  P %s | M %s | E %s" % (i, " ".join("%7d" % rel(v) for v in t[0, i, :3]),
                                          " ".join("%7d" % rel(v) for v in t[1, i, :4]),
                                          " ".join("%7d" % rel(v) for v in t[2, i, :8])))
    del keep
