"""Per-role event clocks of CTA 0 of one conv_row_kernel launch (needs the DG_ROW_TRACE variant build:
UNITS=conv_row scripts/build_variant.sh rowtrace -DDG_ROW_TRACE; DEPGAN_B200_LIB=build_ab/librowtrace.so)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kbench  # noqa: E402

L = kbench.L
buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
import os
if os.environ.get("TRACE_ROWG"):  # the generic kernel: UNITS=conv_rowg scripts/build_variant.sh rowgtrace -DDG_ROWG_TRACE
    L.depgan_dbg_set_row_trace = L.depgan_dbg_set_rowg_trace
L.depgan_dbg_set_row_trace.argtypes = [C.c_void_p]
for mk in kbench.CASES:
    name, run, flops, nbytes, keep = mk()
    if not any(s == name for s in sys.argv[1:]):
        continue
    run(); run()
    torch.cuda.synchronize()
    assert L.depgan_dbg_set_row_trace(C.c_void_p(buf.data_ptr())) == 0
    buf.zero_()
    run()
    torch.cuda.synchronize()
    L.depgan_dbg_set_row_trace(C.c_void_p(0))
    t = buf.cpu().view(4, 64, 8)
    if os.environ.get("TRACE_ROWG"):  # 32-bit clocks from shared memory: unwrap relative to the first event
        t = ((t - t[0, 0, 0]) & 0xFFFFFFFF) * (t != 0) + (t != 0)
    t0 = int(t[0, 0, 0])
    print("== %s  P: top, emptyA ok | M: top, blocks acquired, fullA ok, committed | E0/E1 (warps 0 / 4): top, rowDone ok, "
          "ld done, zeroed+arrived, math done, store slot free, staged+fenced, store issued" % name)
    for i in range(64):
        f = lambda r, n: " ".join("%7d" % (int(v) - t0 if int(v) else -1) for v in t[r, i, :n])
        print("%3d  P %s | M %s | E0 %s | E1 %s" % (i, f(0, 2), f(1, 5 if os.environ.get("TRACE_ROWG") else 4), f(2, 8), f(3, 8)))
