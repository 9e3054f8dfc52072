"""Builds csrc/*.cu into dep-gan-im_b200/libdepgan_b200.so with nvcc for sm_100a (in-tree, no torch involved)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libdepgan_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = sorted(CSRC.glob("*.cu"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include" / "depgan_b200.h"]
    stamp = HERE / "build" / "stamp"
    dg = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dg:
        return LIB
    if not Path(NVCC).exists():
        if LIB.exists():  # GPU box without a toolkit change: use the shipped library
            return LIB
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, LIB))
    (HERE / "build").mkdir(exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = HERE / "build" / (s.stem + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC, *FLAGS, "-c", str(s), "-o", str(o)], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s.name, out))
        failed |= p.returncode != 0
    (HERE / "build" / "nvcc.log").write_text("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see %s" % (HERE / "build" / "nvcc.log"))
    subprocess.check_call([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart_static", "-ldl", "-lrt",
                           "-lpthread"])
    stamp.write_text(dg)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
