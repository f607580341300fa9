"""Builds point2cyl_b200/libp2c.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libp2c.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"] + os.environ.get("P2C_NVCC_FLAGS", "").split()   # e.g. -DP2C_SS_DEBUG (tools/igr_exp.sh)
OBJ = os.path.join(PKG, "_build")     # per-source objects (git- and gpurun-ignored): only stale sources recompile


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    # link into a private name, then rename: a concurrent reader (another rank of a torchrun launch) never maps a
    # half-written library, and two concurrent builders cannot interleave their output
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    hdr_t = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, None
        tmp_o = f"{obj}.{os.getpid()}.tmp"
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", tmp_o, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp_o):
                os.remove(tmp_o)
            return None, res.stdout + res.stderr
        os.replace(tmp_o, obj)
        return obj, res.stderr if verbose else None

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources()))
    for obj, log in results:
        if log:
            sys.stderr.write(log)
    if any(obj is None for obj, _ in results):
        raise RuntimeError("nvcc failed building libp2c.so")
    tmp = f"{LIB}.{os.getpid()}.tmp"
    res = subprocess.run([NVCC, "-shared", "-o", tmp] + [o for o, _ in results] + ["-lcuda"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("linking libp2c.so failed")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
