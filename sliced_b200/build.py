"""
In-tree build of libsliced_b200.so (the C-ABI library) with nvcc for sm_100a.

    python -m sliced_b200.build [--force]

Every .cu under sliced_b200/csrc is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false
(-fmad=false: the reference never contracts a*b+c, so the element-wise kernels must not either; kernels that want an
FMA say fmaf explicitly) and linked with a static cudart so that the library loads on a box without a GPU driver.
The .so stays in-tree (git-ignored) so it travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsliced_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=default", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    srcs = []
    for d in (CSRC, HOST):
        if os.path.isdir(d):
            srcs += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cpp"))]
    return srcs


def _deps_mtime():
    m = 0.0
    for d in (CSRC, HOST, os.path.join(HERE, "..", "include")):
        if not os.path.isdir(d):
            continue
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src: str, force: bool) -> tuple[str, str]:
    obj = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
    log = obj + ".log"
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= _deps_mtime()):
        return obj, ""
    cmd = [nvcc(), *NVCC_FLAGS, "-x", "cu", "-c", src, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{p.stdout[-6000:]}")
    return obj, p.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, out in results:
            if out:
                print(out)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc(), "-shared", "-cudart", "static", "-o", LIB, *objs, "-ldl", "-lpthread", "-lrt"]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout[-4000:]}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
