"""Builds libadaptigraph_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m adaptigraph_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels with the repo
snapshot to the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadaptigraph_b200.so")
SOURCES = ["api.cu", "graph_build.cu", "sampling.cu", "rewards.cu", "forward.cu", "tc_forward.cu", "train.cu", "tc_wgrad.cu", "optim.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "adaptigraph_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """Compiles every translation unit for sm_100a (in parallel, one nvcc per source) and links the shared object."""
    if not force and not _stale() and out == LIB:
        return LIB
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]
    with tempfile.TemporaryDirectory(prefix="agx_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            res = subprocess.run([nvcc] + compile_flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
            return src, obj, res
        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
            results = list(pool.map(compile_one, SOURCES))
        failed = False
        for src, _, res in results:
            if verbose or res.returncode:
                sys.stderr.write(res.stdout + res.stderr)
            failed = failed or res.returncode != 0
        if failed:
            raise RuntimeError("nvcc failed building " + os.path.basename(out))
        link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] +
                              [obj for _, obj, _ in results] + ["-o", out], capture_output=True, text=True)
        if verbose or link.returncode:
            sys.stderr.write(link.stdout + link.stderr)
        if link.returncode:
            raise RuntimeError("nvcc failed linking " + os.path.basename(out))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
