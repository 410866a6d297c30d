"""In-tree build of ``libsustaingym_b200.so`` (hand-written sm_100a CUDA behind a C ABI).

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting
``.so`` is git-ignored but travels to the GPU box with the repo snapshot.  Each ``.cu`` is
compiled to an object in parallel and relinked only when a source/header is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
BUILD_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.environ.get("SG_LIB_PATH", os.path.join(PKG_DIR, "libsustaingym_b200.so"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
] + os.environ.get("SG_NVCC_EXTRA", "").split()


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(d) for d in deps)


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [NVCC, *NVCC_FLAGS, "-Xptxas", "-v", "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every CUDA source for sm_100a and links the shared library. Returns its path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = _sources()
    hdr_time = _deps_mtime()
    jobs = []
    for src in srcs:
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        jobs.append((src, obj, stale))
    todo = [(s, o) for s, o, stale in jobs if stale]
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            logs = list(ex.map(lambda so: _compile(so[0], so[1], verbose), todo))
        with open(os.path.join(BUILD_DIR, "ptxas.log"), "a") as f:
            for (s, _), log in zip(todo, logs):
                f.write(f"==== {os.path.basename(s)}\n{log}\n")
    objs = [o for _, o, _ in jobs]
    if todo or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
