"""Build libegx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m emotiongestures_b200.build [--force] [--verbose]

The shared library links the CUDA runtime statically and has no torch / Python
dependency: it is the C-ABI product of include/egx.h.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libegx.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, attribution: bool | None = None) -> str:
    """attribution=True (or EGX_ATTRIBUTION=1 in the environment of the BUILD) compiles the EGX_* variant /
    attribution switches in; the shipped library has none (csrc/egx_common.cuh: env_switch)."""
    nvcc = _nvcc()
    if attribution is None:
        attribution = os.environ.get("EGX_ATTRIBUTION") == "1"
    flags = NVCC_FLAGS + (["-DEGX_ATTRIBUTION"] if attribution else [])
    stamp = os.path.join(OBJ_DIR, ".attribution")
    os.makedirs(OBJ_DIR, exist_ok=True)
    was = os.path.exists(stamp)
    if was != bool(attribution):
        force = True
        if attribution:
            open(stamp, "w").close()
        else:
            os.remove(stamp)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "egx.h"))
    srcs = sources()
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, [src] + headers):
            return ""
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if verbose:
        for log in logs:
            sys.stderr.write(log)
    if force or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                 attribution=True if "--attribution" in sys.argv else None)
    print(path)
