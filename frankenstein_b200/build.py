"""In-tree build of the C-ABI library ``libfk_b200.so`` (sm_100a only).

``python -m frankenstein_b200.build`` or ``frankenstein_b200.build.build()``.  Each ``csrc/*.cu`` is
compiled with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles without a
GPU) and linked into ``frankenstein_b200/libfk_b200.so``; the built library is git-ignored but
travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libfk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", os.path.join(os.path.dirname(HERE), "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the sm_100a library cannot be built")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), variant: str = "") -> str:
    """Build the library.  ``defines`` / ``variant`` produce an experiment build ``libfk_b200_<variant>.so`` (own object
    directory) for kernel diagnosis; ``FK_LIB_PATH`` makes ``_lib`` load it.  The default build is the product."""
    OBJ = os.path.join(CSRC, "build" + ("_" + variant if variant else ""))
    LIB = os.path.join(HERE, "libfk_b200" + ("_" + variant if variant else "") + ".so")
    NVCC_FLAGS = list(globals()["NVCC_FLAGS"]) + ["-D" + d for d in defines]
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "fk_b200.h"))
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, file=sys.stderr)

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    print(build(force="--force" in sys.argv, verbose=True, defines=_defs, variant=_var))
