"""Build libmavlm.so (CUDA kernels + C ABI) for sm_100a, in-tree, with nvcc.

    python memory-augmented-vlm_b200/build.py [--force] [--verbose]

Used by ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.  Objects are cached under
``csrc/build/`` and rebuilt when the source (or a header) is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmavlm.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    f"-I{INCLUDE}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmavlm.so cannot be built (there is no non-CUDA fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(obj: str, src: str, headers) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(p) > t for p in [src, *headers])


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(BUILD, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, src, headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj[:-2] + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        return src, log

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, log in ex.map(compile_one, jobs):
            if verbose:
                print(f"--- {os.path.basename(src)}\n{log}")
    if jobs or force or not os.path.exists(LIB):
        # the CUDA runtime is linked as a SHARED library (libcudart.so.12): in a PyTorch process the already-loaded
        # runtime is reused; stand-alone (C / ctypes hosts) the rpath finds the toolkit's copy.  A static runtime would
        # embed its whole symbol table (every entry point's name as a string) in the shipped artefact.
        cudart_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64")
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
               "-Xlinker", f"-rpath={cudart_dir}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
