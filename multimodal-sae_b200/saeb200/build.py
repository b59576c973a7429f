"""Build `lib/libsaeb200.so` (hand-written sm_100a kernels + the C ABI of include/saeb200.h) with nvcc.

The library is built IN-TREE so that it travels to the GPU box with the repo snapshot; it links the CUDA runtime
statically and resolves the one driver entry point it needs (cuTensorMapEncodeTiled) at run time, so it loads -- and
its symbols can be checked -- on a machine without a GPU driver.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(LIB_DIR, "libsaeb200.so")
SOURCES = ["capi.cu", "encode_topk.cu", "decode.cu", "pack.cu", "coo_scan.cu", "refine.cu", "exchange.cu", "decode_bwd.cu", "probe.cu"]
# every header a translation unit may include: common helpers, the device code (kernels_*.cuh) and the C ABI
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "saeb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]

    def compile_one(src: str):
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(BUILD_DIR, src.replace(".cu", ".log"))
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB_PATH, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
