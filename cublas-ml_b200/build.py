"""Build the CUDA engine (libb2nn.so) and the C++ façade (libcublasNN_b200.so) in-tree.

    python cublas-ml_b200/build.py [--force]

nvcc cross-compiles for sm_100a without a GPU.  The shared objects land in cublas-ml_b200/lib/ (git-ignored, shipped to
the GPU box with the repo snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "build")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA, "bin", "nvcc")
CXX = "/usr/bin/g++"

CU_SOURCES = ["engine.cu", "gemm_simt.cu", "gemm_tc.cu", "elementwise.cu", "tail.cu", "skinny.cu", "micro.cu", "mlp3.cu", "p2p.cu", "builtins.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wall", "-ccbin", CXX, "-I", os.path.join(ROOT, "include")]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)


def build(force: bool = False, verbose: bool = False) -> dict:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objs, jobs = [], []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + NVCC_FLAGS + extra + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(_run, jobs))
    lib = os.path.join(LIB, "libb2nn.so")
    if force or jobs or _newer(lib, objs):
        # the arch is named on the link line too: nvcc's device-link stub otherwise defaults to sm_52
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-cudart", "static", "-L", os.path.join(CUDA, "lib64"), "-lcurand",
             "-ldl", "-Xlinker", f"-rpath={os.path.join(CUDA, 'lib64')}", "-ccbin", CXX])
    out = {"libb2nn": lib}
    facade_src = os.path.join(HOST, "cublasNN.cpp")
    if os.path.exists(facade_src):
        flib = os.path.join(LIB, "libcublasNN_b200.so")
        if force or _newer(flib, [facade_src, lib] + headers):
            _run([CXX, "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", "-I", os.path.join(ROOT, "include"), facade_src,
                  "-o", flib, "-L", LIB, "-lb2nn", "-Wl,-rpath,$ORIGIN"])
        out["facade"] = flib
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
