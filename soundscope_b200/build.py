"""Builds soundscope_b200/libsoundscope_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree).

    python -m soundscope_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree to the GPU box.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libsoundscope_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "soundscope_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return hdrs


def _compile(src, verbose):
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
    extra = os.environ.get("SSB_NVCC_EXTRA", "").split()
    cmd = [NVCC, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources()
    newest = max(os.path.getmtime(p) for p in srcs + _deps())
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    cmd = [NVCC, "-shared", "-cudart", "static", "-ccbin", HOST_CXX, "-o", SO, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
