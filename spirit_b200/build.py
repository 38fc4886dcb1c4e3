"""Builds the C-ABI library of spirit_b200 in-tree: spirit_b200/libSpirit.so.

CUDA sources are compiled for sm_100a only (`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`); nvcc
cross-compiles without a GPU. Objects are cached under spirit_b200/build/ by mtime of the source and of every
header in the source tree. The built library is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libSpirit.so")

# experiments: SPIRIT_B200_VARIANT=name SPIRIT_B200_DEFINES="-DX=1 -DY=2" builds libSpirit_<name>.so with extra defines
VARIANT = os.environ.get("SPIRIT_B200_VARIANT", "")
EXTRA_DEFINES = os.environ.get("SPIRIT_B200_DEFINES", "").split()
if VARIANT:
    BUILD = os.path.join(HERE, "build_" + VARIANT)
    LIB = os.path.join(HERE, "libSpirit_%s.so" % VARIANT)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("SPIRIT_B200_CXX", "/usr/bin/g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
INCLUDES = ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I/usr/local/cuda/include"]
CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function"]
NVCCFLAGS = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden",
             "-ccbin", CXX] + ARCH


def _sources():
    out = []
    for sub in ("core", "api", "device"):
        d = os.path.join(CSRC, sub)
        for f in sorted(os.listdir(d)):
            if f.endswith(".cpp") or f.endswith(".cu"):
                out.append(os.path.join(d, f))
    return out


def _headers_mtime():
    newest = 0.0
    for base in (CSRC, os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".hpp", ".h", ".cuh")):
                    newest = max(newest, os.path.getmtime(os.path.join(dirpath, f)))
    return newest


def _compile(src, obj, verbose):
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCCFLAGS + EXTRA_DEFINES + INCLUDES + ["-c", src, "-o", obj]
    else:
        cmd = [CXX] + CXXFLAGS + INCLUDES + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compilation failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose and r.stderr.strip():
        print(r.stderr, flush=True)


def build(verbose=False, force=False):
    """Compile what is out of date and link spirit_b200/libSpirit.so. Returns the library path."""
    os.makedirs(BUILD, exist_ok=True)
    hdr = _headers_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(BUILD, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(lambda j: _compile(j[0], j[1], verbose), jobs))
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-ccbin", CXX, "-lcudart_static", "-lpthread", "-ldl", "-lrt",
                                                            "-Xlinker", "--no-undefined"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
