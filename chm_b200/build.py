"""Build libpbsm3d_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpbsm3d_b200.so")
SOURCES = ["pbsm3d_capi.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [os.path.join("..", "..", "include", "pbsm3d.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def nccl_paths():
    """Prefer the NCCL that ships with torch (nvidia-nccl wheel): torch's libtorch_cuda needs its symbols, and the
    dynamic loader shares one libnccl.so.2 per process whichever of us loads first.  Fall back to the system one."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        base = list(spec.submodule_search_locations)[0]
        inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
        if os.path.exists(os.path.join(lib, "libnccl.so.2")) and os.path.exists(os.path.join(inc, "nccl.h")):
            return inc, lib
    except Exception:
        pass
    return None, None


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    inc, libdir = nccl_paths()
    nccl = ["-lnccl"]
    if inc:
        nccl = ["-I", inc, "-L", libdir, "-l:libnccl.so.2", "-Xlinker", "-rpath", "-Xlinker", libdir]
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES], *nccl]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building libpbsm3d_b200.so")
    os.replace(tmp, LIB)  # atomic: concurrent builders (the ranks of a torchrun job) never load a half-written library
    if verbose:
        sys.stderr.write(res.stderr)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(res.stderr)
    return LIB


HOST = os.path.join(HERE, "host")
DRIVER = os.path.join(HOST, "standalone_driver")


def build_adaptor(force: bool = False) -> str:
    """g++ build of the CHM adaptor (host/PBSM3D_gpu.cpp) over the stand-in CHM types (host/chm_shim.hpp) plus the
    stand-alone driver, linked against libpbsm3d_b200.so.  Inside CHM the same PBSM3D_gpu.cpp is compiled against
    CHM's own headers instead."""
    srcs = [os.path.join(HOST, f) for f in ("PBSM3D_gpu.cpp", "snow_slide_gpu.cpp", "standalone_driver.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in ("PBSM3D_gpu.hpp", "snow_slide_gpu.hpp", "chm_shim.hpp")] + [os.path.join(HERE, "..", "include", "pbsm3d.h"), LIB]
    if not force and os.path.exists(DRIVER) and all(os.path.getmtime(d) <= os.path.getmtime(DRIVER) for d in deps):
        return DRIVER
    cmd = ["g++", "-std=c++17", "-O2", "-fopenmp", "-DPBSM3D_GPU_STANDALONE", "-I", os.path.join(HERE, "..", "include"), "-I", HOST,
           *srcs, "-o", DRIVER, "-L", HERE, "-l:libpbsm3d_b200.so", "-Wl,-rpath," + HERE]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the stand-alone adaptor driver")
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_adaptor(force="--force" in sys.argv))
