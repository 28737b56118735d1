"""Build helpers: compile the CUDA library (product) in-tree with nvcc for sm_100a.

The oracle (test infrastructure) has its own Makefile under oracle/; `build_oracle` just runs it.
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "ndsmf.so")
SOURCES = ["kernels.cu", "pool.cu", "hostsink.cu", "mg.cu", "mg_batch.cu", "nccl_comm.cu", "peer.cu", "vecpot.cu", "abi.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "pool.hpp", "hostsink.hpp", "mg.hpp", "vecpot.hpp", "sym_alloc.hpp", os.path.join(REPO_DIR, "include", "ndsm_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the reference is built without FMA contraction (gfortran -O3, baseline x86-64): keep the
    # same rounding on the device (and on the host side of the library)
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared", "-cudart", "static",
    "-ldl",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile ndsm_b200/lib/ndsmf.so (all CUDA kernels + C ABI) for sm_100a."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB_PATH, deps):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()   # link elsewhere, then rename: a snapshot never sees half a library
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # nvcc must use the system g++ (the CC/CXX wrappers exported in this image lack libgomp specs etc.)
    env.pop("CC", None)
    env.pop("CXX", None)
    try:
        subprocess.run(cmd, check=True, cwd=CSRC, env=env)
        os.replace(tmp, LIB_PATH)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB_PATH


def build_oracle(force=False):
    """Compile the CPU oracle (oracle/_build/libndsm_oracle.so).  Test infrastructure only."""
    odir = os.path.join(REPO_DIR, "oracle")
    out = os.path.join(odir, "_build", "libndsm_oracle.so")
    if force or _stale(out, [os.path.join(odir, "ndsm_oracle.c"), os.path.join(odir, "Makefile")]):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", odir, "-B" if force else "-s"], check=True, env=env)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
