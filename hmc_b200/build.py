"""Builds the CUDA extension in-tree: hmc_b200/libhmc_b200.so (sm_100a only, nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libhmc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # alpha / likelihood products must round like the reference's (no FMA contraction)
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(os.path.dirname(HERE), "include", "hmc_b200.h")]


HOST_DIR = os.path.join(HERE, "host")
HOST_SO = os.path.join(HERE, "libhmc_host.so")
CLI = os.path.join(HERE, "bin", "hmc")


def build_host(force: bool = False) -> str:
    """libhmc_host.so (host classes + flat C API) and bin/hmc (the CLI); both link libhmc_b200.so next to them."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("hmc_host.cpp", "hmc_host.hpp", "host_capi.cpp", "hmc_cli.cpp")]
    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++17", "-O3", "-fPIC", "-pthread", "-ffp-contract=off", "-Wall", "-Wno-sign-compare"]
    if force or not os.path.exists(HOST_SO) or any(os.path.getmtime(s) > os.path.getmtime(HOST_SO) for s in srcs + [SO]):
        subprocess.check_call([cxx] + flags + ["-shared", "-o", HOST_SO, srcs[0], srcs[2], "-L" + HERE, "-lhmc_b200", "-Wl,-rpath,$ORIGIN"])
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    if force or not os.path.exists(CLI) or any(os.path.getmtime(s) > os.path.getmtime(CLI) for s in srcs + [SO]):
        subprocess.check_call([cxx] + flags + ["-o", CLI, srcs[0], srcs[3], "-L" + HERE, "-lhmc_b200", "-Wl,-rpath,$ORIGIN/.."])
    return HOST_SO


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    if not force and os.path.exists(SO) and all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("HMC_B200_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "capi.cu")]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
