"""Builds real_b200/libreal_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libreal_gpu.so")
SOURCES = ["real_gpu.cu"]
HEADERS = ["common.cuh", "prims.cuh", "index.cuh", "scan.cuh", "post.cuh", "ingest.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas=-v",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "real_gpu.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or p.returncode != 0:
        sys.stderr.write(p.stdout)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed building libreal_gpu.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
BIN_DIR = os.path.join(HERE, "bin")
HOST_BIN = os.path.join(BIN_DIR, "real")
HOST_DUMP = os.path.join(BIN_DIR, "real_host_dump")


def build_host(force: bool = False) -> str:
    """Compiles the C++ host driver (the `real` command line on the C ABI) and its test helper."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("real_host.cpp", "real_host.hpp", "real_main.cpp", "real_host_dump.cpp")]
    newest = max(os.path.getmtime(f) for f in srcs + [os.path.join(HERE, "..", "include", "real_gpu.h")])
    if not force and all(os.path.exists(b) and os.path.getmtime(b) >= newest for b in (HOST_BIN, HOST_DUMP)) and os.path.getmtime(HOST_BIN) >= os.path.getmtime(LIB):
        return HOST_BIN
    os.makedirs(BIN_DIR, exist_ok=True)
    gxx = shutil.which("g++") or "g++"
    common = [gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread"]
    subprocess.check_call(common + ["-o", HOST_BIN, os.path.join(HOST_DIR, "real_main.cpp"), os.path.join(HOST_DIR, "real_host.cpp"),
                                    "-L" + HERE, "-lreal_gpu", "-Wl,-rpath,$ORIGIN/.."])
    subprocess.check_call(common + ["-o", HOST_DUMP, os.path.join(HOST_DIR, "real_host_dump.cpp"), os.path.join(HOST_DIR, "real_host.cpp"),
                                    "-L" + HERE, "-lreal_gpu", "-Wl,-rpath,$ORIGIN/.."])
    return HOST_BIN


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
    print(build_host(force=True))
