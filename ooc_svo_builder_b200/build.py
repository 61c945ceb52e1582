"""In-tree build of the CUDA library and the CLI executables (nvcc, sm_100a only).

    python -m ooc_svo_builder_b200.build

Outputs (git-ignored, but they travel to the GPU box with the snapshot):
    ooc_svo_builder_b200/lib/libsvo_b200.so      C ABI + kernels
    ooc_svo_builder_b200/bin/svo_builder         payload CLI   (reference: svo_builder)
    ooc_svo_builder_b200/bin/svo_builder_binary  geometry CLI  (reference: svo_builder_binary)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib", "libsvo_b200.so")
BIN = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",                      # float parity with the reference's non-FMA x86 build
              "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_library(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "svo_b200.h")]
    if not force and _newer(LIB, srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    _run(["nvcc", *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "svo_api.cu")])
    return LIB


def build_cli(force: bool = False) -> list[str]:
    if not os.path.isdir(HOST):
        return []
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST))] + [os.path.join(ROOT, "include", "svo_b200.h")]
    outs = []
    os.makedirs(BIN, exist_ok=True)
    for name, defs in (("svo_builder", []), ("svo_builder_binary", ["-DBINARY_VOXELIZATION"])):
        out = os.path.join(BIN, name)
        outs.append(out)
        if not force and _newer(out, srcs + [LIB]):
            continue
        _run(["g++", "-std=c++17", "-O2", "-Wall", *defs, "-I" + os.path.join(ROOT, "include"),
              os.path.join(HOST, "svo_builder_main.cpp"), "-o", out,
              "-L" + os.path.dirname(LIB), "-lsvo_b200", "-lpthread", "-Wl,-rpath,$ORIGIN/../lib"])
    return outs


def build_tools(force: bool = False) -> list[str]:
    """Stand-alone CUDA tools (not part of the product): the NVLink staging microbenchmark."""
    src = os.path.join(ROOT, "tools", "native", "p2p_probe.cu")
    out = os.path.join(ROOT, "tools", "native", "p2p_probe")
    if not os.path.exists(src):
        return []
    if force or not _newer(out, [src]):
        _run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-o", out, src])
    return [out]


def build_all(force: bool = False) -> None:
    build_library(force)
    build_cli(force)
    build_tools(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
