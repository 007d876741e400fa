"""Build libpbd_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpbd_b200.so")
# (pbd_server.cpp is not part of the library: build_server links it against the library)
SOURCES = ["pbd_plan.cpp", "pbd_tileplan.cpp", "pbd_placement.cpp", "pbd_stream.cu", "pbd_tile.cu", "pbd_batch.cu", "pbd_capi.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [os.path.join("..", "..", "include", "pbd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other archs, no fallbacks
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                                   # belt and braces: the math uses explicit *_rn intrinsics
    "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
    "--shared", "-cudart", "static",
]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


SERVER = os.path.join(HERE, "pbd_server")
SERVER_SRC = os.path.join(CSRC, "pbd_server.cpp")


def build_server(force: bool = False) -> str:
    """The PBD1 wire server (csrc/pbd_server.cpp) linked against the in-tree library."""
    if not force and os.path.exists(SERVER) and os.path.getmtime(SERVER) >= max(os.path.getmtime(SERVER_SRC), os.path.getmtime(LIB)):
        return SERVER
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-Wpedantic", "-Werror", SERVER_SRC, "-L", HERE, "-lpbd_b200",
           "-Wl,-rpath,$ORIGIN", "-o", SERVER]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building pbd_server failed")
    return SERVER


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into cs121-softbodysim_b200/libpbd_b200.so (and the wire server
    cs121-softbodysim_b200/pbd_server on top of it); returns the library's path."""
    if not force and not stale():
        build_server()
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    build_server(force=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
