"""In-tree build of the native library (nvcc, sm_100a only).  `python -m metricsfm_b200.build [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(CSRC, "libmsfm_match.so")
SOURCES = ["msfm_api.cu"]
HEADERS = ["match_kernel.cuh", "aux_kernels.cuh", "geo_kernels.cuh", "sm100_ptx.cuh", os.path.join("..", "..", "include", "msfm_match.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "-ccbin", "/usr/bin/g++",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libmsfm_match.so (skipped when up to date).  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MSFM_NVCC_EXTRA", "").split()  # kernel experiments, e.g. -DMSFM_PRODUCER_AUX=7
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmsfm_match.so")
    with open(os.path.join(CSRC, "ptxas_info.txt"), "w") as f:
        f.write(proc.stderr)
    return LIB_PATH


HOST_DIR = os.path.join(PKG_DIR, "host")
SHIM_BIN = os.path.join(HOST_DIR, "shim_selftest")


def build_host_shim(force: bool = False) -> str:
    """Compile the C++ mirror of the reference's matcher interface + its self-test driver against the C ABI library."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("shim_selftest.cc", "feature_matching_b200.cc")]
    deps = srcs + [os.path.join(HOST_DIR, f) for f in ("feature_matching_b200.h", "cv_standin.h")] + [LIB_PATH]
    if not force and os.path.exists(SHIM_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(SHIM_BIN) for d in deps):
        return SHIM_BIN
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", SHIM_BIN] + srcs + ["-L" + CSRC, "-lmsfm_match", "-Wl,-rpath,$ORIGIN/../csrc"]
    subprocess.check_call(cmd)
    return SHIM_BIN


STORE_LIB = os.path.join(HOST_DIR, "libmsfm_store.so")
GRAPH_LIB = os.path.join(HOST_DIR, "libmsfm_graph.so")


def _newer(target: str, deps) -> bool:
    return os.path.exists(target) and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps if os.path.exists(d))


def build_host_libs(force: bool = False):
    """libmsfm_store.so (the reference's on-disk formats, host only) and libmsfm_graph.so (the fine-matching-graph
    driver = store + GPU matcher)."""
    inc = os.path.join(PKG_DIR, "..", "include")
    store_src = os.path.join(HOST_DIR, "msfm_store.cc")
    graph_src = os.path.join(HOST_DIR, "msfm_graph.cc")
    common = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared"]
    if force or not _newer(STORE_LIB, [store_src, os.path.join(inc, "msfm_store.h"), os.path.abspath(__file__)]):
        subprocess.check_call(common + ["-o", STORE_LIB, store_src])
    if force or not _newer(GRAPH_LIB, [graph_src, store_src, LIB_PATH, os.path.join(inc, "msfm_graph.h"), os.path.abspath(__file__)]):
        subprocess.check_call(common + ["-o", GRAPH_LIB, graph_src, store_src, "-L" + CSRC, "-lmsfm_match", "-Wl,-rpath,$ORIGIN/../csrc"])
    return STORE_LIB, GRAPH_LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
    print(build_host_shim(force="--force" in sys.argv))
    print(*build_host_libs(force="--force" in sys.argv))
