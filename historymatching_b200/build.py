"""Build libhm_b200.so (in-tree) with nvcc for sm_100a.

    python -m historymatching_b200.build [--force] [--verbose]

The shared library is the product's C ABI (include/hm_b200.h).  It is built
in-tree so that it travels with the repo snapshot to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhm_b200.so")
SOURCES = ["hm_api.cu", "hm_sim.cu", "hm_pressure.cu", "hm_small.cu", "hm_gemm.cu", "hm_analysis.cu"]
HEADERS = [os.path.join(CSRC, "hm_common.cuh"), os.path.join(CSRC, "hm_sim_common.cuh"),
           os.path.join(CSRC, "hm_mg_onchip.cuh"),
           os.path.join(HERE, "..", "include", "hm_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhm_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc_path())), "lib64")
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo", *os.environ.get("HM_NVCC_EXTRA", "").split(),
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
        "-shared",
        "-o", LIB,
        *[os.path.join(CSRC, s) for s in SOURCES],
        f"-L{cuda_lib}", "-lcusolver",
        "-Xlinker", f"-rpath={cuda_lib}",
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
