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
SOURCES = ["hm_api.cu", "hm_sim.cu", "hm_transport.cu", "hm_pressure.cu", "hm_small.cu", "hm_gemm.cu", "hm_analysis.cu"]
HEADERS = [os.path.join(CSRC, "hm_common.cuh"), os.path.join(CSRC, "hm_sim_common.cuh"),
           os.path.join(CSRC, "hm_mg_onchip.cuh"), os.path.join(CSRC, "hm_ptx.cuh"),
           os.path.join(HERE, "..", "include", "hm_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhm_b200.so cannot be built")


STAMP = LIB + ".stamp"


def source_hash() -> str:
    """Hash of every source and header the library is built from (mtimes do not survive a repo snapshot)."""
    import hashlib

    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read())
    h.update(os.environ.get("HM_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def needs_build() -> bool:
    """True when the library is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    import fcntl

    with open(LIB + ".lock", "w") as lock:  # one builder at a time (several ranks may start together)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc_path())), "lib64")
    tmp = LIB + ".tmp"
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo", *os.environ.get("HM_NVCC_EXTRA", "").split(),
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
        "-shared",
        "-o", tmp,
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
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
