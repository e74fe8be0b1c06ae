"""In-tree build of libhrn_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python highres-net_b200/build.py [--force]

The .so lands next to the sources (highres-net_b200/csrc/libhrn_b200.so); it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libhrn_b200.so")
SOURCES = ["api.cu", "conv3x3_umma.cu", "fuse_wave_umma.cu", "enc_wave_umma.cu", "resblock64_umma.cu", "conv_init_umma.cu", "decoder_umma.cu", "pointwise.cu", "scoring.cu", "lanczos7_tma.cu", "shiftnet.cu", "imageset_io.cu"]
HEADERS = ["internal.h", "ptx.cuh", "umma_common.cuh", "strips.cuh", os.path.join("..", "..", "include", "hrn_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built (no CPU fallback exists)")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))] + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log.append(res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart", "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
