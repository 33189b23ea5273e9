"""Build libcosmos_b200.so (and the primitive self-test binary) in-tree with nvcc for sm_100a.

    python -m cosmos_b200.build [--force] [-v]

The .so lands next to this file (cosmos_b200/libcosmos_b200.so); it is git-ignored but travels
to the GPU box with the repository snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcosmos_b200.so")
SELFTEST = os.path.join(HERE, "selftest_sm100")
STAMP = os.path.join(HERE, ".build_stamp")

LIB_SOURCES = ["api.cu", "ema.cu", "infonce_fwd.cu", "infonce_bwd.cu", "infonce_bwd_pair.cu", "infonce_bwd_quad.cu", "infonce_bwd_e2.cu", "infonce_bwd_e2t.cu", "infonce_aux.cu", "xpool.cu", "gemm.cu", "retrieval.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-diag-suppress", "177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cosmos_b200 needs the CUDA toolkit to build its sm_100a kernels")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def is_current() -> bool:
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and is_current():
        return LIB
    nvcc = _nvcc()
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [nvcc, *ARCH, *COMMON, "-Xptxas", "-v", "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    st = os.path.join(CSRC, "selftest.cu")
    if os.path.exists(st):
        cmd = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-diag-suppress", "177", "-o", SELFTEST, st]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for cmd, p in procs:
        out, _ = p.communicate()
        log.append(" ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(objdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.run([nvcc, *ARCH, "-shared", "-o", LIB, *objs], check=True)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
