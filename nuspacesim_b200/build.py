"""Build libnss_b200.so in-tree with nvcc for sm_100a.

    python -m nuspacesim_b200.build [--force]

Two translation units: the float64 stage kernels are compiled with -fmad=false so their
arithmetic rounds like the reference's NumPy float64 expressions; the float32 EAS kernel
keeps FMA contraction (its two rounding-sensitive spots use explicit IEEE intrinsics).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("NSS_B200_LIB") or os.path.join(HERE, "libnss_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
UNITS = [
    ("nss_eas_kernel.cu", []),
    ("nss_f64_kernels.cu", ["-fmad=false"]),
    ("nss_microbench.cu", []),
]
DEPS = ["nss_eas_core.cuh", "nss_eas_tables.h", os.path.join("..", "..", "include", "nss_b200.h")]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnss_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [u for u, _ in UNITS] + DEPS
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs = []
    log = []
    for src, extra in UNITS:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
