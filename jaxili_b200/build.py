"""In-tree build of the native library: nvcc -> jaxili_b200/csrc/libmafb200.so (sm_100a only)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libmafb200.so")
SOURCES = ["mafb200.cu"]
HEADERS = ["plan.h", "common.cuh", "flow_kernel.cuh", "chain_kernel.cuh", "opt_kernels.cuh", "sample_kernel.cuh", "dp_kernel.cuh", "hmc_kernels.cuh", "wide_kernels.cuh", "wide_tc.cuh", "pipeline.inc",
           os.path.join("..", "..", "include", "mafb200.h")]


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", OUT] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("MAFB_TRACE"):
        cmd.insert(1, "-DMAFB_TRACE")
    for exp in os.environ.get("MAFB_EXP", "").split(","):
        if exp:
            cmd.insert(1, "-DMAFB_EXP_" + exp)
    if os.environ.get("MAFB_CTA_TIMES"):
        cmd.insert(1, "-DMAFB_CTA_TIMES")
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
