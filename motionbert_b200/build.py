"""In-tree build of libmotionbert_b200.so (sm_100a only) with nvcc, and of its test twin.

    python -m motionbert_b200.build [--force] [--verbose]

libmotionbert_b200.so      the product: the C ABI of include/motionbert_b200.h, production kernels only
libmotionbert_b200_test.so the same source with -DMB_TEST_KERNELS: additionally the CUDA-core / first-generation
                           reference kernels and the kernel-level hooks of include/motionbert_b200_test.h (tests only)

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  There is exactly one code path: compute_100a / sm_100a.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmotionbert_b200.so")
TEST_LIB = os.path.join(HERE, "libmotionbert_b200_test.so")
SOURCES = ["mb_api.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "gemm_tc2.cuh", "mlp_fused.cuh", "attn_t_tc.cuh", "attn_s_tc.cuh", "attn_s_f16c.cuh", "attn_t_f16c.cuh", "attn_bwd_tc.cuh", "backward_kernels.cuh", "loss_kernels.cuh", "optim_kernels.cuh", "simt_kernels.cuh", "wgrad_tc.cuh",
           os.path.join("..", "..", "include", "motionbert_b200.h"), os.path.join("..", "..", "include", "motionbert_b200_test.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(lib: str, defines, verbose: bool):
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-DMB_BUILD"] + defines + ["-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd


def build(force: bool = False, verbose: bool = False) -> str:
    """Build both libraries (in parallel) if their sources are newer; returns the product library's path."""
    jobs = []
    for lib, defs in ((LIB, []), (TEST_LIB, ["-DMB_TEST_KERNELS"])):
        if force or needs_build(lib):
            jobs.append((lib,) + _compile(lib, defs, verbose))
    for lib, proc, cmd in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose:
            print(out)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
