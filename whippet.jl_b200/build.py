"""In-tree build of the CUDA library (nvcc, sm_100a)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
CUDA_SO = os.path.join(CSRC, "libwhippet_b200.so")
_CU_DEPS = ["whippet_b200.cu", "wq_align.cuh", "wq_kernels.cuh", "wq_types.h", "wq_host_index.h",
            "wq_host_quant.h", "wq_em.cuh", "wq_events.h", "wq_fastq.h", "wq_sam.h", "wq_index_file.h", "wq_inflate.h", "wq_bias.cuh"]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def cuda_stale():
    if not os.path.exists(CUDA_SO):
        return True
    deps = [os.path.join(CSRC, f) for f in _CU_DEPS] + [os.path.join(_HERE, "..", "include", "whippet_b200.h")]
    return os.path.getmtime(CUDA_SO) < max(os.path.getmtime(d) for d in deps)


def build_cuda(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> csrc/libwhippet_b200.so"""
    if not force and not cuda_stale():
        return CUDA_SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
           "-Xcompiler", "-fPIC", "-shared", "-o", CUDA_SO, os.path.join(CSRC, "whippet_b200.cu"), "-lz", "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return CUDA_SO


def build_all(force=False):
    return build_cuda(force)
