"""Builds mofo_b200/libmofo_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mofo_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmofo_sm100.so")
SOURCES = ["runtime.cu", "simple_kernels.cu", "gemm.cu", "attention.cu", "attention_small.cu", "optimizer.cu", "motion_kernels.cu"]
HEADERS = ["common.cuh", "attn_helpers.cuh", os.path.join("..", "..", "include", "mofo_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# --use_fast_math (flush-to-zero, approximate division / sqrt) is for the GEMM / attention epilogues and the streaming
# kernels whose results are rounded to bf16 anyway; the optimizer keeps IEEE division / sqrt and denormals so that
# adamw_kernel really is torch.optim.AdamW's arithmetic (exp_avg_sq of tiny gradients must not flush to zero)
FAST_MATH = {"runtime.cu": True, "simple_kernels.cu": True, "gemm.cu": True, "attention.cu": True, "attention_small.cu": True, "optimizer.cu": False,
             "motion_kernels.cu": False}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["--use_fast_math"] if FAST_MATH.get(s, True) else []) + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- nvcc {s} ---\n{out}\n")
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
