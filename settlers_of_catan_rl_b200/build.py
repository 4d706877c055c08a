"""Build the CUDA extension in-tree: ``settlers_of_catan_rl_b200/csrc/libcatan_b200.so``.

nvcc cross-compiles for sm_100a without a GPU.  The built ``.so`` is git-ignored but travels to the
GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SO = os.path.join(CSRC, "libcatan_b200.so")
SOURCES = ("catan_kernels.cu", "ppo_kernels.cu", "policy_kernels.cu")
HEADERS = ("catan_core.cuh", "catan_game.cuh", "device_scope.cuh", os.path.join("..", "..", "include", "catan_b200.h"),
           os.path.join("..", "..", "include", "catan_layout.h"), os.path.join("..", "..", "include", "catan_topology.h"))
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU path)")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


PROF_SO = os.path.join(CSRC, "libcatan_b200_prof.so")


def build_profiling_extension() -> str:
    """instrumented twin (-DCATAN_PROFILE_PHASES): per-phase clock64 timers, used only by profiles/phase_profile.py"""
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-DCATAN_PROFILE_PHASES", "-o", PROF_SO] + list(SOURCES)
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return PROF_SO


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + list(SOURCES)
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return SO


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
