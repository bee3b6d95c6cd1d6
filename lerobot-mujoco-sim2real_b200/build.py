"""Build the CUDA C-ABI library in-tree: lerobot-mujoco-sim2real_b200/libso101_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "so101_capi.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("so101_physics.cuh", "so101_model.h", "so101_koopman.cuh",
                                                        "so101_ik.cuh", "so101_contact.cuh")] + [os.path.join(ROOT, "include", "so101_b200.h")]
LIB = os.path.join(HERE, "libso101_b200.so")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC", "-ccbin", ccbin, "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "csrc", "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see csrc/build.log")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
