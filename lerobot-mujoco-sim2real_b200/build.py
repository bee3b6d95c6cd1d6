"""Build the CUDA C-ABI library in-tree: lerobot-mujoco-sim2real_b200/libso101_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with gpurun snapshots.
Five translation units, compiled in parallel: the host side (csrc/so101_capi.cu, with the small kernels, the IK and the
Koopman kernels) and csrc/so101_kernels.cu once per (dtype, kernel family) - the stepping kernels are ~10 k SASS
instructions each and dominate the build time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
       [os.path.join(ROOT, "include", "so101_b200.h"), os.path.abspath(__file__)]
LIB = os.path.join(HERE, "libso101_b200.so")
UNITS = [("capi", "so101_capi.cu", [])] + [
    (f"kernels_{'f64' if t == 'double' else 'f32'}_{'team' if s == 'true' else 'onewarp'}", "so101_kernels.cu",
     [f"-DSO101_TU_T={t}", f"-DSO101_TU_SPLIT={s}"]) for t in ("double", "float") for s in ("false", "true")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = os.path.join(OBJ, "extra_flags.txt")   # the -D flags of the build that produced LIB (experiments); absent = none


def up_to_date() -> bool:
    """The library is newer than its sources AND was built without experiment flags (a library built with
    `-DSO101_EXP_...` for a measurement must never pass for the product build)."""
    if not (os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in DEPS)):
        return False
    try:
        with open(STAMP) as f:
            return f.read().strip() == ""
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if not force and up_to_date():
        return LIB
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    os.makedirs(OBJ, exist_ok=True)
    base = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v",
            "-Xcompiler", "-fPIC", "-ccbin", ccbin, *extra]

    def one(unit):
        name, src, defs = unit
        obj = os.path.join(OBJ, name + ".o")
        cmd = base + defs + ["-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return name, obj, " ".join(cmd), res

    with ThreadPoolExecutor(len(UNITS)) as ex:
        done = list(ex.map(one, UNITS))
    log = "".join(f"{cmd}\n{res.stdout}{res.stderr}\n" for _, _, cmd, res in done)
    bad = [name for name, _, _, res in done if res.returncode != 0]
    if not bad:
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-ccbin", ccbin,
               "-o", LIB] + [obj for _, obj, _, _ in done]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        if res.returncode != 0:
            bad = ["link"]
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(log)
    if verbose or bad:
        sys.stderr.write(log if verbose else "\n".join(l for l in log.split("\n") if "error" in l or "rror:" in l)[:8000] + "\n")
    if bad:
        raise RuntimeError(f"nvcc failed ({', '.join(bad)}), see csrc/build.log")
    with open(STAMP, "w") as f:
        f.write(" ".join(extra) + "\n")
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv or bool(extra), verbose="-v" in sys.argv, extra=extra))
