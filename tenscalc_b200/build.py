"""Build of the native library (nvcc, sm_100a only, in-tree so that the .so travels)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtenscalc_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tenscalc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + sources() + ["-o", LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode:
            print(res.stdout + res.stderr)
        if res.returncode:
            raise RuntimeError("nvcc failed")
        with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
