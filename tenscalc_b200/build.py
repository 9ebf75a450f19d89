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


def build_colamd(force: bool = False) -> str:
    """``tenscalc_b200/_gen/libcolamd.so``: COLAMD (``symamd_l``) compiled, where they lie, from the
    sources of a SuiteSparse tree -- ``TENSCALC_SUITESPARSE`` or the reference's vendored submodule --
    the third-party library behind MATLAB's ``symamd`` that ``sparsity_ldl.m:91`` calls.  Returns ""
    when no SuiteSparse tree is present (``ordering.symamd`` then uses its committed cache)."""
    from .ordering import COLAMD_LIB
    root = os.environ.get("TENSCALC_SUITESPARSE", "/root/reference/SuiteSparse")
    src = os.path.join(root, "COLAMD", "Source", "colamd.c")
    if not os.path.exists(src):
        return COLAMD_LIB if os.path.exists(COLAMD_LIB) else ""
    if force or not os.path.exists(COLAMD_LIB):
        os.makedirs(os.path.dirname(COLAMD_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-w", "-shared", "-fPIC", "-DDLONG",
                               "-I" + os.path.join(root, "COLAMD", "Include"),
                               "-I" + os.path.join(root, "SuiteSparse_config"), src,
                               os.path.join(root, "SuiteSparse_config", "SuiteSparse_config.c"),
                               "-o", COLAMD_LIB, "-lm"])
    return COLAMD_LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))


# --------------------------------------------------------------------------------------------
# per-problem compiled modules (the counterpart of the dynamic library cmex2optimizeCS builds
# with the C compiler at generation time, lib/cmex2optimizeCS.m:403-433)
# --------------------------------------------------------------------------------------------
GEN_DIR = os.path.join(HERE, "_gen")                                  # the modules (.so): travel to the GPU box
GEN_SRC = os.path.join(os.path.dirname(HERE), "build", "gen_src")    # emitted CUDA + objects: stay here
MODULE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
                "-rdc=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", CSRC]


def _toolchain_tag(emit_opts: dict) -> str:
    """hash of everything besides the instruction list that shapes the module: the device headers,
    the emitter, the compiler flags and the chunking options (content-based: mtimes do not survive
    the copy to the GPU box)"""
    import hashlib
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "emit_cuda.py"), "rb").read())
    h.update(repr((MODULE_FLAGS[:-1], sorted(emit_opts.items()))).encode())   # includes warps_per_sm -> -maxrregcount
    return h.hexdigest()[:8]


def default_module_opts(gen, **given) -> dict:
    """launch shape per problem size (measured on B200, profiles/): small scratchbooks (mpc_dcmotor,
    sls) run fastest with many light warps -- 24 warps/SM at 80 registers, 28 operands per chunk, the
    next chunk's operands prefetched into L2.  Large ones (mpc_quadcopter, l1l2) run 12 warps/SM at 168
    registers with 72 operands per chunk, the warps grouped in CTAs of 4 that walk the straight-line
    code in lockstep (one barrier per chunk), so that the instruction stream -- instruction fetch is a
    fifth of these kernels' time, half of the lines that miss the SM come from DRAM -- is fetched once
    per CTA: +17 % over 8 independent warps/SM at 255 registers (profiles/r02i_*); and WITHOUT the L2
    prefetch: the bulk phases are DRAM-throughput bound (4.2-4.4 TB/s of 256-byte segments), so the
    prefetch instructions (one per operand, 14 % of all instructions) only cost issue slots -- +10 % on
    mpc_quadcopter and l1l2 without them (profiles/r03_experiments.md).  Lockstep needs the
    warp-synchronous driver with one instance per lane."""
    solo = given.get("driver", "warp") != "warp" or given.get("tiles", 1) != 1
    if gen.problem.n_slots <= 8192:        # lockstep 8: mpc_dcmotor +13 %, sls +7 % (profiles/r02j_*)
        return dict(warps_per_sm=24, max_loads=28, **({} if solo else dict(lockstep=8)))
    return dict(prefetch=0) if solo else dict(lockstep=4, warps_per_sm=12, max_loads=72, prefetch=0)


def module_path(gen, **emit_opts) -> str:
    from .csparse import fingerprint
    return os.path.join(GEN_DIR, f"{gen.classname}_{fingerprint(gen.problem)[:10]}_{_toolchain_tag(emit_opts)}.so")


def build_module(gen, force: bool = False, verbose: bool = False, jobs: int = 0, **emit_opts) -> str:
    """Emit the CUDA module of a generated solver and compile it for sm_100a (in-tree, so the
    .so travels to the GPU box).  Returns the path of the shared object."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    from .emit_cuda import emit_module
    emit_opts = {**default_module_opts(gen, **emit_opts), **emit_opts}
    so = module_path(gen, **emit_opts)
    if not force and os.path.exists(so):
        return so
    # fmad=1 (experiment, profiles/r03u_*): let nvcc contract a*b+c into DFMA.  NOT the arithmetic contract of the
    # product -- the reference's x86-64 build rounds products and sums separately, and so does the default module
    fmad = int(emit_opts.get("fmad", 0))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        raise RuntimeError(f"compiled module {so} is missing and nvcc is not available to build it")
    t0 = time.time()
    files, stats = emit_module(gen, **emit_opts)
    src_dir = os.path.join(GEN_SRC, os.path.basename(so)[:-3])
    cache_dir = os.path.join(GEN_SRC, "objcache")     # part files do not depend on the launch shape
    os.makedirs(src_dir, exist_ok=True)
    os.makedirs(cache_dir, exist_ok=True)
    os.makedirs(GEN_DIR, exist_ok=True)
    for f in os.listdir(src_dir):
        os.remove(os.path.join(src_dir, f))
    for name, src in files.items():
        with open(os.path.join(src_dir, name), "w") as f:
            f.write(src)
    import hashlib
    import shutil

    # register budget of the chunk functions: the kernel runs warps_per_sm warps on the 64K-register SM
    # (registers are allocated to warps in groups of four: 14 warps/SM cost what 16 do)
    reg_budget = min(255, 65536 // (32 * ((int(emit_opts.get("warps_per_sm", 8)) + 3) // 4 * 4)) // 8 * 8)
    flags = [("-fmad=true" if fmad and f == "-fmad=false" else f) for f in MODULE_FLAGS] + [f"-maxrregcount={reg_budget}"]

    def compile_one(name: str) -> str:
        obj = os.path.join(src_dir, name[:-3] + ".o")
        key = hashlib.sha1((files[name] + repr(flags[:-3]) + flags[-1]).encode()).hexdigest()
        cached = os.path.join(cache_dir, key + ".o")
        if name != "main.cu" and os.path.exists(cached):
            shutil.copyfile(cached, obj)
            return ""
        res = subprocess.run([nvcc] + flags + ["-Xptxas", "-v", "-dc", os.path.join(src_dir, name), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(f"nvcc failed on {name}:\n{res.stdout}{res.stderr}")
        if name != "main.cu":
            shutil.copyfile(obj, cached)
        return res.stderr

    names = sorted(files, key=lambda n: -len(files[n]))
    with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as pool:
        logs = list(pool.map(compile_one, names))
    objs = [os.path.join(src_dir, n[:-3] + ".o") for n in names]
    res = subprocess.run([nvcc] + flags + ["-shared"] + objs + ["-o", so], capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError(f"nvcc link failed:\n{res.stdout}{res.stderr}")
    import re
    text = "\n".join(logs)
    regs = [int(x) for x in re.findall(r"Used (\d+) registers", text)]
    spills = [int(a) for a in re.findall(r"(\d+) bytes spill stores", text)]
    with open(os.path.join(src_dir, "build.log"), "w") as f:
        f.write(f"emit+compile {time.time() - t0:.1f}s files {len(files)} functions {stats['n_functions']} "
                f"max_regs {max(regs) if regs else 0} spill_store_bytes {sum(spills)} traffic {stats['total']}\n")
        f.write(text)
    if verbose:
        print(f"module {os.path.basename(so)}: {time.time() - t0:.1f}s, {len(files)} files, "
              f"{stats['n_functions']} functions, max regs {max(regs) if regs else 0}, "
              f"spill stores {sum(spills)} B, {stats['total']}")
    return so
