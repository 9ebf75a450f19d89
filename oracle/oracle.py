"""CPU ORACLE harness (test infrastructure only -- never imported by the product path).

Builds the per-problem oracle library and runs batches through it, one instance after the
other, exactly as a MATLAB user would drive the generated class:
``set parameters -> set primal initial values -> ipmPD_CSsolver -> getOutputs``.

kinds
-----
``ref``   reference emitter + reference driver (needs /root/reference at BUILD time only;
          the resulting ``oracle/_ref/<name>_ref.so`` is what travels to the GPU box)
``port``  this repo's restatement of both (``emit_c.statement`` + ``ipm_driver_port.c``);
          pinned against ``ref`` by tests/test_oracle.py
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

from tenscalc_b200.optimize import GeneratedSolver

from . import emit_c
from .build_ref import build_itable

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def pack_inputs(gen: GeneratedSolver, P: Dict[str, np.ndarray], V: Dict[str, np.ndarray], batch: int):
    """[batch, nP] and [batch, nV] C-contiguous blocks (all set functions concatenated)"""
    from tenscalc_b200.runtime import pack_inputs as pack_sets
    Ps, Vs = pack_sets(gen, P, V, batch)
    cat = lambda d: np.ascontiguousarray(np.concatenate(list(d.values()), axis=1)) if d else np.zeros((batch, 0))
    return cat(Ps), cat(Vs)


def unpack_outputs(gen: GeneratedSolver, Y: np.ndarray):
    outs, off = {}, 0
    B = Y.shape[0]
    for name, shape in gen.outputs:
        n = int(np.prod(shape)) if shape else 1
        outs[name] = Y[:, off:off + n].T.reshape(tuple(shape) + (B,), order="F")
        off += n
    return outs


class Oracle:
    def __init__(self, gen: GeneratedSolver, kind: str = "ref", opt: str = "-O1", tag: Optional[str] = None,
                 variant: str = ""):
        self.gen = gen
        self.kind = kind
        os.makedirs(OUT, exist_ok=True)
        if tag is None:
            # the library is tied to the exact instruction list (not to the device tape format)
            from tenscalc_b200.csparse import fingerprint
            tag = gen.classname + "_" + fingerprint(gen.problem)[:10]
        self.opt = opt
        if variant not in ("", "crpow", "count"):
            raise ValueError("variant must be '', 'crpow' (libm-sensitivity probe) or 'count' (group execution counters)")
        self.variant = variant
        self.so = os.path.join(OUT, f"{tag}_{kind}{opt.replace('-', '_').replace(' ', '').replace('=', '')}"
                                    f"{'_' + variant if variant else ''}.so")
        self.lib = None

    def build(self, force: bool = False) -> str:
        if os.path.exists(self.so) and not force:
            return self.so
        if self.kind == "ref":
            build_itable()
        src = emit_c.emit_translation_unit(self.gen, emitter=self.kind, driver=self.kind)
        cfile = self.so[:-3] + ".c"
        with open(cfile, "w") as f:
            f.write(src)
        # plain gcc, no -march/-ffast-math: the reference is built the same way by CmexTools
        extra = ["-DORACLE_CRPOW", "-mfma", "-ffp-contract=off"] if self.variant == "crpow" else []
        if self.variant == "count":
            extra = ["-DORACLE_COUNT"]
        subprocess.check_call(["gcc"] + self.opt.split() + extra + ["-w", "-shared", "-fPIC", "-fvisibility=hidden",
                               cfile, "-o", self.so, "-lm"])
        return self.so

    def load(self):
        if self.lib is None:
            self.build()
            lib = ctypes.CDLL(self.so)
            lib.oracle_run_batch.restype = None
            dp = ctypes.POINTER(ctypes.c_double)
            ip = ctypes.POINTER(ctypes.c_int32)
            lib.oracle_run_batch.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int32,
                                             ctypes.c_int32, dp, ip, ip, dp, dp]
            self.lib = lib
        return self.lib

    def solve_batch(self, P, V, mu0: float, maxIter: int, saveIter: int = -1,
                    addEye2Hessian=(1e-9, 1e-9)):
        """returns status[B], iter[B], outputs{name: array(...,B)}, seconds[B]"""
        lib = self.load()
        first = next(iter(list(P.values()) + list(V.values())))
        B = np.asarray(first).shape[-1]
        Pp, Vp = pack_inputs(self.gen, P, V, B)
        nY = sum(int(np.prod(s)) if s else 1 for _, s in self.gen.outputs)
        Y = np.zeros((B, nY))
        status = np.zeros(B, dtype=np.int32)
        it = np.zeros(B, dtype=np.int32)
        sec = np.zeros(B)
        ae = list(addEye2Hessian)
        ae = np.asarray(ae + [ae[-1]] * (3 - len(ae)), dtype=np.float64)      # {U, Eq} or, min-max, {U, D, Eq}
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.oracle_run_batch(B, Pp.ctypes.data_as(dp), Vp.ctypes.data_as(dp), mu0, maxIter, saveIter,
                             ae.ctypes.data_as(dp), status.ctypes.data_as(ip), it.ctypes.data_as(ip),
                             Y.ctypes.data_as(dp), sec.ctypes.data_as(dp))
        return status, it, unpack_outputs(self.gen, Y), sec

    def group_counts(self, reset: bool = True) -> np.ndarray:
        """variant 'count': how often every dependency group ran since the last reset"""
        lib = self.load()
        out = np.zeros(max(len(self.gen.problem.groups), 1), dtype=np.int64)
        lib.oracle_group_counts(out.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), ctypes.c_int(int(reset)))
        return out[:len(self.gen.problem.groups)]
