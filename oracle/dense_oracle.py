"""CPU ORACLE for the dense least-squares configuration (BASELINE config 5) -- test infrastructure only.

Builds ``oracle/dense_ls.c`` (the reference's own ``lib/ipmPD_CSsolver.c`` around hand-written dense group
functions, see that file) for one problem size into ``oracle/_ref/dense_ls_<N>_<n>_*.so`` and runs batches
through the same ``oracle_run_batch`` entry point as the generated oracles (``oracle/oracle.py``).
``kind="ref"`` #includes the reference driver where it lies under /root/reference (build container only; the
.so travels to the GPU box), ``kind="port"`` uses ``oracle/ipm_driver_port.c`` (buildable anywhere).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


class DenseOracle:
    def __init__(self, N: int, n: int, lo: float, hi: float, kind: str = "ref", opt: str = "-O1"):
        self.N, self.n, self.lo, self.hi, self.kind, self.opt = int(N), int(n), float(lo), float(hi), kind, opt
        os.makedirs(OUT, exist_ok=True)
        tag = f"dense_ls_{N}_{n}_{lo!r}_{hi!r}".replace("-", "m").replace(".", "p")
        self.so = os.path.join(OUT, f"{tag}_{kind}{opt.replace('-', '_')}.so")
        self.lib = None

    def build(self, force: bool = False) -> str:
        if os.path.exists(self.so) and not force:
            return self.so
        if self.kind == "ref" and not os.path.isdir("/root/reference"):
            raise RuntimeError("the reference driver is only available in the build container; use kind='port'")
        cmd = ["gcc"] + self.opt.split() + ["-w", "-shared", "-fPIC", "-fvisibility=hidden", f"-DDN={self.N}",
                                             f"-DDn={self.n}", f"-DLO=({self.lo!r})", f"-DHI=({self.hi!r})"]
        if self.kind == "port":
            cmd.append("-DDRIVER_PORT")
        subprocess.check_call(cmd + [os.path.join(HERE, "dense_ls.c"), "-o", self.so, "-lm"])
        return self.so

    def load(self):
        if self.lib is None:
            self.build()
            lib = ctypes.CDLL(self.so)
            dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
            lib.oracle_run_batch.restype = None
            lib.oracle_run_batch.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int32, ctypes.c_int32,
                                             dp, ip, ip, dp, dp]
            self.lib = lib
        return self.lib

    def solve_batch(self, A: np.ndarray, b: np.ndarray, x0: np.ndarray, mu0: float, maxIter: int,
                    addEye2Hessian=(1e-9, 1e-9)):
        """A [N,n,B], b [N,B], x0 [n,B] (batch last, like the generated class).  Returns status[B], iter[B],
        J[B], x[n,B], seconds[B]."""
        lib = self.load()
        B = A.shape[-1]
        P = np.ascontiguousarray(np.concatenate([A.reshape(self.N * self.n, B, order="F"), b], axis=0).T)
        V = np.ascontiguousarray(x0.T)
        Y = np.zeros((B, 1 + self.n))
        st = np.zeros(B, np.int32)
        it = np.zeros(B, np.int32)
        sec = np.zeros(B)
        ae = np.asarray(addEye2Hessian, dtype=np.float64)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        lib.oracle_run_batch(B, P.ctypes.data_as(dp), V.ctypes.data_as(dp), mu0, maxIter, -1, ae.ctypes.data_as(dp),
                             st.ctypes.data_as(ip), it.ctypes.data_as(ip), Y.ctypes.data_as(dp), sec.ctypes.data_as(dp))
        return st, it, Y[:, 0].copy(), np.ascontiguousarray(Y[:, 1:].T), sec
