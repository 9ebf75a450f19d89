"""CPU ORACLE (test infrastructure / benchmark baseline only -- never on the product path).

Runs a prebuilt per-problem oracle library (``oracle/_ref/*.so``: the reference's own
``ipmPD_CSsolver.c`` around code printed by the reference's own ``writeCinstructionsC``) on
all host cores.  The generated C keeps its scratchbook in process-global storage
(``writeCfunctionpergroup.m:148,264-270``: not re-entrant), so parallelism is one PROCESS per
core, each solving a contiguous slice of the batch sequentially -- "one instance per core".
"""
from __future__ import annotations

import ctypes
import multiprocessing as mp
import os
import time
from typing import Tuple

import numpy as np


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _run_slice(args):
    so, Pp, Vp, nY, mu0, maxIter, addEye = args
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.oracle_run_batch.restype = None
    lib.oracle_run_batch.argtypes = [ctypes.c_int64, dp, dp, ctypes.c_double, ctypes.c_int32, ctypes.c_int32,
                                     dp, ip, ip, dp, dp]
    B = Pp.shape[0]
    Y = np.zeros((B, nY))
    status = np.zeros(B, dtype=np.int32)
    it = np.zeros(B, dtype=np.int32)
    sec = np.zeros(B)
    ae = list(addEye)
    ae = np.asarray(ae + [ae[-1]] * (3 - len(ae)), dtype=np.float64)     # the generated runners read three entries
    Pp = np.ascontiguousarray(Pp)
    Vp = np.ascontiguousarray(Vp)
    t0 = time.perf_counter()
    lib.oracle_run_batch(B, Pp.ctypes.data_as(dp), Vp.ctypes.data_as(dp), mu0, maxIter, -1,
                         ae.ctypes.data_as(dp), status.ctypes.data_as(ip), it.ctypes.data_as(ip),
                         Y.ctypes.data_as(dp), sec.ctypes.data_as(dp))
    wall = time.perf_counter() - t0
    return status, it, Y, sec, wall


class CpuPool:
    """A pool of worker processes (spawned, so that it is safe next to an initialised CUDA
    context) that each dlopen the oracle library once."""

    def __init__(self, so: str, cores: int | None = None):
        self.so = so
        self.cores = cores or host_cores()
        self.pool = mp.get_context("spawn").Pool(self.cores)

    def solve(self, Pp: np.ndarray, Vp: np.ndarray, nY: int, mu0: float, maxIter: int,
              addEye=(1e-9, 1e-9)) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, float]:
        """returns status, iter, Y, per-instance solver seconds, wall seconds of the whole call
        (set parameters + ipmPD_CSsolver + getOutputs for every instance, all cores busy)"""
        B = Pp.shape[0]
        bounds = np.linspace(0, B, self.cores + 1).astype(int)
        jobs = [(self.so, Pp[a:b], Vp[a:b], nY, mu0, maxIter, addEye)
                for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        t0 = time.perf_counter()
        res = self.pool.map(_run_slice, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        status = np.concatenate([r[0] for r in res])
        it = np.concatenate([r[1] for r in res])
        Y = np.concatenate([r[2] for r in res])
        sec = np.concatenate([r[3] for r in res])
        return status, it, Y, sec, wall

    def close(self):
        self.pool.close()
        self.pool.join()
