"""Dense least squares with box constraints (BASELINE config 5) -- host-side mirror of the generated class.

``examples/sls.m:136-166`` generates, with ``cmex2optimizeCS``, a class with ``setP_A, setP_b, setV_x, solve,
getOutputs`` for ``min ||A x - b||^2/N  s.t.  lo <= x <= hi``.  For N=4000, n=256 the reference cannot emit the
scalar code; :class:`DenseLSSolver` exposes the same methods on a BATCH of instances and runs
the dense kernels of ``csrc/tc_dense.cu`` through the C ABI (``tc_dense_*`` in ``include/tenscalc_b200.h``).
Arrays carry the batch as trailing dimension like :class:`tenscalc_b200.runtime.BatchedSolver`.  There is no CPU
path: construction fails without a CUDA device.

``dense_trailing_block`` is the structural test that routes a problem here: the size of the trailing block of
the symbolic LDL' factor (``ldl.py``) whose columns are completely dense.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import numpy as np

from .runtime import TenscalcError, _check, load_library

# algorithmic flop counts of the path, as the reference's emitter would count the scalar code
# (instructionsTableUTHash.c:816-832: one multiply-add = 2)
def gram_flops(N: int, n: int) -> int:
    """H0=(2/N)A'A (lower triangle), g0=(2/N)A'b, b'b: once per set of (A, b)"""
    return 2 * N * (n * (n + 1) // 2) + 2 * N * n + 2 * N


def newton_step_flops(n: int) -> int:
    """one Newton step: dense LDL' of the n x n Schur block (n^3/3), two triangular solves (2 n^2), three
    matrix-vector products with H0 (gradient, curvature / direction error: 2 n^2 each) and O(n) vector work"""
    return n ** 3 // 3 + 2 * n * n + 3 * 2 * n * n + 40 * n


from .ldl import dense_trailing_block  # noqa: E402,F401  (the structural test that routes a problem here)


class DenseLSSolver:
    def __init__(self, N: int, n: int, batch: int, lo: float, hi: float, device: int = 0):
        lib = load_library()
        if lib.tc_device_count() == 0:
            raise TenscalcError("no CUDA device visible: the batched solver has no CPU path")
        self.N, self.n, self.batch, self.lo, self.hi = int(N), int(n), int(batch), float(lo), float(hi)
        h = ctypes.c_void_p()
        _check(lib.tc_dense_create(ctypes.byref(h), self.N, self.n, self.batch, self.lo, self.hi, int(device)))
        self._h, self._lib = h, lib
        self._keep = {}

    # -- reference surface: arrays [..., batch] on the host -------------------------------------------
    def _host(self, value, shape: Tuple[int, ...]) -> np.ndarray:
        a = np.asarray(value, dtype=np.float64)
        if a.shape == shape:
            a = np.broadcast_to(a[..., None], shape + (self.batch,))
        if a.shape != shape + (self.batch,):
            raise TenscalcError(f"expected shape {shape} or {shape + (self.batch,)}, got {a.shape}")
        return np.ascontiguousarray(a.reshape((-1, self.batch), order="F").T)      # [batch][column-major entries]

    def setP_A(self, value) -> None:
        flat = self._host(value, (self.N, self.n))
        _check(self._lib.tc_dense_set_A(self._h, flat.ctypes.data_as(ctypes.c_void_p), 0))
        self._lib.tc_dense_synchronize(self._h)        # the staging copy has left `flat`

    def setP_b(self, value) -> None:
        flat = self._host(value, (self.N,))
        _check(self._lib.tc_dense_set_b(self._h, flat.ctypes.data_as(ctypes.c_void_p), 0))
        self._lib.tc_dense_synchronize(self._h)

    def setV_x(self, value) -> None:
        flat = self._host(value, (self.n,))
        _check(self._lib.tc_dense_set_x(self._h, flat.ctypes.data_as(ctypes.c_void_p), 0))
        self._lib.tc_dense_synchronize(self._h)

    def solve(self, mu0: float = 1.0, maxIter: int = 200, saveIter: int = -1,
              addEye2Hessian: Sequence[float] = (1e-9, 1e-9)):
        """returns (status[batch] int32, iter[batch] int32, device time of the batch in seconds)"""
        mu, mi, si = ctypes.c_double(mu0), ctypes.c_int32(maxIter), ctypes.c_int32(saveIter)
        ae = (ctypes.c_double * 2)(*addEye2Hessian)
        st = np.empty(self.batch, np.int32)
        it = np.empty(self.batch, np.int32)
        t = ctypes.c_double(0)
        _check(self._lib.tc_dense_solve(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae,
                                        st.ctypes.data_as(ctypes.c_void_p), it.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.byref(t)))
        return st, it, t.value

    def getOutputs(self):
        """[J (batch,), x (n, batch)] -- the outputExpressions {J, x} of examples/sls.m"""
        J = np.empty(self.batch)
        x = np.empty((self.batch, self.n))
        _check(self._lib.tc_dense_get_outputs(self._h, J.ctypes.data_as(ctypes.c_void_p),
                                              x.ctypes.data_as(ctypes.c_void_p), None, None, 0))
        return [J, np.ascontiguousarray(x.T)]

    # -- zero-copy variants (device pointers; A and b are borrowed until the next solve is enqueued) ----
    def set_device_ptrs(self, A_ptr: int = 0, b_ptr: int = 0, x_ptr: int = 0) -> None:
        if A_ptr:
            _check(self._lib.tc_dense_set_A(self._h, ctypes.c_void_p(A_ptr), 1))
        if b_ptr:
            _check(self._lib.tc_dense_set_b(self._h, ctypes.c_void_p(b_ptr), 1))
        if x_ptr:
            _check(self._lib.tc_dense_set_x(self._h, ctypes.c_void_p(x_ptr), 1))

    def set_host_ptrs(self, A: np.ndarray = None, b: np.ndarray = None, x: np.ndarray = None) -> None:
        """C-contiguous [batch][...] host arrays (e.g. pinned); copies are enqueued on the handle's stream"""
        for arr, fn in ((A, self._lib.tc_dense_set_A), (b, self._lib.tc_dense_set_b), (x, self._lib.tc_dense_set_x)):
            if arr is not None:
                _check(fn(self._h, arr.ctypes.data_as(ctypes.c_void_p), 0))

    def solve_async(self, mu0: float = 1.0, maxIter: int = 200, saveIter: int = -1,
                    addEye2Hessian: Sequence[float] = (1e-9, 1e-9)) -> None:
        mu, mi, si = ctypes.c_double(mu0), ctypes.c_int32(maxIter), ctypes.c_int32(saveIter)
        ae = (ctypes.c_double * 2)(*addEye2Hessian)
        _check(self._lib.tc_dense_solve_async(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae))

    def get_outputs_ptrs(self, J_ptr: int, x_ptr: int, status_ptr: int = 0, iter_ptr: int = 0, device: bool = True) -> None:
        _check(self._lib.tc_dense_get_outputs(self._h, ctypes.c_void_p(J_ptr), ctypes.c_void_p(x_ptr),
                                              ctypes.c_void_p(status_ptr) if status_ptr else None,
                                              ctypes.c_void_p(iter_ptr) if iter_ptr else None, int(device)))

    def times_ms(self) -> Tuple[float, float]:
        """CUDA-event times of the last solve: (Gram kernel, Newton-loop kernel)"""
        a, b = ctypes.c_float(0), ctypes.c_float(0)
        _check(self._lib.tc_dense_times_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def counters(self) -> Tuple[int, int]:
        """(kernel launches, FP64 tensor-core flops issued by the LDL' trailing updates) since creation"""
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _check(self._lib.tc_dense_counters(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def stream_ptr(self) -> int:
        return int(self._lib.tc_dense_stream(self._h) or 0)

    def synchronize(self) -> None:
        _check(self._lib.tc_dense_synchronize(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.tc_dense_destroy(self._h)
            self._h = None

    delete = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synthetic_inputs(N: int, n: int, batch: int, seed: int = 0):
    """A ~ N(0,1)/sqrt(N), b ~ N(0,1), x0 mid-box; with lo=-1, hi=1 about 30 % of the bounds are
    active at the optimum.  Returns A [N,n,B], b [N,B], x0 [n,B]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = rng.standard_normal((N, n, batch)) / np.sqrt(N)
    b = rng.standard_normal((N, batch))
    return A, b, np.zeros((n, batch))
