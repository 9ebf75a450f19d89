"""Batched counterpart of the class the reference generates
(``obj=classname(); setP_*(obj,v); setV_*(obj,v); [status,iter,time]=solve(obj,mu0,maxIter,
saveIter,addEye2Hessian); y=getOutputs(obj)`` -- lib/cmex2optimizeCS.m:280-294,
doc/sphinx/optimization.rst:487-556), bound to the C ABI of ``include/tenscalc_b200.h``
through ctypes.  Every array carries the batch as trailing dimension.

There is no CPU fallback: constructing a solver without the native library or without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import tape
from .build import LIB
from .optimize import GeneratedSolver

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

# every symbol include/tenscalc_b200.h declares: (name, restype, argtypes)
ABI = [
    ("tc_version", ctypes.c_int, []),
    ("tc_device_count", ctypes.c_int, []),
    ("tc_create", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_int64, ctypes.c_int]),
    ("tc_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_attach_module", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]),
    ("tc_has_module", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_batch", ctypes.c_int64, [ctypes.c_void_p]),
    ("tc_num_sets", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_num_outputs", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_set_size", ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    ("tc_output_size", ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    ("tc_scratch_bytes", ctypes.c_int64, [ctypes.c_void_p]),
    ("tc_find_set", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    ("tc_find_output", ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    ("tc_set_name", ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int]),
    ("tc_output_name", ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int]),
    ("tc_set", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]),
    ("tc_set_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]),
    ("tc_solve", ctypes.c_int, [ctypes.c_void_p, _dp, _ip, _ip, _dp, ctypes.c_void_p, ctypes.c_void_p, _dp]),
    ("tc_solve_async", ctypes.c_int, [ctypes.c_void_p, _dp, _ip, _ip, _dp, ctypes.c_void_p, ctypes.c_void_p]),
    ("tc_solve_device", ctypes.c_int, [ctypes.c_void_p, _dp, _ip, _ip, _dp, ctypes.c_void_p, ctypes.c_void_p]),
    ("tc_set_phased", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    ("tc_set_phases", ctypes.c_int, [ctypes.c_void_p, _ip, ctypes.c_int]),
    ("tc_get_outputs", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    ("tc_get_outputs_device", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    ("tc_get_outputs_async", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    ("tc_feedback", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _ip, _ip]),
    ("tc_synchronize", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_stream", ctypes.c_void_p, [ctypes.c_void_p]),
    ("tc_last_solve_ms", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    ("tc_kernel_launches", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    ("tc_total_newton_steps", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    ("tc_num_groups", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_group_counts", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    ("tc_group_warp_counts", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    ("tc_error_string", ctypes.c_char_p, [ctypes.c_int]),
    # dense path (tenscalc_b200/dense.py)
    ("tc_dense_create", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    ("tc_dense_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_dense_set_A", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    ("tc_dense_set_b", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    ("tc_dense_set_x", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    ("tc_dense_solve", ctypes.c_int, [ctypes.c_void_p, _dp, _ip, _ip, _dp, ctypes.c_void_p, ctypes.c_void_p, _dp]),
    ("tc_dense_solve_async", ctypes.c_int, [ctypes.c_void_p, _dp, _ip, _ip, _dp]),
    ("tc_dense_get_outputs", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int]),
    ("tc_dense_synchronize", ctypes.c_int, [ctypes.c_void_p]),
    ("tc_dense_stream", ctypes.c_void_p, [ctypes.c_void_p]),
    ("tc_dense_times_ms", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    ("tc_dense_counters", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
]


def load_library():
    """dlopen the in-tree native library and bind the ABI; fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"native library {LIB} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(LIB)
        for name, res, args in ABI:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class TenscalcError(RuntimeError):
    pass


def pack_inputs(gen: GeneratedSolver, P: Dict[str, np.ndarray], V: Dict[str, np.ndarray], batch: int):
    """Host layout of the C ABI: per set function one C-contiguous ``double[batch][n]`` block.
    Value arrays carry the batch as LAST dimension (or none: broadcast) and are flattened
    column-major per instance (MATLAB layout).  Returns ({name: block}, {name: block})."""
    def pack(spec, vals):
        out = {}
        for name, shape in spec:
            a = np.asarray(vals[name], dtype=np.float64)
            n = int(np.prod(shape)) if shape else 1
            if a.shape != tuple(shape) + (batch,):
                a = np.broadcast_to(a.reshape(tuple(shape) + (1,)) if a.shape == tuple(shape) else a,
                                    tuple(shape) + (batch,))
            out[name] = np.ascontiguousarray(a.reshape((n, batch), order="F").T)
        return out
    return pack(gen.parameters, P), pack(gen.variables, V)


def unpack_outputs(gen: GeneratedSolver, blocks: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """inverse of the output layout: double[batch][n_k] blocks -> arrays shaped (..., batch)"""
    outs = {}
    for (name, shape), b in zip(gen.outputs, blocks):
        outs[name] = b.T.reshape(tuple(shape) + (b.shape[0],), order="F")
    return outs


def _check(rc: int) -> None:
    if rc != 0:
        raise TenscalcError(f"tenscalc_b200: {load_library().tc_error_string(rc).decode()} (code {rc})")


def _regularisers(values):
    """addEye2Hessian of solve(): {U, Eq} (ipmPD_CSsolver.c:139), {1, 2} (ipmPDeq_CSsolver) or, for min-max classes,
    {U, D, Eq} (ipmPDminmax_CSsolver.m:46-60: missing entries repeat the last one)"""
    v = [float(x) for x in values]
    v += [v[-1]] * (3 - len(v))
    return (ctypes.c_double * 3)(*v)


class BatchedSolver:
    def __init__(self, gen: GeneratedSolver, batch: int, device: Optional[int] = None,
                 backend: str = "compiled", **module_opts):
        """``backend="compiled"`` (default) runs the per-problem sm_100a module that
        :func:`tenscalc_b200.build.build_module` emits and compiles (built on first use when it is
        not in ``tenscalc_b200/_gen/``, like cmex2optimizeCS compiles its C at generation time);
        ``backend="tape"`` runs the generic tape interpreter (no per-problem compilation).
        ``module_opts`` are emitter options (``lockstep``, ``max_refs``, ``max_loads``, ``driver``), see
        :func:`tenscalc_b200.emit_cuda.emit_module`."""
        if backend not in ("compiled", "tape"):
            raise ValueError("backend must be 'compiled' or 'tape'")
        lib = load_library()
        if lib.tc_device_count() == 0:
            raise TenscalcError("no CUDA device visible: the batched solver has no CPU path")
        self.gen = gen
        self.batch = int(batch)
        self.device = int(device if device is not None else os.environ.get("LOCAL_RANK", 0))
        self._image = tape.serialize(gen, compiled=(backend == "compiled"))
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(self._image, len(self._image))
        _check(lib.tc_create(ctypes.byref(h), buf, len(self._image), self.batch, self.device))
        self._h = h
        self._lib = lib
        self.backend = backend
        if backend == "compiled":
            from .build import build_module
            from .csparse import fingerprint
            try:
                so = build_module(gen, **module_opts)
                _check(lib.tc_attach_module(h, so.encode(), fingerprint(gen.problem).encode()))
            except Exception:
                lib.tc_destroy(h)
                self._h = None
                raise
        # set functions by name, as the reference's `<class>_set_<name>` symbols: the handle resolves them
        self._set_index: Dict[str, int] = {}
        for kind, items in (("P_", gen.parameters), ("V_", gen.variables)):
            for n, _ in items:
                k = lib.tc_find_set(h, n.encode())
                if k < 0:
                    lib.tc_destroy(h)
                    self._h = None
                    raise TenscalcError(f"the problem image has no set function named {n}")
                self._set_index[kind + n] = k
        self._shapes = {("P_" + n): s for n, s in gen.parameters}
        self._shapes.update({("V_" + n): s for n, s in gen.variables})

    # -- reference surface ---------------------------------------------------------------------
    def __getattr__(self, name: str):
        if name.startswith("setP_") or name.startswith("setV_"):
            key = name[3:]
            if key in self.__dict__.get("_set_index", {}):
                return lambda value: self._set(key, value)
        raise AttributeError(name)

    def _set(self, key: str, value) -> None:
        shape = tuple(self._shapes[key])
        n = int(np.prod(shape)) if shape else 1
        idx = self._set_index[key]
        if hasattr(value, "data_ptr"):                       # torch CUDA tensor [batch, n], zero-copy
            if not value.is_cuda or value.dtype.itemsize != 8 or not value.is_contiguous():
                raise TenscalcError("device values must be contiguous float64 CUDA tensors")
            if value.numel() != n * self.batch:
                raise TenscalcError(f"{key}: expected {n}x{self.batch} values")
            import torch
            torch.cuda.current_stream(value.device).synchronize()
            _check(self._lib.tc_set_device(self._h, idx, ctypes.c_void_p(value.data_ptr()), n))
            return
        a = np.asarray(value, dtype=np.float64)
        if a.shape == shape:                                 # one value for the whole batch
            flat = np.ascontiguousarray(a.reshape(-1, order="F"))
            _check(self._lib.tc_set(self._h, idx, flat.ctypes.data_as(ctypes.c_void_p), 0))
        elif a.shape == shape + (self.batch,):
            flat = np.ascontiguousarray(a.reshape((n, self.batch), order="F").T)
            _check(self._lib.tc_set(self._h, idx, flat.ctypes.data_as(ctypes.c_void_p), n))
        else:
            raise TenscalcError(f"set{key}: expected shape {shape} or {shape + (self.batch,)}, got {a.shape}")

    def set_packed(self, key: str, packed: np.ndarray) -> None:
        """values already laid out as C-contiguous [batch, n] (pinned host memory in bench.py)"""
        idx = self._set_index[key]
        _check(self._lib.tc_set(self._h, idx, packed.ctypes.data_as(ctypes.c_void_p), packed.shape[1]))

    def solve(self, mu0: float = 1.0, maxIter: int = 200, saveIter: int = -1,
              addEye2Hessian: Sequence[float] = (1e-9, 1e-9)):
        """returns (status[batch] int32, iter[batch] int32, time in seconds for the batch)"""
        mu = ctypes.c_double(mu0)
        mi = ctypes.c_int32(maxIter)
        si = ctypes.c_int32(saveIter)
        ae = _regularisers(addEye2Hessian)
        status = np.empty(self.batch, dtype=np.int32)
        it = np.empty(self.batch, dtype=np.int32)
        t = ctypes.c_double(0)
        _check(self._lib.tc_solve(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae,
                                  status.ctypes.data_as(ctypes.c_void_p), it.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.byref(t)))
        return status, it, t.value

    def getOutputs(self) -> Union[List[np.ndarray], Dict[str, np.ndarray]]:
        bufs = [np.empty((self.batch, int(np.prod(s)) if s else 1)) for _, s in self.gen.outputs]
        ptrs = (ctypes.c_void_p * len(bufs))(*[b.ctypes.data_as(ctypes.c_void_p) for b in bufs])
        _check(self._lib.tc_get_outputs(self._h, ptrs))
        outs = [b.T.reshape(tuple(s) + (self.batch,), order="F") for b, (_, s) in zip(bufs, self.gen.outputs)]
        if self.gen.output_is_struct:
            return {n: o for (n, _), o in zip(self.gen.outputs, outs)}
        return outs

    # -- zero-copy variants (device pointers / caller-owned pinned buffers) ------------------------
    def set_index(self, key: str) -> int:
        """index of 'P_<name>' / 'V_<name>' in the C ABI's set table"""
        return self._set_index[key]

    def set_device_ptr(self, key: str, ptr: int, n: int) -> None:
        """values are double[batch][n] in device memory at address `ptr`"""
        _check(self._lib.tc_set_device(self._h, self._set_index[key], ctypes.c_void_p(ptr), n))

    def solve_device_ptr(self, status_ptr: int, iter_ptr: int, mu0: float = 1.0, maxIter: int = 200,
                         saveIter: int = -1, addEye2Hessian: Sequence[float] = (1e-9, 1e-9)) -> None:
        """asynchronous solve; status/iter (int32[batch]) stay in device memory"""
        mu = ctypes.c_double(mu0)
        mi = ctypes.c_int32(maxIter)
        si = ctypes.c_int32(saveIter)
        ae = _regularisers(addEye2Hessian)
        _check(self._lib.tc_solve_device(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae,
                                         ctypes.c_void_p(status_ptr), ctypes.c_void_p(iter_ptr)))

    def solve_into(self, status: np.ndarray, it: np.ndarray, mu0: float = 1.0, maxIter: int = 200,
                   saveIter: int = -1, addEye2Hessian: Sequence[float] = (1e-9, 1e-9)) -> float:
        """`solve` writing status/iter into caller-owned (e.g. pinned) int32 host arrays"""
        mu = ctypes.c_double(mu0)
        mi = ctypes.c_int32(maxIter)
        si = ctypes.c_int32(saveIter)
        ae = _regularisers(addEye2Hessian)
        t = ctypes.c_double(0)
        _check(self._lib.tc_solve(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae,
                                  status.ctypes.data_as(ctypes.c_void_p), it.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.byref(t)))
        return t.value

    def solve_async(self, status: np.ndarray, it: np.ndarray, mu0: float = 1.0, maxIter: int = 200,
                    saveIter: int = -1, addEye2Hessian: Sequence[float] = (1e-9, 1e-9)) -> None:
        """enqueue the solve and the copies of status/iter into (pinned) host arrays; they are valid
        after the next getOutputs / get_outputs_ptrs / synchronize on this handle"""
        mu = ctypes.c_double(mu0)
        mi = ctypes.c_int32(maxIter)
        si = ctypes.c_int32(saveIter)
        ae = _regularisers(addEye2Hessian)
        _check(self._lib.tc_solve_async(self._h, ctypes.byref(mu), ctypes.byref(mi), ctypes.byref(si), ae,
                                        status.ctypes.data_as(ctypes.c_void_p), it.ctypes.data_as(ctypes.c_void_p)))

    def get_outputs_ptrs(self, ptrs: Sequence[int], device: bool, sync: bool = True) -> None:
        """getOutputs into caller-owned buffers, one double[batch][n_k] per output; ``sync=False`` with host
        buffers only enqueues the gathers and the copies (``tc_get_outputs_async``: pinned buffers, complete
        after ``synchronize()`` or an event recorded on the handle's stream)"""
        arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])
        fn = self._lib.tc_get_outputs_device if device else (self._lib.tc_get_outputs if sync else self._lib.tc_get_outputs_async)
        _check(fn(self._h, arr))

    def feedback(self, pairs: Dict[str, str]) -> None:
        """Device-resident warm start (``tc_feedback``): ``pairs`` maps an output name to the optimisation
        variable (or parameter) it initialises for the next solve, e.g. ``{"xWarm": "x", "uWarm": "u"}`` for
        ``setV_x(obj,out.xWarm); setV_u(obj,out.uWarm)`` of examples/mpcmhe/mpc_dcmotor.m:223-224.  Nothing
        crosses the host; outputs given as a list are addressed as ``y1, y2, ...``."""
        names = [n for n, _ in self.gen.outputs]
        oi, si = [], []
        for o, v in pairs.items():
            if o not in names:
                raise TenscalcError(f"feedback: no output named {o!r}")
            key = v if v in self._set_index else ("V_" + v if "V_" + v in self._set_index else "P_" + v)
            if key not in self._set_index:
                raise TenscalcError(f"feedback: no variable or parameter named {v!r}")
            oi.append(names.index(o))
            si.append(self._set_index[key])
        a = (ctypes.c_int32 * max(len(oi), 1))(*oi)
        b = (ctypes.c_int32 * max(len(si), 1))(*si)
        _check(self._lib.tc_feedback(self._h, len(oi), a, b))

    def set_phased(self, on: bool = True, phases: Sequence[int] = ()) -> None:
        """phased solve on/off (csrc/tc_driver.cuh); ``phases``: iteration budget of phase 0,1,... (last repeats)"""
        _check(self._lib.tc_set_phased(self._h, int(bool(on))))
        if on:
            arr = (ctypes.c_int32 * max(len(phases), 1))(*phases)
            _check(self._lib.tc_set_phases(self._h, arr, len(phases)))

    def stream_ptr(self) -> int:
        return int(self._lib.tc_stream(self._h) or 0)

    def group_counts(self, reset: bool = False) -> np.ndarray:
        n = self._lib.tc_num_groups(self._h)
        c = np.zeros(n, dtype=np.int64)
        _check(self._lib.tc_group_counts(self._h, c.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(reset)))
        return c

    # -- measurement hooks -------------------------------------------------------------------------
    def group_warp_counts(self, reset: bool = False) -> np.ndarray:
        """warp-executions per dependency group (32*this / group_counts = lanes busy per execution)"""
        n = self._lib.tc_num_groups(self._h)
        out = np.zeros(n, dtype=np.int64)
        _check(self._lib.tc_group_warp_counts(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(reset)))
        return out

    def last_solve_ms(self) -> float:
        ms = ctypes.c_float(0)
        _check(self._lib.tc_last_solve_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def kernel_launches(self) -> int:
        c = ctypes.c_int64(0)
        _check(self._lib.tc_kernel_launches(self._h, ctypes.byref(c)))
        return c.value

    def total_newton_steps(self) -> int:
        c = ctypes.c_int64(0)
        _check(self._lib.tc_total_newton_steps(self._h, ctypes.byref(c)))
        return c.value

    def has_module(self) -> bool:
        """True when the per-problem compiled module is attached (native SASS for every group)"""
        return bool(self._lib.tc_has_module(self._h))

    def scratch_bytes(self) -> int:
        return int(self._lib.tc_scratch_bytes(self._h))

    def synchronize(self) -> None:
        _check(self._lib.tc_synchronize(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.tc_destroy(self._h)
            self._h = None

    delete = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
