"""Tensor-valued symbolic expressions over the scalar instruction table.

Host-side, offline front end mirroring the user surface of the reference's
``@Tcalculus`` class (``lib/@Tcalculus/Tcalculus.m``; ``Tvariable.m``, ``Tconstant.m``,
``Tzeros.m``, ``Tones.m``, ``Teye.m``, ``norm2.m``, ``ts*.m``) for the operators the
BASELINE problem scripts use.  Unlike the reference, which keeps a tensor-level
expression tree and scalarises it later (``computeScalarInstructions.m``), a tensor
here *is* its scalarisation: an object array whose entries are instruction ids, with
``None`` marking structural zeros.

Index conventions are Python's (0-based, half-open slices); storage order for
``reshape``, for packing variables and for all host I/O is column-major as in MATLAB.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import itable as it
from .itable import InstructionTable

_TABLE: Optional[InstructionTable] = None


def table() -> InstructionTable:
    global _TABLE
    if _TABLE is None:
        _TABLE = InstructionTable()
    return _TABLE


def clear() -> None:
    """``Tcalculus.clear`` -- forget all symbolic expressions."""
    global _TABLE
    _TABLE = InstructionTable()


def _obj(shape, fill=None) -> np.ndarray:
    a = np.empty(shape, dtype=object)
    a[...] = fill
    return a


Scalar = Union[int, float]


class Tcalculus:
    __array_priority__ = 1000  # numpy defers to our reflected operators

    def __init__(self, e: np.ndarray, name: str = ""):
        assert e.dtype == object
        self.e = e
        self.name = name

    # -- basic properties ----------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return self.e.shape

    size = shape

    def numel(self) -> int:
        return int(self.e.size)

    def __len__(self) -> int:
        return self.e.shape[0] if self.e.ndim else 1

    def flat(self) -> List[Optional[int]]:
        """entries in column-major order"""
        return list(self.e.reshape(-1, order="F"))

    def is_variable(self) -> bool:
        tab = table()
        return bool(self.name) and all(x is not None and tab.types[x] == it.I_set for x in self.flat())

    def __repr__(self) -> str:
        nz = sum(x is not None for x in self.flat())
        return f"Tcalculus(name={self.name!r}, shape={self.shape}, nnz={nz})"

    # -- elementwise helpers ---------------------------------------------------------
    @staticmethod
    def _lift(x) -> "Tcalculus":
        if isinstance(x, Tcalculus):
            return x
        return Tconstant(x)

    def _binary(self, other, fn) -> "Tcalculus":
        a, b = self, Tcalculus._lift(other)
        ea, eb = a.e, b.e
        if ea.shape != eb.shape:
            if ea.size == 1:
                ea = np.broadcast_to(ea.reshape(()), eb.shape)
            elif eb.size == 1:
                eb = np.broadcast_to(eb.reshape(()), ea.shape)
            else:
                raise ValueError(f"size mismatch {a.shape} vs {b.shape}")
        out = _obj(ea.shape)
        for idx in np.ndindex(*ea.shape):
            out[idx] = fn(ea[idx], eb[idx])
        return Tcalculus(out)

    def _unary(self, fn) -> "Tcalculus":
        out = _obj(self.shape)
        for idx in np.ndindex(*self.shape):
            out[idx] = fn(self.e[idx])
        return Tcalculus(out)

    # -- arithmetic -------------------------------------------------------------------
    def __add__(self, o):
        tab = table()
        return self._binary(o, lambda x, y: tab.add([(1, x), (1, y)]))

    __radd__ = __add__

    def __sub__(self, o):
        tab = table()
        return self._binary(o, lambda x, y: tab.add([(1, x), (-1, y)]))

    def __rsub__(self, o):
        return Tcalculus._lift(o).__sub__(self)

    def __neg__(self):
        tab = table()
        return self._unary(tab.neg)

    def __mul__(self, o):
        """elementwise product (MATLAB ``.*``; scalars broadcast)"""
        tab = table()
        return self._binary(o, lambda x, y: tab.mul([x, y]))

    __rmul__ = __mul__

    def __truediv__(self, o):
        tab = table()

        def f(x, y):
            if y is None:
                raise ZeroDivisionError("./ by a structural zero")
            return tab.div(x, y)
        return self._binary(o, f)

    def __rtruediv__(self, o):
        return Tcalculus._lift(o).__truediv__(self)

    def __pow__(self, p):
        if p == 2:
            return sqr(self)
        if p == 3:
            return cube(self)
        if p == 1:
            return self
        raise NotImplementedError("only .^2 and .^3")

    def __matmul__(self, o):
        return mtimes(self, o)

    def __rmatmul__(self, o):
        return mtimes(Tcalculus._lift(o), self)

    # -- indexing / shape -----------------------------------------------------------
    def __getitem__(self, idx) -> "Tcalculus":
        r = self.e[idx]
        if not isinstance(r, np.ndarray):
            r0 = _obj(())
            r0[()] = r
            r = r0
        return Tcalculus(r)

    @property
    def T(self) -> "Tcalculus":
        return Tcalculus(self.e.T.copy())

    # -- constraints (Tcalculus.m:2334-2413) ------------------------------------------
    def __ge__(self, o):
        return Constraint("ispositive", self - o)

    __gt__ = __ge__

    def __le__(self, o):
        return Constraint("ispositive", Tcalculus._lift(o) - self)

    __lt__ = __le__

    def __eq__(self, o):  # type: ignore[override]
        return Constraint("iszero", self - o)

    __hash__ = None  # type: ignore[assignment]


class Constraint:
    """``iszero(expr)`` or ``ispositive(expr)``"""

    def __init__(self, kind: str, expr: Tcalculus):
        self.kind = kind
        self.expr = expr


# -------------------------------------------------------------------------------------
# constructors
# -------------------------------------------------------------------------------------
def _shape(shape) -> Tuple[int, ...]:
    if shape is None:
        return ()
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def Tvariable(name: str, shape=()) -> Tcalculus:
    """``Tvariable name [sizes]`` -- full tensor of fresh I_set instructions."""
    tab = table()
    shp = _shape(shape)
    n = int(np.prod(shp)) if shp else 1
    ids = [tab.new_set(f"{name}[{k}]") for k in range(n)]  # column-major numbering
    flat = np.empty(n, dtype=object)
    flat[:] = ids
    e = flat.reshape(shp, order="F") if shp else flat.reshape(())
    return Tcalculus(e.copy(), name)


def Tconstant(value, shape=None) -> Tcalculus:
    tab = table()
    v = np.asarray(value, dtype=float)
    if shape is not None:
        v = v.reshape(_shape(shape), order="F")
    e = _obj(v.shape)
    for idx in np.ndindex(*v.shape):
        e[idx] = tab.const(v[idx])
    return Tcalculus(e)


def Tzeros(*shape) -> Tcalculus:
    shp = _shape(shape[0]) if len(shape) == 1 else _shape(shape)
    return Tcalculus(_obj(shp))


def Tones(*shape) -> Tcalculus:
    shp = _shape(shape[0]) if len(shape) == 1 else _shape(shape)
    return Tcalculus(_obj(shp, table().const(1.0)))


def Teye(*shape) -> Tcalculus:
    shp = _shape(shape[0]) if len(shape) == 1 else _shape(shape)
    if len(shp) == 1:
        shp = (shp[0], shp[0])
    e = _obj(shp)
    one = table().const(1.0)
    for k in range(min(shp)):
        e[k, k] = one
    return Tcalculus(e)


# -------------------------------------------------------------------------------------
# functions
# -------------------------------------------------------------------------------------
def _compose(x: Tcalculus, typ: int) -> Tcalculus:
    tab = table()
    return Tcalculus._lift(x)._unary(lambda v: tab.unary(typ, v))


def sqrt(x):
    return _compose(x, it.I_sqrt)


def sqr(x):
    return _compose(x, it.I_sqr)


def cube(x):
    return _compose(x, it.I_cube)


def exp(x):
    return _compose(x, it.I_exp)


def log(x):
    return _compose(x, it.I_log)


def sin(x):
    return _compose(x, it.I_sin)


def cos(x):
    return _compose(x, it.I_cos)


def atan(x):
    return _compose(x, it.I_atan)


def abs_(x):
    return _compose(x, it.I_abs)


def relu(x):
    return _compose(x, it.I_relu)


def srelu(x):
    return _compose(x, it.I_srelu)


def componentwise(x, fun, cfun: str, derivative1=None, derivative2=None) -> Tcalculus:
    """``y=componentwise(x,fun,cfun,derivative1,derivative2)`` (lib/@Tcalculus/Tcalculus.m:1836-1880): a user-defined
    scalar function applied entry by entry.  ``fun`` is its Python form (constant folding, structural zeros),
    ``cfun`` the C text with ``%s`` for the operand that the back ends compile (``I_componentwise``),
    ``derivative1(x)`` returns the derivative as a Tcalculus expression (which may itself be a
    ``componentwise`` carrying ``derivative2`` as ITS derivative, as in the reference's example
    ``componentwise(x,@sin,'sin(%s)',@(x)componentwise(x,@cos,'cos(%s)',@(x)-componentwise(x,@sin,'sin(%s)')))``)."""
    if not callable(fun):
        raise TypeError("componentwise 2nd argument must be a function")
    tab = table()
    d1 = derivative1
    if derivative1 is not None and derivative2 is not None:
        def d1(z, _d1=derivative1, _d2=derivative2):   # the derivative's own derivative rides on its instruction
            y = _d1(z)
            for i in Tcalculus._lift(y).flat():
                if i is not None and tab.types[i] == it.I_componentwise:
                    text = it.componentwise_text(tab.params[i], "%s")
                    f0, dd = tab.cw_funcs.get(text, (None, None))
                    if dd is None:
                        tab.cw_funcs[text] = (f0, _d2)
            return y
    return Tcalculus._lift(x)._unary(lambda v: tab.componentwise(cfun, v, fun, d1))


def mtimes(a, b) -> Tcalculus:
    """matrix product ``A*B`` (scalars broadcast like MATLAB's ``*``)"""
    a, b = Tcalculus._lift(a), Tcalculus._lift(b)
    if a.e.size == 1 or b.e.size == 1:
        return a * b
    tab = table()
    if a.e.ndim == 1 and b.e.ndim == 1:
        if a.shape != b.shape:
            raise ValueError("inner dimensions differ")
        out = _obj(())
        out[()] = tab.dot((x, y) for x, y in zip(a.e, b.e) if x is not None and y is not None)
        return Tcalculus(out)
    A = a.e if a.e.ndim == 2 else a.e.reshape(1, -1)
    B = b.e if b.e.ndim == 2 else b.e.reshape(-1, 1)
    if A.shape[1] != B.shape[0]:
        raise ValueError(f"inner dimensions differ {a.shape} * {b.shape}")
    out = _obj((A.shape[0], B.shape[1]))
    rowsA = [[(k, A[i, k]) for k in range(A.shape[1]) if A[i, k] is not None] for i in range(A.shape[0])]
    for j in range(B.shape[1]):
        colB = B[:, j]
        for i, ra in enumerate(rowsA):
            out[i, j] = tab.dot((x, colB[k]) for k, x in ra if colB[k] is not None)
    if a.e.ndim == 1:
        out = out.reshape(B.shape[1])
    elif b.e.ndim == 1:
        out = out.reshape(A.shape[0])
    return Tcalculus(out)


def tsum(x: Tcalculus, dim: Optional[int] = None) -> Tcalculus:
    """``sum(x,dim)`` with a 0-based dimension; ``dim=None`` sums everything."""
    tab = table()
    if dim is None:
        out = _obj(())
        out[()] = tab.add((1, v) for v in x.flat())
        return Tcalculus(out)
    moved = np.moveaxis(x.e, dim, 0)
    out = _obj(moved.shape[1:])
    for idx in np.ndindex(*moved.shape[1:]):
        out[idx] = tab.add((1, moved[(k,) + idx]) for k in range(moved.shape[0]))
    return Tcalculus(out)


def norm2(x: Tcalculus) -> Tcalculus:
    """sum of squares of all entries (``lib/norm2.m``; I_plus_sqr)"""
    out = _obj(())
    out[()] = table().plus_sqr(x.flat())
    return Tcalculus(out)


def norm1(x: Tcalculus) -> Tcalculus:
    out = _obj(())
    out[()] = table().plus_abs(x.flat())
    return Tcalculus(out)


def norminf(x: Tcalculus) -> Tcalculus:
    out = _obj(())
    out[()] = table().max_abs(x.flat())
    return Tcalculus(out)


def _minmax(x, y, dim, is_min: bool) -> Tcalculus:
    tab = table()
    if y is not None:
        return Tcalculus._lift(x)._binary(y, lambda a, b: tab.minmax([a, b], is_min))
    if dim is None:
        out = _obj(())
        out[()] = tab.minmax(x.flat(), is_min)
        return Tcalculus(out)
    moved = np.moveaxis(x.e, dim, 0)
    out = _obj(moved.shape[1:])
    for idx in np.ndindex(*moved.shape[1:]):
        out[idx] = tab.minmax([moved[(k,) + idx] for k in range(moved.shape[0])], is_min)
    return Tcalculus(out)


def tmin(x, y=None, dim: Optional[int] = None) -> Tcalculus:
    """``min(x,y)`` elementwise or ``min(x,[],dim)``"""
    return _minmax(x, y, dim, True)


def tmax(x, y=None, dim: Optional[int] = None) -> Tcalculus:
    return _minmax(x, y, dim, False)


def reshape(x: Tcalculus, shape) -> Tcalculus:
    return Tcalculus(x.e.reshape(_shape(shape), order="F").copy())


def repmat(x: Tcalculus, reps) -> Tcalculus:
    reps = _shape(reps)
    e = x.e
    while e.ndim < len(reps):
        e = e.reshape(e.shape + (1,))
    return Tcalculus(np.tile(e, reps))


def cat(dim: int, *xs) -> Tcalculus:
    arrs = [Tcalculus._lift(x).e for x in xs]
    nd = max(max(a.ndim for a in arrs), dim + 1)
    arrs = [a.reshape(a.shape + (1,) * (nd - a.ndim)) for a in arrs]
    return Tcalculus(np.concatenate(arrs, axis=dim))


def vertcat(*xs) -> Tcalculus:
    return cat(0, *xs)


def horzcat(*xs) -> Tcalculus:
    return cat(1, *xs)


def diag(x: Tcalculus) -> Tcalculus:
    if x.e.ndim == 1:
        n = x.shape[0]
        e = _obj((n, n))
        for k in range(n):
            e[k, k] = x.e[k]
        return Tcalculus(e)
    n = min(x.shape)
    e = _obj((n,))
    for k in range(n):
        e[k] = x.e[k, k]
    return Tcalculus(e)


def full(x: Tcalculus) -> Tcalculus:
    """materialise structural zeros as the constant 0 (``I_load 0``)"""
    tab = table()
    zero = tab.append_unique(it.I_load, (0.0,))
    return x._unary(lambda v: zero if v is None else v)


# -- time-series stencils (scalar sample time) -------------------------------------------
def tsDerivative(x: Tcalculus, ts) -> Tcalculus:
    """lib/tsDerivative.m, scalar-Ts branch: one-sided 3-point ends, central interior."""
    ts = Tcalculus._lift(ts)
    t1 = Tconstant([-1.5, 2, -.5]) / ts
    t2 = Tconstant(-.5) / ts
    t3 = Tconstant([.5, -2, 1.5]) / ts
    n = x.shape[0]
    first = reshape(mtimes(x[:, 0:3], t1), (n, 1))
    mid = t2 * (x[:, :-2] - x[:, 2:])
    last = reshape(mtimes(x[:, -3:], t3), (n, 1))
    return horzcat(first, mid, last)


def tsDerivative2(x: Tcalculus, ts) -> Tcalculus:
    """lib/tsDerivative2.m, scalar-Ts branch."""
    ts = Tcalculus._lift(ts)
    inv = Tconstant(1.0) / (ts * ts)
    n = x.shape[0]
    two = Tconstant(2.0)
    first = reshape(x[:, 0] - two * x[:, 1] + x[:, 2], (n, 1))
    mid = x[:, :-2] - two * x[:, 1:-1] + x[:, 2:]
    last = reshape(x[:, -3] - two * x[:, -2] + x[:, -1], (n, 1))
    return inv * horzcat(first, mid, last)


def tsIntegral(x: Tcalculus, ts) -> Tcalculus:
    """lib/tsIntegral.m, scalar-Ts branch: trapezoid weights [.5Ts,Ts,...,.5Ts]."""
    ts = Tcalculus._lift(ts)
    nt = x.shape[-1]
    half = Tconstant(.5) * ts
    dt = vertcat(reshape(half, (1,)), repmat(reshape(ts, (1,)), (nt - 2,)), reshape(half, (1,)))
    return mtimes(x, dt)


def tsCross(x1: Tcalculus, x2: Tcalculus) -> Tcalculus:
    """lib/tsCross.m"""
    a, b = [1, 2, 0], [2, 0, 1]
    return x1[a, :] * x2[b, :] - x1[b, :] * x2[a, :]
