"""Scalar instruction table (hash-consing store) for the batched IPM backend.

Role in the reference: ``lib/csparse/instructionsTableUTHash.c`` (store, :175-204,
``appendUniqueInstruction`` :534-619) together with the opcode list of
``lib/instructionTypes.m:8-117``.  Every scalar value of a problem is one
*instruction* ``(type, parameters, operands)``; structurally identical
instructions are merged when they are created (hash-consing), ``I_set``
instructions are never merged (``lib/@csparse/csparse.m:1118,1150``).

This is a host-side, offline data structure: it is built once per problem
structure and then lowered to (i) the device tape executed by the sm_100a
kernels and (ii) C text for the CPU oracle.  Opcode numbers are the reference's
(``lib/csparse/instructionsTableTypes.h:2-56``) so that the same list can be fed
unchanged to the reference's own C emitter when the oracle is built.

Structural zeros are represented by ``None`` everywhere (the reference stores
only structurally non-zero entries, ``computeScalarInstructions.m:259-262``).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

# ---------------------------------------------------------------------------------
# opcodes (values = reference enum, instructionsTableTypes.h)
# ---------------------------------------------------------------------------------
I_set = 1
I_load = 2
I_sum = 3
I_sumprod = 4
I_inv = 5
I_minus_inv_sqr = 6
I_2_inv_cube = 7
I_div = 8
I_minus_dot = 9
I_minus_dot_div = 10
I_plus_minus_dot = 11
I_plus_sqr = 12
I_plus_abs = 13
I_plus_minus_dot_div = 14
I_min = 15
I_min0 = 16
I_max = 17
I_max0 = 18
I_max_abs = 19
I_clp = 20
I_exp = 21
I_log = 22
I_cos = 23
I_minus_cos = 24
I_sin = 25
I_minus_sin = 26
I_round = 27
I_ceil = 28
I_floor = 29
I_abs = 30
I_sign = 31
I_sqrt = 32
I_Dsqrt = 33
I_DDsqrt = 34
I_sqr = 35
I_2times = 36
I_2 = 37
I_cube = 38
I_3sqr = 39
I_6times = 40
I_atan = 41
I_Datan = 42
I_DDatan = 43
I_srelu = 44
I_dsrelu = 45
I_ddsrelu = 46
I_relu = 47
I_heaviside = 48
I_componentwise = 50   # user-defined scalar function given as C text (instructionsTableTypes.h:53; 49 = I_zero, never emitted)

OPNAMES = {v: k for k, v in list(globals().items()) if k.startswith("I_")}

# unary functions f(x) with their numeric definition (C forms of
# instructionsTableUTHash.c:1347-1539) -- used for constant folding only
_UNARY_EVAL = {
    I_inv: lambda x: 1 / x,
    I_minus_inv_sqr: lambda x: -1 / (x * x),
    I_2_inv_cube: lambda x: 2 / (x * x * x),
    I_exp: math.exp,
    I_log: math.log,
    I_cos: math.cos,
    I_minus_cos: lambda x: -math.cos(x),
    I_sin: math.sin,
    I_minus_sin: lambda x: -math.sin(x),
    I_round: lambda x: float(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1),
    I_ceil: lambda x: float(math.ceil(x)),
    I_floor: lambda x: float(math.floor(x)),
    I_abs: abs,
    I_sign: lambda x: 1.0 if x >= 0 else -1.0,
    I_sqrt: math.sqrt,
    I_Dsqrt: lambda x: 0.5 * x ** -0.5,
    I_DDsqrt: lambda x: -0.25 * x ** -1.5,
    I_sqr: lambda x: x * x,
    I_2times: lambda x: 2 * x,
    I_cube: lambda x: x * x * x,
    I_3sqr: lambda x: 3 * x * x,
    I_6times: lambda x: 6 * x,
    I_atan: math.atan,
    I_Datan: lambda x: 1 / (1 + x * x),
    I_DDatan: lambda x: -2 * x / (1 + x * x) / (1 + x * x),
    I_srelu: lambda x: math.log(1 + math.exp(x)),
    I_dsrelu: lambda x: 1 / (1 + math.exp(-x)),
    I_ddsrelu: lambda x: 1 / (2 + math.exp(-x) + math.exp(x)),
    I_relu: lambda x: x if x > 0 else 0.0,
    I_heaviside: lambda x: 1.0 if x > 0 else 0.0,
}
UNARY_TYPES = frozenset(_UNARY_EVAL)

# derivative chain of the unary functions (reference table:
# lib/@csparse/computeScalarInstructions.m:374-400)
_UNARY_DERIV = {
    I_log: I_inv,
    I_inv: I_minus_inv_sqr,
    I_minus_inv_sqr: I_2_inv_cube,
    I_exp: I_exp,
    I_sin: I_cos,
    I_cos: I_minus_sin,
    I_minus_sin: I_minus_cos,
    I_minus_cos: I_sin,
    I_sqrt: I_Dsqrt,
    I_Dsqrt: I_DDsqrt,
    I_sqr: I_2times,
    I_2times: I_2,
    I_cube: I_3sqr,
    I_3sqr: I_6times,
    I_atan: I_Datan,
    I_Datan: I_DDatan,
    I_srelu: I_dsrelu,
    I_dsrelu: I_ddsrelu,
    I_relu: I_heaviside,
}

# longest product list the reference emitter prints on one C line; longer lists
# get split into `m[r]=...; m[r]+=...;` which changes the association order
# (instructionsTableUTHash.c:173,1013-1022).  The front end never creates longer ones unless
# InstructionTable.split_long is switched off (then, like MATLAB's csparse, it leaves the line breaks to the
# emitters); tables read from .cg files may hold them.
MAX_TERMS_PERLINE = 100


def _cw_split(cfun: str):
    """characters of a componentwise C text, None at every ``%s``"""
    out = []
    parts = cfun.split("%s")
    for k, part in enumerate(parts):
        out += list(part)
        if k + 1 < len(parts):
            out.append(None)
    return out


def componentwise_text(params, operand: str) -> str:
    """C expression of an ``I_componentwise`` instruction: its characters with ``operand`` at every 0
    (instructionsTableUTHash.c:1337-1343)"""
    return "".join(operand if c == 0 else chr(int(c)) for c in params)


def sumprod_lines(m: int, o):
    """The emitted statement(s) of ``I_sumprod [m, s]`` over operand expressions ``o``: one string per LINE of
    the reference's C (instructionsTableUTHash.c:1007-1029).  The emitter counts ``m`` factors per product and
    breaks the statement once a line holds MAX_TERMS_PERLINE of them (:1019-1026), continuing with
    ``m[d]+=...``: each line is summed left to right on its own, then added to the running value."""
    per = -(-MAX_TERMS_PERLINE // m)            # products per line: the first count that reaches the limit
    prods = ["*".join(o[k:k + m]) for k in range(0, len(o), m)]
    return ["+".join(prods[k:k + per]) for k in range(0, len(prods), per)]

Id = Optional[int]


class InstructionTable:
    """Append-only table of scalar instructions with structural de-duplication."""

    def __init__(self) -> None:
        self.types: List[int] = []
        self.params: List[Tuple[float, ...]] = []
        self.operands: List[Tuple[int, ...]] = []
        self.names: Dict[int, str] = {}  # I_set only: debugging label
        self._index: Dict[tuple, int] = {}
        self.sqrt_by_division = True   # see differentiate.Differentiator
        self.cw_funcs: Dict[str, tuple] = {}   # componentwise C text -> (Python form, derivative)
        self.split_long = True     # chunk sums of products longer than one emitted line (see MAX_TERMS_PERLINE)

    # -- raw access --------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.types)

    def append(self, typ: int, params: Sequence[float] = (), operands: Sequence[int] = ()) -> int:
        self.types.append(typ)
        self.params.append(tuple(params))
        self.operands.append(tuple(operands))
        return len(self.types) - 1

    def append_unique(self, typ: int, params: Sequence[float] = (), operands: Sequence[int] = ()) -> int:
        key = (typ, tuple(params), tuple(operands))
        i = self._index.get(key)
        if i is None:
            i = self.append(typ, key[1], key[2])
            self._index[key] = i
        return i

    # -- leaves ------------------------------------------------------------------
    def new_set(self, name: str = "") -> int:
        i = self.append(I_set)
        if name:
            self.names[i] = name
        return i

    def const(self, c: float) -> Id:
        c = float(c)
        if c == 0.0:
            return None
        return self.append_unique(I_load, (c,))

    def is_const(self, i: Id) -> bool:
        return i is None or self.types[i] == I_load or self.types[i] == I_2

    def const_value(self, i: Id) -> float:
        if i is None:
            return 0.0
        if self.types[i] == I_2:
            return 2.0
        assert self.types[i] == I_load
        return self.params[i][0]

    # -- sums --------------------------------------------------------------------
    def _is_neg(self, i: int) -> bool:
        """i is a unary minus  -x  (I_sum with the single parameter -1)."""
        return self.types[i] == I_sum and len(self.operands[i]) == 1 and self.params[i][0] < 0

    def neg(self, a: Id) -> Id:
        return self.add([(-1, a)])

    def add(self, terms: Iterable[Tuple[int, Id]]) -> Id:
        """Signed sum  sum_k s_k*x_k  with s_k=+-1.

        Normalisation follows the reference's `plus` (lib/@Tcalculus/Tcalculus.m:2228-2311:
        nested sums flattened, zeros dropped) and operand sorting
        (lib/@csparse/csparse.m:1176-1178).  Constant terms are folded, +x-x cancels.
        """
        cval = 0.0
        count: Dict[int, int] = {}
        stack = list(terms)
        while stack:
            s, x = stack.pop()
            if x is None:
                continue
            t = self.types[x]
            if t == I_load or t == I_2:
                cval += s * self.const_value(x)
            elif t == I_sum:
                for p, o in zip(self.params[x], self.operands[x]):
                    stack.append((s if p > 0 else -s, o))
            else:
                count[x] = count.get(x, 0) + s
        flat: List[Tuple[int, int]] = []
        for x in sorted(count):
            c = count[x]
            flat.extend([(1 if c > 0 else -1, x)] * abs(c))
        if cval != 0.0:
            cid = self.const(abs(cval))
            flat.append((1 if cval > 0 else -1, cid))
            flat.sort(key=lambda sx: sx[1])
        if not flat:
            return None
        if len(flat) == 1 and flat[0][0] > 0:
            return flat[0][1]
        # all-positive products with the same number of factors -> one I_sumprod [m,s]
        if all(s > 0 and self.types[x] == I_sumprod for s, x in flat):
            ms = {self.params[x][0] for _, x in flat}
            nfac = sum(len(self.operands[x]) for _, x in flat)
            if len(ms) == 1 and (nfac <= MAX_TERMS_PERLINE or not self.split_long):
                prods: List[Tuple[int, ...]] = []
                m = int(next(iter(ms)))
                for _, x in flat:
                    ops = self.operands[x]
                    prods.extend(ops[k:k + m] for k in range(0, len(ops), m))
                return self._sumprod(m, prods)
        return self.append_unique(I_sum, tuple(float(s) for s, _ in flat), tuple(x for _, x in flat))

    def sub(self, a: Id, b: Id) -> Id:
        return self.add([(1, a), (-1, b)])

    # -- products ----------------------------------------------------------------
    def _sumprod(self, m: int, prods: List[Tuple[int, ...]]) -> int:
        # factors sorted inside each product, then products sorted (csparse.m:1170-1175)
        prods = sorted(tuple(sorted(p)) for p in prods)
        s = len(prods)
        if m * s > MAX_TERMS_PERLINE and self.split_long:
            # keep every emitted C statement on one line: chunk and add the chunks
            per = max(1, MAX_TERMS_PERLINE // m)
            parts = [self._sumprod(m, prods[k:k + per]) for k in range(0, s, per)]
            return self.append_unique(I_sum, (1.0,) * len(parts), tuple(sorted(parts)))
        flat = tuple(o for p in prods for o in p)
        return self.append_unique(I_sumprod, (float(m), float(s)), flat)

    def mul(self, factors: Iterable[Id]) -> Id:
        """Product of scalars.  Constants folded, nested single products flattened,
        unary minus pulled out (so that signs live in I_sum parameters)."""
        cval = 1.0
        sign = 1
        out: List[int] = []
        stack = list(factors)
        while stack:
            x = stack.pop()
            if x is None:
                return None
            t = self.types[x]
            if t == I_load or t == I_2:
                cval *= self.const_value(x)
            elif t == I_sumprod and self.params[x][1] == 1:
                stack.extend(self.operands[x])
            elif self._is_neg(x):
                sign = -sign
                stack.append(self.operands[x][0])
            else:
                out.append(x)
        if cval == 0.0:
            return None
        if cval < 0:
            sign, cval = -sign, -cval
        if not out:
            r = self.const(cval)
        else:
            if cval != 1.0:
                out.append(self.const(cval))
            out.sort()
            r = out[0] if len(out) == 1 else self.append_unique(
                I_sumprod, (float(len(out)), 1.0), tuple(out))
        return r if sign > 0 else self.neg(r)

    def dot(self, pairs: Iterable[Tuple[Id, ...]]) -> Id:
        """sum_k prod(pairs[k]) -- the scalarised `tprod` (sparsity_tprod.m:156-272)."""
        terms = [(1, self.mul(p)) for p in pairs]
        return self.add(terms)

    def div(self, a: Id, b: Id) -> Id:
        if b is None:
            raise ZeroDivisionError("division by a structural zero")
        if a is None:
            return None
        if self.is_const(b):
            bv = self.const_value(b)
            if self.is_const(a):
                return self.const(self.const_value(a) / bv)
            if bv == 1.0:
                return a
            if bv == -1.0:
                return self.neg(a)
        if self._is_neg(a):
            return self.neg(self.div(self.operands[a][0], b))
        if self.is_const(a) and self.const_value(a) == 1.0:
            return self.unary(I_inv, b)
        return self.append_unique(I_div, (), (a, b))

    # -- unary functions ---------------------------------------------------------
    def componentwise(self, cfun: str, x: Id, pyfun=None, derivative=None) -> Id:
        """``I_componentwise`` (sparsity_componentwise.m:33-35): the user's C text ``cfun`` with every ``%s``
        standing for the operand; the parameter vector holds the characters of the text, 0 at every operand
        position (the form writeCinstructionsC prints, instructionsTableUTHash.c:1333-1345).  ``pyfun`` (a
        Python callable of a float) folds constants, ``derivative`` (a callable of a Tcalculus returning a
        Tcalculus) is the chain rule's factor; both are remembered per C text."""
        if "%s" not in cfun:
            raise ValueError("componentwise: the C text must hold %s where the operand goes")
        if pyfun is not None or derivative is not None or cfun not in self.cw_funcs:
            old = self.cw_funcs.get(cfun, (None, None))
            self.cw_funcs[cfun] = (pyfun or old[0], derivative or old[1])
        if self.is_const(x):   # (structural zeros included: the reference fills them with I_load 0 when f(0) is not 0)
            f = self.cw_funcs[cfun][0]
            if f is None:
                raise ValueError("componentwise of a constant needs the Python form of the function")
            return self.const(float(f(self.const_value(x))))
        params = tuple(0.0 if part is None else float(ord(part)) for part in _cw_split(cfun))
        return self.append_unique(I_componentwise, params, (x,))

    def unary(self, typ: int, x: Id) -> Id:
        assert typ in UNARY_TYPES or typ == I_2
        if typ == I_2:
            return self.const(2.0)
        if self.is_const(x):
            return self.const(_UNARY_EVAL[typ](self.const_value(x)))
        if typ == I_2times:
            return self.mul([self.const(2.0), x]) if self._is_neg(x) else self.append_unique(typ, (), (x,))
        return self.append_unique(typ, (), (x,))

    # -- reductions --------------------------------------------------------------
    def plus_sqr(self, xs: Sequence[Id]) -> Id:
        ops = sorted(x for x in xs if x is not None)
        if not ops:
            return None
        if len(ops) > MAX_TERMS_PERLINE // 2 and self.split_long:
            per = MAX_TERMS_PERLINE // 2
            return self.add([(1, self.plus_sqr(ops[k:k + per])) for k in range(0, len(ops), per)])
        return self.append_unique(I_plus_sqr, (), tuple(ops))

    def plus_abs(self, xs: Sequence[Id]) -> Id:
        ops = sorted(x for x in xs if x is not None)
        if not ops:
            return None
        return self.append_unique(I_plus_abs, (), tuple(ops))

    def max_abs(self, xs: Sequence[Id]) -> Id:
        """norminf (sparsity_norminf.m): fabs chain, instructionsTableUTHash.c:1296-1313."""
        ops = sorted(set(x for x in xs if x is not None))
        if not ops:
            return None
        if len(ops) == 1:
            return self.unary(I_abs, ops[0])
        return self.append_unique(I_max_abs, (), tuple(ops))

    def minmax(self, xs: Sequence[Id], is_min: bool) -> Id:
        """min/max over a list; structural zeros take part through the `...0`
        variants (sparsity_min.m; instructionsTableUTHash.c:1228-1294)."""
        has_zero = any(x is None for x in xs)
        ops = sorted(set(x for x in xs if x is not None))
        if not ops:
            return None
        if has_zero:
            return self.append_unique(I_min0 if is_min else I_max0, (), tuple(ops))
        if len(ops) == 1:
            return ops[0]
        return self.append_unique(I_min if is_min else I_max, (), tuple(ops))

    def clp(self, pairs: Sequence[Tuple[Id, Id]]) -> Id:
        """largest alpha>=0 with a+alpha*b>=0 (instructionsTableUTHash.c:1315-1331).
        Pairs whose b is a structural zero never bind and are dropped; a structural
        zero `a` is materialised as the constant 0 would be in `full()`."""
        ops: List[Tuple[int, int]] = []
        for a, b in pairs:
            if b is None:
                continue
            if a is None:
                raise ValueError("clp(x,dx): pattern(dx) must be inside pattern(x) (sparsity_clp.m)")
            ops.append((a, b))
        if not ops:
            return None
        ops.sort()
        return self.append_unique(I_clp, (), tuple(o for p in ops for o in p))

    # -- LDL-specific forms (sparsity_ldl.m:144-236) --------------------------------
    def plus_minus_dot(self, a0: Id, pairs: Sequence[Tuple[int, int]]) -> Id:
        pairs = sorted(pairs)
        flat = tuple(o for p in pairs for o in p)
        if not pairs:
            return a0
        if a0 is None:
            return self.append_unique(I_minus_dot, (), flat)
        return self.append_unique(I_plus_minus_dot, (), (a0,) + flat)

    def plus_minus_dot_div(self, a0: Id, d: int, pairs: Sequence[Tuple[int, int]]) -> Id:
        pairs = sorted(pairs)
        flat = tuple(o for p in pairs for o in p)
        if not pairs:
            return None if a0 is None else self.append_unique(I_div, (), (a0, d))
        if a0 is None:
            return self.append_unique(I_minus_dot_div, (), (d,) + flat)
        return self.append_unique(I_plus_minus_dot_div, (), (a0, d) + flat)

    # -- derivative of the unary chain ----------------------------------------------
    def unary_derivative_type(self, typ: int) -> int:
        if typ not in _UNARY_DERIV:
            raise NotImplementedError(f"derivative of {OPNAMES[typ]} is not available")
        return _UNARY_DERIV[typ]

    # -- graph queries -------------------------------------------------------------
    def ancestors(self, roots: Iterable[int]) -> List[int]:
        """All instructions the roots depend on (roots included), ascending ids."""
        seen = set()
        stack = [r for r in roots if r is not None]
        while stack:
            x = stack.pop()
            if x in seen:
                continue
            seen.add(x)
            stack.extend(o for o in self.operands[x] if o not in seen)
        return sorted(seen)

    # -- structural rebuild / substitution -------------------------------------------
    def rebuild(self, i: int, new_ops: Sequence[Id]) -> Id:
        """Re-create instruction ``i`` with other operands, through the normalising
        constructors (so that zeros/constants that appear are folded)."""
        t = self.types[i]
        p = self.params[i]
        if t == I_sum:
            return self.add((1 if s > 0 else -1, o) for s, o in zip(p, new_ops))
        if t == I_sumprod:
            m = int(p[0])
            return self.dot(tuple(new_ops[k:k + m]) for k in range(0, len(new_ops), m))
        if t == I_div:
            return self.div(new_ops[0], new_ops[1])
        if t in UNARY_TYPES:
            return self.unary(t, new_ops[0])
        if t == I_componentwise:
            return self.componentwise(componentwise_text(p, "%s"), new_ops[0])
        if t == I_plus_sqr:
            return self.plus_sqr(new_ops)
        if t == I_plus_abs:
            return self.plus_abs(new_ops)
        if t == I_max_abs:
            return self.max_abs(new_ops)
        if t in (I_min, I_min0, I_max, I_max0):
            ops = list(new_ops) + ([None] if t in (I_min0, I_max0) else [])
            return self.minmax(ops, t in (I_min, I_min0))
        raise NotImplementedError(f"rebuild of {OPNAMES.get(t, t)}")

    def substitute(self, roots: Sequence[Id], mapping: Dict[int, Id]) -> List[Id]:
        """``substitute(expr,var,newExpr)`` of the reference, for lists of scalars:
        every occurrence of a key of ``mapping`` is replaced by its value."""
        memo: Dict[int, Id] = dict(mapping)
        for i in self.ancestors(roots):
            if i in memo:
                continue
            ops = self.operands[i]
            new_ops = [memo.get(o, o) for o in ops]
            if all(a == b for a, b in zip(ops, new_ops)):
                memo[i] = i
            else:
                memo[i] = self.rebuild(i, new_ops)
        return [None if r is None else memo[r] for r in roots]
