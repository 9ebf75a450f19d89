"""Symbolic first derivatives on the scalar instruction table.

The reference differentiates at tensor level (``lib/@Tcalculus/gradient.m:75-390``)
and scalarises afterwards; here the same rules are applied per scalar instruction
(sum rule :113-132, product rule :134-171, ``norm2`` -> ``2*tprod(grad x,x)`` :180-184,
``compose`` chain rule with the derivative table of
``computeScalarInstructions.m:374-400`` :200-212, ``rdivide`` :306-349), which
yields the structurally sparse gradient directly: ``grad(i)`` is a dict
``{variable instruction -> derivative instruction}`` holding only the structural
non-zeros.  ``hessian(f,u)`` is ``gradient(gradient(f,u),u)`` as in
``Tcalculus.m:3097-3100``.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

from . import itable as it
from .itable import InstructionTable, Id


class Differentiator:
    def __init__(self, tab: InstructionTable, wrt: Iterable[int], sqrt_by_division: bool = True):
        self.tab = tab
        self.sqrt_by_division = sqrt_by_division and getattr(tab, "sqrt_by_division", True)
        self.wrt = set(wrt)
        self._memo: Dict[int, Dict[int, int]] = {}
        self._one = tab.const(1.0)

    def grad(self, root: Id) -> Dict[int, int]:
        if root is None:
            return {}
        tab = self.tab
        memo = self._memo
        # iterative post-order so that deep chains do not hit the recursion limit
        stack = [root]
        while stack:
            x = stack[-1]
            if x in memo:
                stack.pop()
                continue
            pending = [o for o in tab.operands[x] if o not in memo]
            if pending:
                stack.extend(pending)
                continue
            memo[x] = self._rule(x)
            stack.pop()
        return memo[root]

    # ------------------------------------------------------------------------------
    def _accumulate(self, contributions: List[Tuple[int, Tuple[Id, ...]]]) -> Id:
        """sum_k sign_k * prod(factors_k)"""
        tab = self.tab
        return tab.add([(s, tab.mul(f)) for s, f in contributions])

    def _rule(self, x: int) -> Dict[int, int]:
        tab = self.tab
        memo = self._memo
        t = tab.types[x]
        ops = tab.operands[x]
        if t == it.I_set:
            return {x: self._one} if x in self.wrt else {}
        if t in (it.I_load, it.I_2):
            return {}
        if not any(memo[o] for o in ops):
            return {}
        per_var: Dict[int, List[Tuple[int, Tuple[Id, ...]]]] = {}

        def contribute(sign: int, o: int, cofactors: Tuple[Id, ...]) -> None:
            for v, d in memo[o].items():
                per_var.setdefault(v, []).append((sign, cofactors + (d,)))

        if t == it.I_sum:
            for p, o in zip(tab.params[x], ops):
                contribute(1 if p > 0 else -1, o, ())
        elif t == it.I_sumprod:
            m = int(tab.params[x][0])
            for k in range(0, len(ops), m):
                prod = ops[k:k + m]
                for j, o in enumerate(prod):
                    if memo[o]:
                        contribute(1, o, prod[:j] + prod[j + 1:])
        elif t == it.I_div:
            a, b = ops
            if memo[a]:
                contribute(1, a, (tab.unary(it.I_inv, b),))
            if memo[b]:
                contribute(1, b, (a, tab.unary(it.I_minus_inv_sqr, b)))
        elif t == it.I_plus_sqr:
            two = tab.const(2.0)
            for o in ops:
                contribute(1, o, (o,))
            out = {}
            for v, c in per_var.items():
                r = tab.mul([two, self._accumulate(c)])
                if r is not None:
                    out[v] = r
            return out
        elif t in it.UNARY_TYPES:
            (o,) = ops
            if t in (it.I_abs, it.I_sign, it.I_round, it.I_ceil, it.I_floor, it.I_heaviside):
                raise NotImplementedError(f"derivative of {it.OPNAMES[t]} requested")
            if t == it.I_sqrt and self.sqrt_by_division:
                # d sqrt(x) = .5/sqrt(x), re-using the sqrt already on the table.  The reference's
                # table maps sqrt -> I_Dsqrt=.5*pow(x,-.5) -> I_DDsqrt=-.25*pow(x,-1.5)
                # (computeScalarInstructions.m:374-400); division and square root are correctly
                # rounded on both the host and the device whereas libm's pow() is not, so this
                # form keeps GPU and CPU-oracle results bit-identical (and is cheaper on both).
                contribute(1, o, (tab.div(tab.const(.5), x),))
            else:
                dt = tab.unary_derivative_type(t)
                contribute(1, o, (tab.unary(dt, o),))
        elif t == it.I_componentwise:
            # chain rule with the derivative the user supplied (Tcalculus.m:1836-1880), applied to the scalar operand
            (o,) = ops
            text = it.componentwise_text(tab.params[x], "%s")
            d = tab.cw_funcs.get(text, (None, None))[1]
            if d is None:
                raise NotImplementedError(f"componentwise '{text}': no derivative was given")
            from .tensor import Tcalculus, _obj
            e = _obj(())
            e[()] = o
            dy = Tcalculus._lift(d(Tcalculus(e)))
            if dy.numel() != 1:
                raise ValueError(f"componentwise '{text}': the derivative of a scalar must be a scalar")
            contribute(1, o, (dy.flat()[0],))
        else:
            raise NotImplementedError(f"derivative of {it.OPNAMES.get(t, t)} requested")
        out = {}
        for v, c in per_var.items():
            r = self._accumulate(c)
            if r is not None:
                out[v] = r
        return out


def gradient(tab: InstructionTable, f: Id, u: List[int]) -> List[Id]:
    d = Differentiator(tab, u).grad(f)
    return [d.get(v) for v in u]


def hessian(tab: InstructionTable, f: Id, u: List[int]):
    """Returns (H, g): g[i]=df/du_i, H[i]={j: d2f/du_i du_j} (structural non-zeros)."""
    D = Differentiator(tab, u)
    pos = {v: k for k, v in enumerate(u)}
    g0 = D.grad(f)
    g = [g0.get(v) for v in u]
    H = []
    for gi in g:
        row = D.grad(gi)
        H.append({pos[v]: e for v, e in row.items()})
    return H, g
