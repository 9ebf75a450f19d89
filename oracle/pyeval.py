"""CPU ORACLE (test infrastructure only): numpy evaluation of scalar instructions.

A second, independent restatement of the per-instruction C forms printed by the reference's
``writeCinstructionsC`` (``lib/csparse/instructionsTableUTHash.c:950-1539``; table in SURVEY.md
Appendix A): every instruction is evaluated left to right in IEEE double, vectorised over a
batch (one numpy array of shape [B] per instruction).  Used by the front-end tests (symbolic
derivatives vs finite differences, ``L*D*L'=A`` identities, lazy-group semantics) where
compiling a C library per case would be slow.  Pinned by tests/test_oracle.py against the
compiled reference oracle.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from tenscalc_b200 import itable as it

DBL_MAX = np.finfo(np.float64).max

_UNARY = {
    it.I_inv: lambda x: 1 / x,                                     # :1427-1431
    it.I_minus_inv_sqr: lambda x: -1 / (x * x),
    it.I_2_inv_cube: lambda x: 2 / (x * x * x),
    it.I_exp: np.exp, it.I_log: np.log, it.I_cos: np.cos, it.I_minus_cos: lambda x: -np.cos(x),
    it.I_sin: np.sin, it.I_minus_sin: lambda x: -np.sin(x),
    it.I_round: lambda x: np.sign(x) * np.floor(np.abs(x) + .5),   # C round(): half away from zero
    it.I_ceil: np.ceil, it.I_floor: np.floor, it.I_abs: np.abs,
    it.I_sign: lambda x: np.where(x >= 0, 1.0, -1.0),              # :1371-1376, sign(0)=+1
    it.I_sqrt: np.sqrt,
    it.I_Dsqrt: lambda x: .5 * np.power(x, -.5), it.I_DDsqrt: lambda x: -.25 * np.power(x, -1.5),
    it.I_sqr: lambda x: x * x, it.I_2times: lambda x: 2 * x, it.I_cube: lambda x: x * x * x,
    it.I_3sqr: lambda x: 3 * x * x, it.I_6times: lambda x: 6 * x,
    it.I_atan: np.arctan, it.I_Datan: lambda x: 1 / (1 + x * x),
    it.I_DDatan: lambda x: -2 * x / (1 + x * x) / (1 + x * x),
    it.I_srelu: lambda x: np.log(1 + np.exp(x)), it.I_dsrelu: lambda x: 1 / (1 + np.exp(-x)),
    it.I_ddsrelu: lambda x: 1 / (2 + np.exp(-x) + np.exp(x)),
    it.I_relu: lambda x: np.where(x > 0, x, 0.0), it.I_heaviside: lambda x: np.where(x > 0, 1.0, 0.0),
}


def _dot_pairs(v, ops):
    acc = v[ops[0]] * v[ops[1]]
    for k in range(2, len(ops), 2):
        acc = acc + v[ops[k]] * v[ops[k + 1]]
    return acc


def eval_instruction(tab: it.InstructionTable, i: int, v: Dict[int, np.ndarray], B: int) -> np.ndarray:
    t, p, o = tab.types[i], tab.params[i], tab.operands[i]
    if t == it.I_load:
        return np.full(B, p[0])
    if t == it.I_2:
        return np.full(B, 2.0)
    if t == it.I_sum:                                              # :964-968
        acc = v[o[0]] if p[0] > 0 else -v[o[0]]
        for s, x in zip(p[1:], o[1:]):
            acc = acc + v[x] if s > 0 else acc - v[x]
        return acc
    if t == it.I_sumprod:                                          # :1007-1029
        m = int(p[0])
        per = m * -(-it.MAX_TERMS_PERLINE // m)                    # operands per emitted line (:1019-1026)
        total = None
        for k0 in range(0, len(o), per):                           # m[d]=line1; m[d]+=line2; ...
            acc = None
            for k in range(k0, min(k0 + per, len(o)), m):
                pr = v[o[k]]
                for x in o[k + 1:k + m]:
                    pr = pr * v[x]
                acc = pr if acc is None else acc + pr
            total = acc if total is None else total + acc
        return total
    if t == it.I_plus_minus_dot_div:                               # :1081-1098
        return (v[o[0]] - _dot_pairs(v, o[2:])) / v[o[1]]
    if t == it.I_minus_dot_div:                                    # :1106-1121
        return -_dot_pairs(v, o[1:]) / v[o[0]]
    if t == it.I_plus_minus_dot:                                   # :1128-1144
        return v[o[0]] - _dot_pairs(v, o[1:])
    if t == it.I_minus_dot:                                        # :1151-1164
        return -_dot_pairs(v, o)
    if t == it.I_div:
        return v[o[0]] / v[o[1]]
    if t == it.I_plus_sqr:
        acc = v[o[0]] * v[o[0]]
        for x in o[1:]:
            acc = acc + v[x] * v[x]
        return acc
    if t == it.I_plus_abs:
        acc = np.abs(v[o[0]])
        for x in o[1:]:
            acc = acc + np.abs(v[x])
        return acc
    if t in (it.I_min, it.I_max, it.I_min0, it.I_max0, it.I_max_abs):   # compare chains :1228-1313
        f = np.abs if t == it.I_max_abs else (lambda x: x)
        if t in (it.I_min0, it.I_max0):
            acc, rest = np.zeros(B), o
        else:
            acc, rest = f(v[o[0]]).copy(), o[1:]
        for x in rest:
            y = f(v[x])
            with np.errstate(invalid="ignore"):
                take = (acc > y) if t in (it.I_min, it.I_min0) else (acc < y)   # NaN compares false
            acc = np.where(take, y, acc)
        return acc
    if t == it.I_clp:                                              # :1315-1331
        acc = np.full(B, DBL_MAX)
        for k in range(0, len(o), 2):
            a, b = v[o[k]], v[o[k + 1]]
            with np.errstate(divide="ignore", invalid="ignore"):
                x = -a / b
                take = (b < 0) & (acc > x)
            acc = np.where(take, x, acc)
        return acc
    if t in _UNARY:
        with np.errstate(all="ignore"):
            return _UNARY[t](v[o[0]])
    if t == it.I_componentwise:                                    # :1333-1345, through the function's Python form
        f = tab.cw_funcs.get(it.componentwise_text(p, "%s"), (None, None))[0]
        if f is None:
            raise NotImplementedError("componentwise without a Python form")
        with np.errstate(all="ignore"):
            return np.vectorize(f, otypes=[np.float64])(np.broadcast_to(v[o[0]], (B,)))
    raise NotImplementedError(it.OPNAMES.get(t, str(t)))


def evaluate(tab: it.InstructionTable, roots: Iterable[Optional[int]], values: Dict[int, np.ndarray]):
    """values: {I_set instruction -> array[B] or scalar}; returns {root -> array[B]} (None -> zeros)"""
    roots = list(roots)
    B = max([np.size(x) for x in values.values()] + [1])
    v: Dict[int, np.ndarray] = {k: np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)) for k, x in values.items()}
    with np.errstate(all="ignore"):
        for i in tab.ancestors(r for r in roots if r is not None):
            if i in v:
                continue
            if tab.types[i] == it.I_set:
                raise KeyError(f"no value for variable {tab.names.get(i, i)}")
            v[i] = eval_instruction(tab, i, v, B)
    return [np.zeros(B) if r is None else v[r] for r in roots]
