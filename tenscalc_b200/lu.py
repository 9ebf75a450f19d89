"""Fixed-sparsity-pattern LU factorisation and solves, as instructions (non-symmetric KKT matrices of
the equilibrium solvers, ``ipm_builder_eq``).

Restates ``lib/@csparse/sparsity_lu.m`` -- the row/column permutations come from a NUMERIC sparse LU with
threshold pivoting of a matrix of typical values with the same pattern (the reference calls MATLAB's
``lu(A,.01,'vector')`` on ``1+.5*rand`` per distinct instruction, :42-50,82; here SciPy's SuperLU with COLAMD
and ``diag_pivot_thresh=.01`` on the same kind of values), then a right-looking symbolic elimination without
further pivoting (:150-221):

    L(row,col) = A(row,col)/A(pivot,col)                          I_div
    A(row,c)   = A(row,c) - L(row,col)*A(pivot,c)    c > col      I_plus_minus_dot / I_minus_dot (fill)

-- and ``sparsity_mldivide_l1.m:66-102`` / ``sparsity_mldivide_u.m:66-92``: permute b by p, forward substitution
with the unit lower factor, backward substitution with U (``I_plus_minus_dot_div`` / ``I_minus_dot_div`` /
``I_div``), inverse column permutation.  The permutations are part of the generated code, as in the
reference: a pivot that is structurally fine but numerically zero at run time is not detected.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .itable import InstructionTable, Id

SparseRows = List[Dict[int, int]]          # rows[i] = {j: instruction}


class LUFactor:
    def __init__(self, n: int, p: List[int], q: List[int], LU: SparseRows):
        self.n = n
        self.p = p          # row k of the factored matrix  <-  row p[k] of A
        self.q = q          # column k                      <-  column q[k] of A
        self.LU = LU        # LU[r] = {c: instr}: c < r strictly lower (unit diagonal implied), c >= r upper

    def nnz(self) -> int:
        return sum(len(r) for r in self.LU)


def pivot_order(A: SparseRows, seed: int = 0):
    """(p, q) of a sparse LU with threshold partial pivoting on typical values (sparsity_lu.m:42-50,82)"""
    from scipy.sparse import csc_matrix
    from scipy.sparse.linalg import splu
    n = len(A)
    ids = sorted({x for r in A for x in r.values()})
    rng = np.random.RandomState(seed)
    val = dict(zip(ids, 1 + .5 * rng.rand(len(ids))))       # one value per DISTINCT instruction: structural symmetries survive
    ri, ci, vv = [], [], []
    for i, r in enumerate(A):
        for j, x in r.items():
            ri.append(i); ci.append(j); vv.append(val[x])
    M = csc_matrix((vv, (ri, ci)), shape=(n, n))
    lu = splu(M, permc_spec="COLAMD", diag_pivot_thresh=.01, options=dict(SymmetricMode=False))
    p = np.argsort(lu.perm_r)       # Pr[perm_r[i], i] = 1: row i of A lands in row perm_r[i]
    q = np.argsort(lu.perm_c)       # Pc[i, perm_c[i]] = 1: column i of A lands in column perm_c[i]
    return [int(x) for x in p], [int(x) for x in q]


def lu(tab: InstructionTable, A: SparseRows, order=None) -> LUFactor:
    """Symbolic LU of the square matrix ``A`` (rows of {column: instruction}).  ``order``: None (numeric
    threshold pivoting on typical values) or an explicit pair (p, q)."""
    n = len(A)
    p, q = pivot_order(A) if order is None else (list(order[0]), list(order[1]))
    if sorted(p) != list(range(n)) or sorted(q) != list(range(n)):
        raise ValueError("lu: the ordering is not a pair of permutations of 0..n-1")
    qinv = [0] * n
    for k, o in enumerate(q):
        qinv[o] = k
    M: SparseRows = [{qinv[j]: x for j, x in A[p[k]].items()} for k in range(n)]
    remaining = list(range(n))
    pp: List[int] = []
    for col in range(n):
        rows = [r for r in remaining if col in M[r]]
        if not rows:
            raise ValueError(f"lu: no pivot in column {col} (sparsity_lu.m:158-161)")
        pivot = rows[0]                                  # first structurally non-zero row (:164)
        remaining.remove(pivot)
        pp.append(pivot)
        prow = M[pivot]
        upper = sorted(c for c in prow if c > col)
        for row in rows[1:]:
            mr = M[row]
            l = tab.div(mr[col], prow[col])              # :181-183
            mr[col] = l
            for c in upper:
                if c in mr:                              # :196-203
                    mr[c] = tab.plus_minus_dot(mr[c], [(l, prow[c])])
                else:                                    # fill: :216-221
                    mr[c] = tab.plus_minus_dot(None, [(l, prow[c])])
                if mr[c] is None:
                    del mr[c]
    return LUFactor(n, [p[k] for k in pp], q, [M[k] for k in pp])


def solve(tab: InstructionTable, F: LUFactor, b: Sequence[Id]) -> List[Id]:
    """x = A \\ b for one right-hand side given as a list of instructions"""
    n = F.n
    y: List[Id] = [b[F.p[k]] for k in range(n)]                                  # b(p)
    for row in range(1, n):                                                        # (I+L) y = b(p)
        pairs = [(l, y[c]) for c, l in sorted(F.LU[row].items()) if c < row and y[c] is not None]
        y[row] = tab.plus_minus_dot(y[row], pairs)
    for row in range(n - 1, -1, -1):                                               # U x = y
        d = F.LU[row].get(row)
        if d is None:
            raise ValueError("mldivide_u: zero in the diagonal (sparsity_mldivide_u.m:70-72)")
        pairs = [(u, y[c]) for c, u in sorted(F.LU[row].items()) if c > row and y[c] is not None]
        y[row] = tab.plus_minus_dot_div(y[row], d, pairs)
    x: List[Id] = [None] * n
    for k in range(n):
        x[F.q[k]] = y[k]                                                           # inverse column permutation
    return x


def evaluate(tab: InstructionTable, ids: Sequence[Optional[int]], values: Dict[int, float]) -> List[float]:
    """numeric evaluation of instructions (tests): ``values`` gives the I_set leaves"""
    from . import itable as it
    memo: Dict[int, float] = dict(values)

    def ev(i: int) -> float:
        if i in memo:
            return memo[i]
        t, p, o = tab.types[i], tab.params[i], tab.operands[i]
        v = [ev(x) for x in o]
        if t == it.I_load:
            r = p[0]
        elif t == it.I_2:
            r = 2.0
        elif t == it.I_sum:
            r = sum((1 if s > 0 else -1) * x for s, x in zip(p, v))
        elif t == it.I_sumprod:
            m = int(p[0])
            r = sum(float(np.prod(v[k:k + m])) for k in range(0, len(v), m))
        elif t == it.I_div:
            r = v[0] / v[1]
        elif t == it.I_plus_minus_dot:
            r = v[0] - sum(v[k] * v[k + 1] for k in range(1, len(v), 2))
        elif t == it.I_minus_dot:
            r = -sum(v[k] * v[k + 1] for k in range(0, len(v), 2))
        elif t == it.I_plus_minus_dot_div:
            r = (v[0] - sum(v[k] * v[k + 1] for k in range(2, len(v), 2))) / v[1]
        elif t == it.I_minus_dot_div:
            r = -sum(v[k] * v[k + 1] for k in range(1, len(v), 2)) / v[0]
        else:
            raise NotImplementedError(it.OPNAMES.get(t, str(t)))
        memo[i] = r
        return r
    return [0.0 if i is None else ev(i) for i in ids]
