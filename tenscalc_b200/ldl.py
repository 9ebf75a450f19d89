"""Fixed-sparsity-pattern LDL' factorisation and triangular solves, as instructions.

Restates the recipe of ``lib/@csparse/sparsity_ldl.m:144-236`` (left-looking column
LDL' on the symmetrically permuted pattern, strictly diagonal D, no numerical
pivoting, symbolic fill) and of ``sparsity_mldivide_l1.m:66-102`` /
``sparsity_mldivide_d.m:60-74`` / ``sparsity_mldivide_u1.m:67-104`` (permute b,
forward substitution, divide by D, backward substitution, inverse permutation).

Ordering: the reference calls MATLAB's ``symamd(A,[1000,0])`` (sparsity_ldl.m:91); so does
this module by default (``ordering.py``: COLAMD's ``symamd_l`` with the same knobs, committed
permutation cache for boxes without the SuiteSparse sources).  Any symmetric ordering is
admissible for the quasi-definite KKT matrices assembled in ``ipm_builder`` -- only fill, not
correctness, depends on it -- and ``order`` may name another method or be an explicit permutation.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .itable import InstructionTable, Id

SparseSym = List[Dict[int, int]]  # rows[i] = {j: instruction}  (full symmetric pattern)


from .ordering import minimum_degree as minimum_degree_ordering  # noqa: E402,F401  (round-1 name)
from . import ordering as _ordering  # noqa: E402


class LDLFactor:
    def __init__(self, n: int, p: List[int], L: List[Dict[int, int]], D: List[int]):
        self.n = n
        self.p = p          # permuted index k  <-  original index p[k]
        self.L = L          # L[row] = {col: instr}, col<row, in permuted indices
        self.D = D          # D[k] instr

    def nnz(self) -> int:
        return sum(len(r) for r in self.L) + self.n


def ldl(tab: InstructionTable, A: SparseSym, order=None) -> LDLFactor:
    """Symbolic LDL' of the symmetric matrix whose LOWER triangle is taken from ``A``.
    ``order``: None (default method of ``ordering.py``: symamd), a method name, a callable
    ``A -> permutation`` or an explicit permutation."""
    n = len(A)
    if order is None or isinstance(order, str) or callable(order):
        order = _ordering.order(A, order)
    p = list(order)
    if sorted(p) != list(range(n)):
        raise ValueError("ldl: the ordering is not a permutation of 0..n-1")
    inv = [0] * n
    for k, o in enumerate(p):
        inv[o] = k
    # permuted lower triangle: Ap[row][col], col<=row, value taken from the lower entry
    Ap: List[Dict[int, int]] = [dict() for _ in range(n)]
    for i, row in enumerate(A):
        for j, x in row.items():
            if j > i:
                continue  # lower entries win (sparsity_ldl.m: "LOWER ONES USED")
            r, c = inv[i], inv[j]
            if r < c:
                r, c = c, r
            Ap[r][c] = x
    # mirror for entries only present in the upper triangle
    for i, row in enumerate(A):
        for j, x in row.items():
            if j > i and i not in A[j]:
                r, c = inv[i], inv[j]
                if r < c:
                    r, c = c, r
                Ap[r].setdefault(c, x)

    L: List[Dict[int, int]] = [dict() for _ in range(n)]      # by row
    Lcol: List[Dict[int, int]] = [dict() for _ in range(n)]   # by column
    Acol: List[List[int]] = [[] for _ in range(n)]
    for r in range(n):
        for c in Ap[r]:
            if c < r:
                Acol[c].append(r)
    D: List[Id] = [None] * n
    for col in range(n):
        # v_i = L(col,i)*D(i)
        v = {i: tab.mul([lci, D[i]]) for i, lci in L[col].items()}
        pairs = [(L[col][i], v[i]) for i in sorted(v)]
        a0 = Ap[col].get(col)
        if a0 is None and not pairs:
            raise ValueError("ldl needs pivoting: structurally zero diagonal entry (sparsity_ldl.m:187-189)")
        D[col] = tab.plus_minus_dot(a0, pairs)
        # rows below with a structural entry: A(row,col)!=0 or L(row,1:col-1)*v!=0
        rows = set(Acol[col])
        for i in v:
            rows.update(r for r in Lcol[i] if r > col)
        for row in sorted(rows):
            lr = L[row]
            pr = [(lr[i], v[i]) for i in sorted(v) if i in lr]
            x = tab.plus_minus_dot_div(Ap[row].get(col), D[col], pr)
            if x is not None:
                lr[col] = x
                Lcol[col][row] = x
    return LDLFactor(n, p, L, D)  # type: ignore[arg-type]


def dense_trailing_block(F: LDLFactor) -> int:
    """largest k such that the last k columns of the factor are completely dense -- L(r,c) structurally
    non-zero for all n-k <= c < r < n.  A large value (box-constrained dense least squares: the whole n x n
    Schur block once the diagonal inequality pivots are eliminated) means the scalar instruction list grows like
    k^3/3 multiply-adds and the problem belongs on the dense kernels (``dense.py``, ``csrc/tc_dense.cu``) whose
    trailing updates run on the FP64 tensor cores."""
    n = F.n
    k = 0
    while k < n:
        c = n - 1 - k
        if all(c in F.L[r] for r in range(c + 1, n)):
            k += 1
        else:
            break
    return k


def solve(tab: InstructionTable, F: LDLFactor, b: Sequence[Id]) -> List[Id]:
    """x = (P'LDL'P) \\ b  for one right-hand side given as a list of instructions."""
    n = F.n
    y: List[Id] = [b[F.p[k]] for k in range(n)]          # apply permutation
    for row in range(n):                                   # (I+L) y = b
        pairs = [(l, y[c]) for c, l in sorted(F.L[row].items()) if y[c] is not None]
        y[row] = tab.plus_minus_dot(y[row], pairs)
    for row in range(n):                                   # D z = y
        if y[row] is not None:
            y[row] = tab.div(y[row], F.D[row])
    Ucol: List[List[Tuple[int, int]]] = [[] for _ in range(n)]
    for r in range(n):
        for c, l in F.L[r].items():
            Ucol[c].append((r, l))                         # U(c,r)=L(r,c)
    for row in range(n - 1, -1, -1):                       # (I+L') x = z
        pairs = [(l, y[k]) for k, l in sorted(Ucol[row]) if y[k] is not None]
        y[row] = tab.plus_minus_dot(y[row], pairs)
    x: List[Id] = [None] * n
    for k in range(n):
        x[F.p[k]] = y[k]                                   # inverse permutation
    return x
