"""Fixed-sparsity-pattern LDL' factorisation and triangular solves, as instructions.

Restates the recipe of ``lib/@csparse/sparsity_ldl.m:144-236`` (left-looking column
LDL' on the symmetrically permuted pattern, strictly diagonal D, no numerical
pivoting, symbolic fill) and of ``sparsity_mldivide_l1.m:66-102`` /
``sparsity_mldivide_d.m:60-74`` / ``sparsity_mldivide_u1.m:67-104`` (permute b,
forward substitution, divide by D, backward substitution, inverse permutation).

Ordering: the reference calls MATLAB's ``symamd(A,[1000,0])`` (sparsity_ldl.m:91),
which is not available offline; a greedy minimum-degree ordering on the elimination
graph is used instead.  Any symmetric ordering is admissible for the quasi-definite
KKT matrices assembled in ``ipm_builder`` -- only fill, not correctness, depends on it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .itable import InstructionTable, Id

SparseSym = List[Dict[int, int]]  # rows[i] = {j: instruction}  (full symmetric pattern)


def minimum_degree_ordering(pattern: List[set]) -> List[int]:
    """Greedy minimum degree on the elimination graph (exact degree update).

    ``pattern[i]`` = set of j!=i with A(i,j) structurally non-zero (symmetric).
    Ties are broken by the original index so that the ordering is deterministic.
    """
    import heapq

    n = len(pattern)
    adj = [set(s) - {i} for i, s in enumerate(pattern)]
    alive = [True] * n
    heap = [(len(adj[i]), i) for i in range(n)]
    heapq.heapify(heap)
    order: List[int] = []
    while heap:
        d, v = heapq.heappop(heap)
        if not alive[v] or d != len(adj[v]):
            continue
        alive[v] = False
        order.append(v)
        nb = adj[v]
        for a in nb:
            s = adj[a]
            s.discard(v)
            s |= nb
            s.discard(a)
        for a in nb:
            heapq.heappush(heap, (len(adj[a]), a))
        adj[v] = set()
    return order


class LDLFactor:
    def __init__(self, n: int, p: List[int], L: List[Dict[int, int]], D: List[int]):
        self.n = n
        self.p = p          # permuted index k  <-  original index p[k]
        self.L = L          # L[row] = {col: instr}, col<row, in permuted indices
        self.D = D          # D[k] instr

    def nnz(self) -> int:
        return sum(len(r) for r in self.L) + self.n


def ldl(tab: InstructionTable, A: SparseSym, order: Optional[Sequence[int]] = None) -> LDLFactor:
    """Symbolic LDL' of the symmetric matrix whose LOWER triangle is taken from ``A``."""
    n = len(A)
    if order is None:
        pat = [set() for _ in range(n)]
        for i, row in enumerate(A):
            for j in row:
                if j != i:
                    pat[i].add(j)
                    pat[j].add(i)
        order = minimum_degree_ordering(pat)
    p = list(order)
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
