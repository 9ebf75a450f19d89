"""Fill-reducing symmetric orderings for the fixed-pattern LDL' of the KKT matrix.

The reference orders with MATLAB's ``p=symamd((A+A')/2,[1000,0])`` (``lib/@csparse/sparsity_ldl.m:91``),
i.e. COLAMD's ``symamd_l`` with ``knobs[COLAMD_DENSE_ROW]=1000``
(``SuiteSparse/COLAMD/MATLAB/symamdmex.c:86-97,151``; algorithm
``SuiteSparse/COLAMD/Source/colamd.c:1115``).  The ordering decides the whole instruction list of the
factorisation (fill, flops, rounding sequence), so ``"symamd"`` is the default here too:

* ``symamd``   the permutation COLAMD returns for the pattern.  Resolution order: (1) the committed
  cache ``tenscalc_b200/orderings/<sha1 of the pattern>.json`` -- one small file per KKT pattern,
  written the first time a pattern is ordered, so that problem generation is reproducible on a box
  without the SuiteSparse sources (the GPU box); (2) ``libcolamd.so`` compiled by
  ``build.build_colamd`` from the COLAMD sources of a SuiteSparse tree (``TENSCALC_SUITESPARSE``,
  default the reference's vendored submodule) -- the same third-party code MATLAB links;
  (3) neither: ``"minfill"`` with a warning.
* ``minfill``  greedy minimum local fill (ties: smaller degree, then smaller index); own
  implementation, no dependency.  On the BASELINE patterns it gives 5-15 % fewer multiply-adds
  than symamd.
* ``mindegree`` greedy exact minimum degree (round 1's ordering; 3.3x more multiply-adds than symamd
  on mpc_quadcopter -- kept for comparison only).

Patterns are ``List[set]``: ``pattern[i]`` = {j != i : A(i,j) or A(j,i) structurally non-zero}.
"""
from __future__ import annotations

import ctypes
import hashlib
import heapq
import json
import os
import warnings
from typing import Callable, Dict, List, Sequence, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE_DIR = os.path.join(HERE, "orderings")
COLAMD_LIB = os.path.join(HERE, "_gen", "libcolamd.so")

Pattern = List[set]


def pattern_of(A: Sequence[Dict[int, int]]) -> Pattern:
    """symmetrised off-diagonal pattern of a matrix given by rows ``A[i] = {j: instruction}``"""
    n = len(A)
    pat: Pattern = [set() for _ in range(n)]
    for i, row in enumerate(A):
        for j in row:
            if j != i:
                pat[i].add(j)
                pat[j].add(i)
    return pat


def _csc(pat: Pattern):
    """sparse((A+A')/2) as MATLAB stores it: CSC, sorted row indices, diagonal included"""
    n = len(pat)
    cols = [sorted(pat[j] | {j}) for j in range(n)]
    p = np.zeros(n + 1, dtype=np.int64)
    p[1:] = np.cumsum([len(c) for c in cols])
    Ai = np.fromiter((r for c in cols for r in c), dtype=np.int64, count=int(p[-1]))
    return Ai, p


def pattern_key(pat: Pattern) -> str:
    Ai, p = _csc(pat)
    h = hashlib.sha1()
    h.update(np.int64(len(pat)).tobytes())
    h.update(p.tobytes())
    h.update(Ai.tobytes())
    return h.hexdigest()


# ---------------------------------------------------------------------------------------------
def minimum_degree(pat: Pattern) -> List[int]:
    """Greedy minimum degree on the elimination graph (exact degree update, ties by index)."""
    n = len(pat)
    adj = [set(s) - {i} for i, s in enumerate(pat)]
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


def min_fill(pat: Pattern) -> List[int]:
    """Greedy minimum local fill: eliminate the vertex whose elimination adds the fewest edges
    (ties: smaller degree, then smaller index).  Exact fill counts, recomputed for every vertex
    within distance two of the eliminated one."""
    n = len(pat)
    adj = [set(s) - {i} for i, s in enumerate(pat)]
    alive = [True] * n

    def fill(v: int) -> int:
        nb = list(adj[v])
        f = 0
        for a in range(len(nb)):
            sa = adj[nb[a]]
            for b in range(a + 1, len(nb)):
                if nb[b] not in sa:
                    f += 1
        return f

    cost = [fill(v) for v in range(n)]
    heap = [(cost[v], len(adj[v]), v) for v in range(n)]
    heapq.heapify(heap)
    order: List[int] = []
    while heap:
        f, d, v = heapq.heappop(heap)
        if not alive[v] or f != cost[v] or d != len(adj[v]):
            continue
        alive[v] = False
        order.append(v)
        nb = adj[v]
        touched = set(nb)
        for a in nb:
            s = adj[a]
            s.discard(v)
            s |= nb
            s.discard(a)
        for a in nb:
            touched |= adj[a]
        for a in touched:
            if alive[a]:
                cost[a] = fill(a)
                heapq.heappush(heap, (cost[a], len(adj[a]), a))
        adj[v] = set()
    return order


# ---------------------------------------------------------------------------------------------
_colamd = None


def _load_colamd():
    global _colamd
    if _colamd is None and os.path.exists(COLAMD_LIB):
        lib = ctypes.CDLL(COLAMD_LIB)
        lib.symamd_l.restype = ctypes.c_int64
        lib.symamd_l.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 7
        _colamd = lib
    return _colamd


def symamd_available() -> bool:
    return _load_colamd() is not None


def _symamd_colamd(pat: Pattern) -> List[int]:
    """``symamd_l(n, A, p, perm, knobs, stats, calloc, free)`` with MATLAB's knobs for [1000,0]"""
    lib = _load_colamd()
    n = len(pat)
    Ai, p = _csc(pat)
    perm = np.zeros(n + 1, dtype=np.int64)
    knobs = (ctypes.c_double * 20)()
    lib.colamd_l_set_defaults(knobs)
    knobs[0] = 1000.0                                   # COLAMD_DENSE_ROW  <-  symamd(.,[1000,0])
    stats = (ctypes.c_int64 * 20)()
    libc = ctypes.CDLL(None)
    ok = lib.symamd_l(n, Ai.ctypes.data, p.ctypes.data, perm.ctypes.data,
                      ctypes.cast(knobs, ctypes.c_void_p), ctypes.cast(stats, ctypes.c_void_p),
                      ctypes.cast(libc.calloc, ctypes.c_void_p), ctypes.cast(libc.free, ctypes.c_void_p))
    if not ok:
        raise RuntimeError(f"symamd failed (status {stats[3]})")
    out = [int(x) for x in perm[:n]]
    if sorted(out) != list(range(n)):
        raise RuntimeError("symamd did not return a permutation")
    return out


def symamd(pat: Pattern, write_cache: bool = True) -> List[int]:
    key = pattern_key(pat)
    path = os.path.join(CACHE_DIR, key + ".json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if d.get("n") == len(pat) and sorted(d["perm"]) == list(range(len(pat))):
            return [int(x) for x in d["perm"]]
    if symamd_available():
        perm = _symamd_colamd(pat)
        if write_cache:
            os.makedirs(CACHE_DIR, exist_ok=True)
            with open(path, "w") as f:
                json.dump({"n": len(pat), "method": "symamd_l knobs[COLAMD_DENSE_ROW]=1000 (sparsity_ldl.m:91)",
                           "nnz_offdiag": int(sum(len(s) for s in pat)), "perm": perm}, f)
        return perm
    warnings.warn("symamd: pattern not in tenscalc_b200/orderings/ and libcolamd.so not built "
                  "(build.build_colamd); using the minimum-fill ordering instead", RuntimeWarning)
    return min_fill(pat)


METHODS: Dict[str, Callable[[Pattern], List[int]]] = {
    "symamd": symamd, "minfill": min_fill, "mindegree": minimum_degree,
}
DEFAULT = "symamd"


def order(A: Sequence[Dict[int, int]], method: Union[None, str, Callable] = None) -> List[int]:
    """elimination order of the symmetric matrix ``A`` (rows of {col: instruction})"""
    if callable(method):
        return list(method(A))
    name = method or DEFAULT
    if name not in METHODS:
        raise ValueError(f"unknown ordering {name!r}: one of {sorted(METHODS)} or a callable")
    return METHODS[name](pattern_of(A))


def symbolic_stats(pat: Pattern, perm: Sequence[int]) -> Dict[str, int]:
    """nnz(L) (strictly lower), multiply-adds of the left-looking LDL' and the widest column"""
    n = len(pat)
    inv = [0] * n
    for k, o in enumerate(perm):
        inv[o] = k
    colpat = [{inv[j] for j in pat[perm[k]] if inv[j] > k} for k in range(n)]
    children: List[List[int]] = [[] for _ in range(n)]
    rows: List[set] = [set() for _ in range(n)]
    for k in range(n):
        s = colpat[k]
        for c in children[k]:
            s |= {j for j in colpat[c] if j > k}
        if s:
            children[min(s)].append(k)
        for r in s:
            rows[r].add(k)
    terms = 0
    for c in range(n):
        Ic = rows[c]
        terms += len(Ic)
        for r in colpat[c]:
            terms += len(Ic & rows[r])
    return dict(nnzL=sum(len(s) for s in colpat), terms=terms, max_col=max((len(s) for s in colpat), default=0))
