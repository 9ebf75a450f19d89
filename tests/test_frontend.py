"""Host-side generator logic (CPU only): symbolic derivatives, fixed-pattern LDL and solves,
dependency groups / lazy re-evaluation, tape lowering.  Patterns follow the reference's
examples/testing scripts: differential tests with seeded inputs and tolerance 1e3*eps
(testLinearAlgebra.m:57-59, testFactorizations.m:46-91, testLDL.m:85-93, numericalGradient.m)."""
import numpy as np
import pytest

from conftest import generated
from tenscalc_b200 import itable as it
from tenscalc_b200 import ldl as ldlmod
from tenscalc_b200 import tensor as T
from tenscalc_b200.differentiate import gradient, hessian
from oracle.pyeval import evaluate

EPS = np.finfo(float).eps


def _vals(var, arr):
    return dict(zip(var.flat(), np.asarray(arr, float).reshape(-1, order="F")))


def test_hash_consing_and_normalisation():
    T.clear()
    tab = T.table()
    a, b = T.Tvariable("a", 3), T.Tvariable("b", 3)
    e1 = (a * b + b * a).flat()
    e2 = (b * a + a * b).flat()
    assert e1 == e2                                   # operands sorted, duplicates merged
    assert (a - a).flat() == [None] * 3               # x-x is a structural zero
    n = len(tab)
    _ = a * b
    assert len(tab) == n                              # appendUniqueInstruction: nothing new
    assert T.Tvariable("c", 2).flat() != T.Tvariable("c", 2).flat()   # I_set never merged


def test_gradient_and_hessian_against_finite_differences():
    T.clear()
    tab = T.table()
    x = T.Tvariable("x", 4)
    p = T.Tvariable("p", 4)
    f = (T.norm2(x * p) + T.tsum(T.sqrt(x * x + 1.0), 0) + T.tsum(T.exp(x) / (p + 2.0), 0)
         + T.tsum(T.sin(x) * T.cos(p), 0) + (x[0] * x[1] * x[2]) + T.tsum(T.atan(x), 0) + T.tsum(T.log(x * x + 2.0), 0))
    u = x.flat()
    H, g = hessian(tab, f.e[()], u)
    rng = np.random.default_rng(0)
    x0, p0 = rng.normal(size=4), rng.uniform(.5, 1.5, size=4)

    def fval(xv):
        return evaluate(tab, [f.e[()]], {**_vals(x, xv), **_vals(p, p0)})[0][0]

    gv = np.array([evaluate(tab, [gi], {**_vals(x, x0), **_vals(p, p0)})[0][0] for gi in g])
    h = 1e-6
    gn = np.array([(fval(x0 + h * e) - fval(x0 - h * e)) / (2 * h) for e in np.eye(4)])
    assert np.abs(gv - gn).max() <= 1e-7 * max(1, np.abs(gn).max())
    Hv = np.zeros((4, 4))
    for i in range(4):
        for j, e in H[i].items():
            Hv[i, j] = evaluate(tab, [e], {**_vals(x, x0), **_vals(p, p0)})[0][0]

    def gval(xv):
        return np.array([evaluate(tab, [gi], {**_vals(x, xv), **_vals(p, p0)})[0][0] for gi in g])
    Hn = np.array([(gval(x0 + h * e) - gval(x0 - h * e)) / (2 * h) for e in np.eye(4)])
    assert np.abs(Hv - Hn).max() <= 1e-6 * max(1, np.abs(Hn).max())
    assert np.abs(Hv - Hv.T).max() <= 1e3 * EPS * np.abs(Hv).max()


def test_componentwise_user_function_derivatives():
    """componentwise(x,fun,cfun,derivative1,derivative2) (lib/@Tcalculus/Tcalculus.m:1836-1880): a user-defined
    tanh with its derivatives written in terms of itself -- gradient and Hessian by the chain rule against finite
    differences; constants fold through the Python form, structural zeros through fun(0); without a derivative the
    differentiator refuses"""
    import math
    T.clear()
    tab = T.table()

    def tanh_(z):
        return T.componentwise(z, math.tanh, "tanh(%s)", lambda w: 1.0 - T.sqr(tanh_(w)))

    x = T.Tvariable("x", 3)
    p = T.Tvariable("p", 3)
    f = T.tsum(tanh_(x * p) * x, 0) + T.tsum(T.sqr(tanh_(x)), 0)
    u = x.flat()
    H, g = hessian(tab, f.e[()], u)
    assert any(tab.types[i] == it.I_componentwise for i in tab.ancestors([f.e[()]]))
    rng = np.random.default_rng(3)
    x0, p0 = rng.normal(size=3), rng.uniform(.5, 1.5, size=3)
    vals = lambda xv: {**_vals(x, xv), **_vals(p, p0)}
    fval = lambda xv: evaluate(tab, [f.e[()]], vals(xv))[0][0]
    gval = lambda xv: np.array([evaluate(tab, [gi], vals(xv))[0][0] for gi in g])
    h = 1e-6
    gn = np.array([(fval(x0 + h * e) - fval(x0 - h * e)) / (2 * h) for e in np.eye(3)])
    assert np.abs(gval(x0) - gn).max() <= 1e-7 * max(1, np.abs(gn).max())
    Hv = np.zeros((3, 3))
    for i in range(3):
        for j, e in H[i].items():
            Hv[i, j] = evaluate(tab, [e], vals(x0))[0][0]
    Hn = np.array([(gval(x0 + h * e) - gval(x0 - h * e)) / (2 * h) for e in np.eye(3)])
    assert np.abs(Hv - Hn).max() <= 1e-6 * max(1, np.abs(Hn).max())
    # constants and structural zeros
    c = T.componentwise(T.Tconstant([0.5, 0.0]), math.cos, "cos(%s)")
    assert [tab.const_value(i) for i in c.flat()] == [math.cos(0.5), 1.0]
    assert T.componentwise(T.Tconstant([0.0]), math.tanh, "tanh(%s)").flat() == [None]
    # the C text is what the emitters print, operand at every %s (instructionsTableUTHash.c:1333-1345)
    y = T.componentwise(x, lambda v: v * abs(v), "%s*fabs(%s)")
    assert it.componentwise_text(tab.params[y.flat()[0]], "m[7]") == "m[7]*fabs(m[7])"
    with pytest.raises(NotImplementedError, match="no derivative"):
        gradient(tab, T.tsum(y, 0).e[()], u)
    with pytest.raises(ValueError):
        T.componentwise(x, math.sin, "sin(x)")


@pytest.mark.parametrize("n,density,seed", [(5, .75, 0), (6, .75, 1), (12, .3, 2), (20, .2, 3)])
def test_ldl_identity_and_solve(n, density, seed):
    """testFactorizations.m:46-91 (L*D*L'=A on random 75%-dense matrices) and testLDL.m
    (ldl(A)\\B vs A\\B for n=6, m=2), tolerance 1e3*eps relative"""
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(n, n)) * (rng.random((n, n)) < density)
    A0 = M + M.T + np.diag(rng.uniform(2, 3, n) * np.where(np.arange(n) % 2, -1, 1) * n)   # quasi-definite-like
    T.clear()
    tab = T.table()
    Av = T.Tvariable("A", [n, n])
    rows = [{j: Av.e[i, j] for j in range(n) if A0[i, j] != 0} for i in range(n)]
    fac = ldlmod.ldl(tab, rows)
    vals = _vals(Av, np.tril(A0) + np.tril(A0, -1).T)
    L = np.eye(n)
    D = np.zeros(n)
    for r in range(n):
        for c, e in fac.L[r].items():
            L[r, c] = evaluate(tab, [e], vals)[0][0]
        D[r] = evaluate(tab, [fac.D[r]], vals)[0][0]
    Pm = np.eye(n)[fac.p]                                  # permuted = P A P'
    assert np.abs(L @ np.diag(D) @ L.T - Pm @ A0 @ Pm.T).max() <= 1e3 * EPS * np.abs(A0).max() * n
    for m in range(2):
        bv = T.Tvariable(f"b{m}", n)
        b0 = rng.normal(size=n)
        xs = ldlmod.solve(tab, fac, bv.flat())
        xv = np.array([v[0] for v in evaluate(tab, xs, {**vals, **_vals(bv, b0)})])
        xref = np.linalg.solve(A0, b0)
        assert np.abs(xv - xref).max() <= 1e4 * EPS * np.linalg.cond(A0) * max(1, np.abs(xref).max())


def test_ldl_structural_zero_pivot_is_an_error():
    """sparsity_ldl.m:187-189: "ldl needs pivoting" is a generation-time error"""
    T.clear()
    tab = T.table()
    a = T.Tvariable("a", 1).flat()[0]
    with pytest.raises(ValueError):
        ldlmod.ldl(tab, [{1: a}, {0: a}], order=[0, 1])


def test_default_ordering_reduces_fill():
    gen, _ = generated("mpc_dcmotor")
    assert gen.info["kkt"] == 323
    assert gen.info["nnzL"] < 323 * 6, "banded MPC KKT must factor with O(n) fill"


def test_orderings_are_permutations_and_symamd_is_the_committed_one():
    """sparsity_ldl.m:91 orders with symamd(.,[1000,0]); the permutation COLAMD returns for every
    BASELINE pattern is committed under tenscalc_b200/orderings/ (the GPU box has no SuiteSparse
    sources) and, where libcolamd.so is built, must agree with a fresh call"""
    import random
    from tenscalc_b200 import ordering
    rng = random.Random(0)
    n = 60
    pat = [set() for _ in range(n)]
    for i in range(n - 1):                       # chain + a few random chords
        pat[i].add(i + 1); pat[i + 1].add(i)
    for _ in range(40):
        i, j = rng.randrange(n), rng.randrange(n)
        if i != j:
            pat[i].add(j); pat[j].add(i)
    stats = {}
    for name in ("mindegree", "minfill") + (("symamd",) if ordering.symamd_available() else ()):
        perm = ordering.METHODS[name](pat) if name != "symamd" else ordering.symamd(pat, write_cache=False)
        assert sorted(perm) == list(range(n))
        stats[name] = ordering.symbolic_stats(pat, perm)
    natural = ordering.symbolic_stats(pat, list(range(n)))
    assert all(s["nnzL"] <= natural["nnzL"] for s in stats.values())
    gen, _ = generated("mpc_quadcopter")
    # symamd on the quadcopter KKT: 3.3x fewer multiply-adds than round 1's greedy minimum degree
    assert gen.info["nnzL"] < 9000


def test_symamd_cache_matches_colamd():
    from tenscalc_b200 import ordering
    import json, os
    if not ordering.symamd_available():
        pytest.skip("libcolamd.so not built (no SuiteSparse tree on this box)")
    from tenscalc_b200 import ipm_builder
    seen = []
    orig = ipm_builder.ldlmod.ldl

    def spy(tab, A, order=None):
        seen.append(ordering.pattern_of(A))
        return orig(tab, A, order)
    ipm_builder.ldlmod.ldl = spy
    try:
        from tenscalc_b200 import problems
        problems.BUILDERS["mpc_dcmotor"]()
    finally:
        ipm_builder.ldlmod.ldl = orig
    assert seen
    for pat in seen:
        path = os.path.join(ordering.CACHE_DIR, ordering.pattern_key(pat) + ".json")
        assert os.path.exists(path), "ordering cache entry missing: regenerate the problem once and commit it"
        assert json.load(open(path))["perm"] == ordering._symamd_colamd(pat)


def test_dependency_groups_lazy_semantics():
    """a set invalidates exactly the groups that depend on it; a get runs exactly its stale
    parents, in topological order (writeCfunctionpergroup.m:566-571,653-657)"""
    gen, _ = generated("sls")
    pr = gen.problem
    tab = pr.tab
    for g, members in enumerate(pr.groups):          # groups are topologically ordered
        for i in members:
            for o in tab.operands[i]:
                go = pr.group_of.get(o)
                assert go is None or go <= g
    mu_children = set(pr.sets["setMu__"].children)
    ap_children = set(pr.sets["setAlphaPrimal__"].children)
    par_children = set(pr.sets["setP_A"].children)
    grad_parents = set(pr.gets["getNorminf_Grad__"].parents)
    minF_parents = set(pr.gets["getMinF_s__"].parents)
    assert not (mu_children & grad_parents), "the gradient of the Lagrangian does not depend on mu"
    assert ap_children & minF_parents, "the line-search value depends on alphaPrimal"
    assert ap_children <= set(pr.copies["updatePrimalDual__"].parents) | minF_parents
    # parameter-only groups (A'A-like subexpressions) are not invalidated by the primal update
    upd_children = set(pr.copies["updatePrimalDual__"].children)
    assert par_children - upd_children, "parameter-only groups must exist (evaluated once per setP_*)"


def test_memory_map_never_aliases_live_values():
    gen, _ = generated("mpc_dcmotor")
    pr = gen.problem
    tab, slot = pr.tab, pr.slot
    owner = {}
    pinned = set()
    for g in pr.gets.values():
        pinned.update(x for s in g.sources for x in s if x is not None)
    for c in pr.copies.values():
        pinned.update(c.src)
    for members in pr.groups:
        for i in members:
            for o in tab.operands[i]:
                if tab.types[o] in (it.I_load, it.I_2) and slot[o] < 0:
                    continue
                assert owner.get(int(slot[o]), o) == o, "operand slot was overwritten before its last use"
            assert all(slot[i] != slot[o] for o in tab.operands[i]), "result aliases an operand"
            owner[int(slot[i])] = i
    for i in pinned:
        assert owner.get(int(slot[i]), i) == i or tab.types[i] == it.I_set


def test_tape_structure():
    """groups start on bundle boundaries, every instruction ends in exactly one F_END cell, late
    operands point into the ring / constant pool, look-ahead operands are never produced inside
    the look-ahead window"""
    from tenscalc_b200 import tape as tp
    gen, _ = generated("mpc_dcmotor")
    pr = gen.problem
    L = tp.lower(gen)
    cells, gs = L["tape"], L["group_start"]
    assert len(cells) % tp.BUNDLE == 0 and all(int(x) % tp.BUNDLE == 0 for x in gs)
    assert tp.RING > 2 * tp.BUNDLE
    for g, members in enumerate(pr.groups):
        grp = cells[int(gs[g]):int(gs[g + 1])]
        ends = [k for k, c in enumerate(grp) if int(c[0]) & tp.F_END]
        assert len(ends) == len(members)
        produced_at = {}                      # scratchbook offset -> cell index of its store
        n = 0
        for k, c in enumerate(grp):
            w0 = int(c[0])
            for w, late in ((int(c[1]), w0 & tp.F_ALATE), (int(c[2]), w0 & tp.F_BLATE)):
                if late:
                    assert w < tp.RING_BYTES + 8 * tp.MAX_CONSTS
                    continue
                assert w % tp.TILE_BYTES == 0 and w // tp.TILE_BYTES < pr.n_slots
                if w in produced_at:          # fetched when the previous bundle starts executing
                    fetch_time = (k // tp.BUNDLE - 1) * tp.BUNDLE
                    assert produced_at[w] < fetch_time, "look-ahead operand produced inside the window"
            if w0 & tp.F_END:
                assert ((w0 >> 16) & 0xFF) == n % tp.RING
                if w0 & tp.F_STORE:
                    produced_at[int(c[3])] = k
                n += 1


def test_reference_rejected_options_raise():
    """ipmPD_CStimesLambda.m:31-48 raises for these; so does the generator here"""
    from tenscalc_b200.optimize import cmex2optimizeCS
    T.clear()
    x = T.Tvariable("x", 2)
    for kw in (dict(smallerNewtonMatrix=True), dict(useLDL=False), dict(useUmfpack=True)):
        with pytest.raises(NotImplementedError):
            cmex2optimizeCS(classname="t", objective=T.norm2(x), optimizationVariables=[x], **kw)
    with pytest.raises(ValueError):
        cmex2optimizeCS(classname="t", objective=x, optimizationVariables=[x])


def test_problem_sizes_match_survey_table():
    sizes = dict(sls=(8, 0, 16), mpc_dcmotor=(89, 60, 174), l1l2=(996, 0, 796), mpc_quadcopter=(700, 406, 300))
    for name, (nU, nG, nF) in sizes.items():
        gen, _ = generated(name)
        d = gen.defines
        assert (d["nU"], d["nG"], d["nF"]) == (nU, nG, nF), name


def _emulate_cells(L, g, m, consts, ring_size):
    """numpy model of csrc/tc_vm.cuh::tc_exec_group (one lane per batch entry): fetches the
    scratchbook operands of bundle b+1 BEFORE executing bundle b, exactly like the kernel, so a
    lowering that let a look-ahead operand be produced inside the window would be caught here"""
    from tenscalc_b200 import tape as tp
    from oracle.pyeval import _UNARY
    cells = L["tape"]
    lo, hi = int(L["group_start"][g]), int(L["group_start"][g + 1])
    B = m.shape[1]
    ring = np.zeros((ring_size, B))
    acc = np.zeros(B)
    p = np.zeros(B)
    U = tp.BUNDLE

    def fetch(b):
        out = []
        for j in range(U):
            w0, wa, wb, _ = [int(v) for v in cells[lo + b * U + j]]
            assert wa % 8 in (0, 1) and wb % 8 in (0, 1)
            out.append([None if w0 & tp.F_ALATE else m[wa // tp.TILE_BYTES].copy(),
                        None if w0 & tp.F_BLATE else m[wb // tp.TILE_BYTES].copy()])
        return out

    def late(w):
        if w & 1:
            assert (w - 1) % (tp.BLOCK * 8) == 0 and w < tp.RING_BYTES
            return ring[(w - 1) // (tp.BLOCK * 8)].copy()
        assert w >= tp.RING_BYTES
        return np.full(B, consts[(w - tp.RING_BYTES) // 8])
    nb = (hi - lo) // U
    nxt = fetch(0)
    with np.errstate(all="ignore"):
        for b in range(nb):
            cur = nxt
            if b + 1 < nb:
                nxt = fetch(b + 1)
            for j in range(U):
                w0, wa, wb, wr = [int(v) for v in cells[lo + b * U + j]]
                op = w0 & 0x7F
                x = late(wa) if w0 & tp.F_ALATE else cur[j][0]
                y = late(wb) if w0 & tp.F_BLATE else cur[j][1]
                sx = -x if w0 & tp.F_NEGA else x
                sy = -y if w0 & tp.F_NEGB else y
                init = bool(w0 & tp.F_INIT)
                if op == tp.OP_MADD:
                    t = x * y
                    if w0 & tp.F_NEGA:
                        t = -t
                    acc = t if w0 & tp.F_FIRST else acc + t
                elif op == tp.OP_NOP: pass
                elif op == tp.OP_SUM1F: acc = sx
                elif op == tp.OP_SUM2F: acc = sx + sy
                elif op == tp.OP_SUM1: acc = acc + sx
                elif op == tp.OP_SUM2: acc = (acc + sx) + sy
                elif op == tp.OP_PSET: p = x * y
                elif op == tp.OP_PMUL1: p = p * x
                elif op == tp.OP_PMUL2: p = (p * x) * y
                elif op == tp.OP_P1SET: acc = p * x
                elif op == tp.OP_P1ADD: acc = acc + p * x
                elif op == tp.OP_P2SET: acc = (p * x) * y
                elif op == tp.OP_P2ADD: acc = acc + (p * x) * y
                elif op == tp.OP_RSUB: acc = x - acc
                elif op == tp.OP_RSUBDIV: acc = (x - acc) / y
                elif op == tp.OP_NEGDIV: acc = (-acc) / x
                elif op == tp.OP_DIV: acc = x / y
                elif op == tp.OP_SQ1F: acc = x * x
                elif op == tp.OP_SQ2F: acc = x * x + y * y
                elif op == tp.OP_SQ1: acc = acc + x * x
                elif op == tp.OP_SQ2: acc = (acc + x * x) + y * y
                elif op == tp.OP_ABS1F: acc = np.abs(x)
                elif op == tp.OP_ABS2F: acc = np.abs(x) + np.abs(y)
                elif op == tp.OP_ABS1: acc = acc + np.abs(x)
                elif op == tp.OP_ABS2: acc = (acc + np.abs(x)) + np.abs(y)
                elif op in (tp.OP_MINF1, tp.OP_MAXF1): acc = x
                elif op == tp.OP_AMAXF1: acc = np.abs(x)
                elif op in (tp.OP_MINF2, tp.OP_MIN1, tp.OP_MIN2, tp.OP_MAXF2, tp.OP_MAX1, tp.OP_MAX2,
                            tp.OP_AMAXF2, tp.OP_AMAX1, tp.OP_AMAX2):
                    is_min = op in (tp.OP_MINF2, tp.OP_MIN1, tp.OP_MIN2)
                    f = np.abs if op in (tp.OP_AMAXF2, tp.OP_AMAX1, tp.OP_AMAX2) else (lambda v: v)
                    vals = [x] if op in (tp.OP_MIN1, tp.OP_MAX1, tp.OP_AMAX1) else [x, y]
                    if op in (tp.OP_MINF2, tp.OP_MAXF2, tp.OP_AMAXF2):
                        acc, vals = f(x), [y]
                    elif init:
                        acc = np.zeros(B)
                    for v in vals:
                        take = (acc > f(v)) if is_min else (acc < f(v))
                        acc = np.where(take, f(v), acc)
                elif op == tp.OP_CLP:
                    if init:
                        acc = np.full(B, np.finfo(float).max)
                    q = (-x) / y
                    acc = np.where((y < 0) & (acc > q), q, acc)
                elif op >= tp.OP_UNARY:
                    acc = _UNARY[op - tp.OP_UNARY](x)
                else:
                    raise AssertionError(op)
                if w0 & tp.F_END:
                    if w0 & tp.F_NEG:
                        acc = -acc
                    ring[(w0 >> 16) & 0xFF] = acc
                    if w0 & tp.F_STORE:
                        m[wr // tp.TILE_BYTES] = acc


@pytest.mark.parametrize("name", ["sls", "mpc_dcmotor", "mpc_quadcopter", "isa_probe", "qsaddle_latent"])
def test_cell_tape_equals_instruction_semantics(name):
    """every group: numpy model of the kernel's cell machine (with its look-ahead operand fetch)
    vs direct evaluation of the reference instruction forms (oracle/pyeval.py), bit for bit"""
    from tenscalc_b200 import tape as tp
    from oracle.pyeval import eval_instruction
    gen, _ = generated(name)
    pr = gen.problem
    tab, slot = pr.tab, pr.slot
    L = tp.lower(gen)
    B = 3
    rng = np.random.default_rng(5)
    groups = range(len(pr.groups)) if name != "mpc_quadcopter" else \
        sorted(range(len(pr.groups)), key=lambda g: -len(pr.groups[g]))[:4]
    external = set()
    for gdecl in pr.gets.values():
        external.update(x for s in gdecl.sources for x in s if x is not None)
    for c in pr.copies.values():
        external.update(c.src)
    for i, gi in pr.group_of.items():
        external.update(o for o in tab.operands[i] if pr.group_of.get(o, -1) != gi)
    for g in groups:
        members = pr.groups[g]
        m0 = rng.uniform(.5, 1.5, size=(pr.n_slots, B)) * rng.choice([-1, 1], size=(pr.n_slots, B))
        # (a) instruction semantics, values keyed by instruction id; operands outside the group
        # come from the scratchbook
        v = {}
        mset = set(members)
        with np.errstate(all="ignore"):
            for i in members:
                for o in tab.operands[i]:
                    if o not in v and o not in mset and tab.types[o] not in (it.I_load, it.I_2):
                        v[o] = m0[slot[o]].copy()
                    elif o not in v and tab.types[o] in (it.I_load, it.I_2):
                        v[o] = eval_instruction(tab, o, v, B)
                v[i] = eval_instruction(tab, i, v, B)
        # (b) cell machine
        m = m0.copy()
        _emulate_cells(L, g, m, L["consts"], tp.RING)
        for i in members:
            if i in external:
                a, b = m[slot[i]], v[i]
                assert np.array_equal(a, b, equal_nan=True), (name, g, i, it.OPNAMES[tab.types[i]])
