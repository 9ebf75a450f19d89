"""Symbolic construction of everything the min-max driver ``ipmPDminmax_CSsolver`` gets/sets/copies.

Restates ``lib/ipmPDminmax_CS.m:1-344`` (the builder behind ``cmex2minmaxCS`` / ``class2minmaxCS``):

    min_u max_d  f(u,d)   s.t.  Fu(u) >= 0, Gu(u) = 0   (minimiser;  must not depend on d, :86-97)
                                Fd(u,d) >= 0, Gd(u,d) = 0 (maximiser)

Lagrangian Lf = f + nuU*Gu + nuD*Gd - lambdaU*Fu + lambdaD*Fd (:158-171), z=[u;d], and the symmetric Newton matrix

    WW = [ Lf_uu + eU*I   Lf_ud          Lf_zn     Lf_zl                     ]      b_s = [ -Lf_z ; -G ;
         [ Lf_du          Lf_dd - eD*I                                       ]              +Fu - mu./lambdaU ;
         [ [Gu;Gd]_z      -eEq*I         0                                   ]              -Fd + mu./lambdaD ]
         [ [-Fu;Fd]_z     0              diag([-Fu./lambdaU ; Fd./lambdaD])  ]

(:214-243) with THREE regularisers (minimiser, maximiser, equalities), factored by the fixed-pattern LDL';
``getHessUinertia__`` counts the signs of its D, ``getHessDinertia__`` those of the LDL' of the maximiser's own
KKT block ``HessD`` (:255-272) -- the driver wants nU+nGd+nFd positive pivots in the first and nD negative ones in
the second.
"""
from __future__ import annotations

from typing import Dict, List

from . import itable as it
from . import ldl as ldlmod
from .csparse import CSparse
from .differentiate import Differentiator
from .itable import InstructionTable, Id


def build_ipm_minmax(code: CSparse, u: List[int], d: List[int], f: Id, Fu: List[Id], Fd: List[Id], Gu: List[Id],
                     Gd: List[Id], lamU: List[int], lamD: List[int], nuU: List[int], nuD: List[int], *,
                     scaleInequalities: bool = True, scaleCost: float = 0.0, addEye2Hessian: bool = True,
                     ordering=None) -> Dict[str, object]:
    tab: InstructionTable = code.tab
    nU, nD = len(u), len(d)
    nZ = nU + nD
    nGu, nGd, nFu, nFd = len(Gu), len(Gd), len(Fu), len(Fd)
    nG, nF = nGu + nGd, nFu + nFd
    if any(e is None for e in Fu + Fd + Gu + Gd):
        raise ValueError("constraints that do not involve any variable or parameter")
    if not addEye2Hessian:
        raise NotImplementedError("addEye2Hessian=false: LDL without pivoting needs a non-zero diagonal")
    one = tab.const(1.0)
    zero = tab.append_unique(it.I_load, (0.0,))
    big = tab.append_unique(it.I_load, (1.7976931348623157e308,))

    if scaleInequalities and nF > 0:                                                         # :44-62
        dst, src = [], []
        if nFu:
            sU = [tab.new_set(f"scale4IneqU__[{k}]") for k in range(nFu)]
            dst += sU
            src += [tab.unary(it.I_abs, tab.div(one, e)) for e in Fu]
            Fu = [tab.mul([s, e]) for s, e in zip(sU, Fu)]
        if nFd:
            sD = [tab.new_set(f"scale4IneqD__[{k}]") for k in range(nFd)]
            dst += sD
            src += [tab.unary(it.I_abs, tab.div(one, e)) for e in Fd]
            Fd = [tab.mul([s, e]) for s, e in zip(sD, Fd)]
        code.declare_copy("scaleIneq__", dst, src)
    if scaleCost > 0:                                                                        # :63-70
        s4c = tab.new_set("scale4Cost__")
        code.declare_copy("scaleCost__", [s4c], [tab.unary(it.I_abs, tab.div(tab.const(scaleCost), f))])
        f = tab.mul([s4c, f])
        code.declare_get("getScale4Cost__", [[s4c]])
    code.declare_get("getf__", [[f]])                                                        # :74

    z = list(u) + list(d)
    nu = list(nuU) + list(nuD)
    lam = list(lamU) + list(lamD)
    G = list(Gu) + list(Gd)
    F = list(Fu) + list(Fd)
    w = z + nu + lam
    pos = {v: k for k, v in enumerate(w)}
    D = Differentiator(tab, w)

    def grad_row(e: Id) -> Dict[int, int]:
        return {} if e is None else {pos[v]: dd for v, dd in D.grad(e).items()}

    for e in Gu + Fu:                                                                        # :86-97
        if any(nU <= c < nZ for c in grad_row(e)):
            raise ValueError("cmex2minmaxCS: the minimiser's constraints cannot depend on the maximiser's variables")

    gap = tab.dot(zip(lam, F)) if nF else None
    mu = None
    if nF > 0:                                                                               # :113-118
        mu = tab.new_set("mu__")
        code.declare_set("setMu__", [mu])
        code.declare_get("getGapMinFMinLambda__", [[gap], [tab.minmax(F, True)], [tab.minmax(lam, True)]])
    if nG > 0:
        code.declare_get("getNorminf_G__", [[tab.max_abs(G)]])
        code.declare_copy("initDualEq__", nu, [one] * nG)                                    # :123-136
    if nF > 0:
        code.declare_copy("initDualIneq__", lam, [tab.div(mu, e) for e in F])                # :137-149

    Lf = tab.add([(1, f)] + ([(1, tab.dot(zip(nuU, Gu)))] if nGu else []) + ([(1, tab.dot(zip(nuD, Gd)))] if nGd else [])
                 + ([(-1, tab.dot(zip(lamU, Fu)))] if nFu else []) + ([(1, tab.dot(zip(lamD, Fd)))] if nFd else []))
    aP = tab.new_set("alphaPrimal__")
    aDe = tab.new_set("alphaDualEq__")
    aDi = tab.new_set("alphaDualIneq__")
    code.declare_set("setAlphaPrimal__", [aP])
    code.declare_set("setAlphaDualEq__", [aDe])
    code.declare_set("setAlphaDualIneq__", [aDi])
    eU = tab.new_set("addEye2HessianU__")                                                    # :181-192
    eD = tab.new_set("addEye2HessianD__")
    eEq = tab.new_set("addEye2HessianEq__")
    code.declare_set("setAddEye2HessianU__", [eU])
    code.declare_set("setAddEye2HessianD__", [eD])
    code.declare_set("setAddEye2HessianEq__", [eEq])

    Lf_z_row = grad_row(Lf)
    Lf_z: List[Id] = [Lf_z_row.get(k) for k in range(nZ)]
    if any(e is None for e in Lf_z):
        raise ValueError("an optimisation variable does not enter the Lagrangian")
    code.declare_get("getNorminf_Grad__", [[tab.max_abs(Lf_z)]])                             # :200
    rows = [grad_row(e) for e in Lf_z]                  # d Lf_z / d [z; nu; lambda]

    n = nZ + nG + nF
    WW: List[Dict[int, int]] = [dict() for _ in range(n)]

    def put(M, r, c, v):
        if v is not None:
            M[r][c] = v

    for i in range(nZ):
        for c, v in rows[i].items():
            put(WW, i, c, v)
        reg = eU if i < nU else tab.neg(eD)
        WW[i][i] = tab.add([(1, rows[i].get(i)), (1, reg)])
    for k, e in enumerate(G):                                                                # [G_z, -eEq I, 0]
        r = nZ + k
        for c, v in grad_row(e).items():
            if c < nZ:
                put(WW, r, c, v)
        WW[r][r] = tab.neg(eEq)
    F_z = [{c: v for c, v in grad_row(e).items() if c < nZ} for e in F]
    for k, e in enumerate(F):                                                                # [[-Fu;Fd]_z, 0, diag([-Fu./lamU; Fd./lamD])]
        r = nZ + nG + k
        sgn = -1 if k < nFu else 1
        for c, v in F_z[k].items():
            put(WW, r, c, tab.neg(v) if sgn < 0 else v)
        q = tab.div(e, lam[k])
        WW[r][r] = tab.neg(q) if sgn < 0 else q
    # the lower triangle defines the symmetric matrix (sparsity_ldl.m: "LOWER ONES USED")
    for r in range(n):
        for c in list(WW[r]):
            if c > r:
                if r in WW[c]:
                    WW[r][c] = WW[c][r]
                else:
                    WW[c][r] = WW[r][c]
    b: List[Id] = [tab.neg(e) for e in Lf_z] + [tab.neg(e) for e in G]
    for k, e in enumerate(F):
        if k < nFu:
            b.append(tab.add([(1, e), (-1, tab.div(mu, lam[k]))]))
        else:
            b.append(tab.add([(-1, e), (1, tab.div(mu, lam[k]))]))
    fac = ldlmod.ldl(tab, WW, ordering)

    def inertia(D_: List[int]):
        npos = tab.add((1, tab.unary(it.I_heaviside, x)) for x in D_)
        nneg = tab.add((1, tab.unary(it.I_heaviside, tab.neg(x))) for x in D_)
        return [[zero if npos is None else npos], [zero if nneg is None else nneg]]

    # HessD: the maximiser's own KKT block (:255-264)
    idxD = list(range(nU, nZ)) + [nZ + nGu + k for k in range(nGd)] + [nZ + nG + nFu + k for k in range(nFd)]
    sel = {g: k for k, g in enumerate(idxD)}
    HessD: List[Dict[int, int]] = [dict() for _ in idxD]
    for g, k in sel.items():
        for c, v in WW[g].items():
            if c in sel:
                HessD[k][sel[c]] = v
    facD = ldlmod.ldl(tab, HessD, ordering) if idxD else None
    code.declare_get("getHessDinertia__", inertia(facD.D) if facD else [[zero], [zero]])     # :266-268
    code.declare_get("getHessUinertia__", inertia(fac.D))                                    # :270-272

    dx = ldlmod.solve(tab, fac, b)
    resid = [tab.add([(1, tab.dot((m, dx[c]) for c, m in sorted(WW[i].items()) if dx[c] is not None)), (-1, b[i])])
             for i in range(n)]
    derr = tab.max_abs(resid)
    code.declare_get("getDirectionError__", [[zero if derr is None else derr]])              # :277
    dZ, dNu, dLam = dx[:nZ], dx[nZ:nZ + nG], dx[nZ + nG:]

    def stepped(vars_, dv, alpha):
        return [tab.add([(1, a), (1, tab.mul([alpha, q]))]) for a, q in zip(vars_, dv)]

    newZ = stepped(z, dZ, aP)
    newNu = stepped(nu, dNu, aDe)
    newLam: List[Id] = []
    if nF > 0:
        FzdZ = [tab.dot((m, dZ[c]) for c, m in sorted(F_z[k].items()) if dZ[c] is not None) for k in range(nF)]
        mAP = tab.clp(list(zip(F, FzdZ)))
        mAD = tab.clp(list(zip(lam, dLam)))
        newLam = stepped(lam, dLam, aDi)
        newF = tab.substitute(F, dict(zip(z, newZ)))
        code.declare_get("getMaxAlphas_s__", [[big if mAP is None else mAP], [big if mAD is None else mAD]])
        code.declare_get("getMinF_s__", [[tab.minmax(newF, True)]])
    for e in newZ + newNu + newLam:
        if e is None:
            raise ValueError("structurally zero update")
    code.declare_copy("updatePrimalDual__", z + nu + lam, newZ + newNu + newLam)             # :318-338
    return dict(factor=fac, dx=dx, WW=WW, b=b, F=F, f=f, mu=mu, nZ=nZ, nU=nU, nD=nD, nG=nG, nF=nF, nGd=nGd, nFd=nFd)
