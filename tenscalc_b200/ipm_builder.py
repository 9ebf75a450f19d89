"""Symbolic construction of everything the primal-dual IPM driver gets/sets/copies.

Restates ``lib/ipmPD_CStimesLambda.m:86-290`` (the builder ``cmex2optimizeCS`` always
selects, ``lib/cmex2optimizeCS.m:83-93``): scaling, Lagrangian, gradient and Hessian,
the "times lambda" symmetric KKT matrix

    WW = [ Lf_uu + eU*I      G_u'        -F_u' diag(lambda) ]
         [ G_u              -eEq*I        0                 ]
         [ -diag(lambda)F_u   0          -diag(F.*lambda)   ]

its fixed-pattern LDL' factorisation, the Newton direction, step bounds and the
multiplicative dual update.  Function names are the reference's ``*__`` names; the
driver (device: ``csrc/tc_driver.cuh``, oracle: the reference's own
``lib/ipmPD_CSsolver.c``) calls exactly these.

Host-side, offline.  Everything is expressed per scalar on the instruction table.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from . import itable as it
from . import ldl as ldlmod
from .csparse import CSparse
from .differentiate import Differentiator
from .itable import InstructionTable, Id


def build_ipm_timeslambda(code: CSparse, u: List[int], f: Id, F: List[Id], G: List[Id],
                          lam: List[int], nu: List[int], *, scaleInequalities: bool = True,
                          scaleCost: float = 0.0, addEye2Hessian: bool = True,
                          useInertia: bool = False, ordering=None) -> Dict[str, object]:
    tab: InstructionTable = code.tab
    nU, nG, nF = len(u), len(G), len(F)
    if any(x is None for x in F) or any(x is None for x in G):
        raise ValueError("constraints that do not involve any variable or parameter")
    out: Dict[str, object] = {}
    one = tab.const(1.0)

    # ---- scaling (ipmPD_CStimesLambda.m:86-103) -------------------------------------
    if nF > 0 and scaleInequalities:
        scale4Ineq = [tab.new_set(f"scale4Ineq__[{k}]") for k in range(nF)]
        code.declare_copy("scaleIneq__", scale4Ineq,
                          [tab.unary(it.I_abs, tab.div(one, Fk)) for Fk in F])
        F = [tab.mul([s, Fk]) for s, Fk in zip(scale4Ineq, F)]
    if scaleCost > 0:
        scale4Cost = tab.new_set("scale4Cost__")
        code.declare_copy("scaleCost__", [scale4Cost],
                          [tab.unary(it.I_abs, tab.div(tab.const(scaleCost), f))])
        f = tab.mul([scale4Cost, f])
        code.declare_get("getScale4Cost__", [[scale4Cost]])

    # ---- regularisation variables (:106-116) ----------------------------------------
    if not addEye2Hessian:
        raise NotImplementedError("addEye2Hessian=false: LDL without pivoting needs a non-zero diagonal")
    eU = tab.new_set("addEye2HessianU__")
    eEq = tab.new_set("addEye2HessianEq__")
    code.declare_set("setAddEye2HessianU__", [eU])
    code.declare_set("setAddEye2HessianEq__", [eEq])

    # ---- cost (:118-121) ---------------------------------------------------------------
    code.declare_get("getJ__", [[f]])

    # ---- Lagrangian and derivatives (:124-187) ---------------------------------------------
    D = Differentiator(tab, u)
    pos = {v: k for k, v in enumerate(u)}

    def grad_row(x: Id) -> Dict[int, int]:
        return {pos[v]: d for v, d in D.grad(x).items()}

    f_u = grad_row(f)
    F_u = [grad_row(Fk) for Fk in F]            # F_u[k] = {i: dF_k/du_i}
    G_u = [grad_row(Gk) for Gk in G]
    gap = tab.dot(zip(lam, F)) if nF else None
    Lf = tab.add([(1, f), (-1, gap)] + ([(1, tab.dot(zip(nu, G)))] if nG else []))
    Lf_u_d = grad_row(Lf)
    Lf_u: List[Id] = [Lf_u_d.get(i) for i in range(nU)]
    Lf_uu: List[Dict[int, int]] = [grad_row(g) for g in Lf_u]
    # the lower triangle defines the (symmetric) matrix, as in sparsity_ldl.m
    for i in range(nU):
        for j in list(Lf_uu[i]):
            if j > i:
                if i in Lf_uu[j]:
                    Lf_uu[i][j] = Lf_uu[j][i]
                else:
                    Lf_uu[j][i] = Lf_uu[i][j]
    # getGrad__ (:186) and initDualEq__ (:210) are declared by the reference but never
    # called by its C driver (ipmPD_CSsolver.c:233-234); they are left out so that they do
    # not split dependency groups.
    code.declare_get("getNorminf_Grad__", [[tab.max_abs(Lf_u)]])

    # ---- barrier variable, dual initialisation, gap (:189-221) --------------------------------
    if nF > 0:
        mu = tab.new_set("mu__")
        code.declare_set("setMu__", [mu])
        code.declare_copy("initDualIneq__", lam, [tab.div(mu, Fk) for Fk in F])
        code.declare_get("getGapMinFMinLambda__",
                         [[gap], [tab.minmax(F, True)], [tab.minmax(lam, True)]])
    else:
        mu = None
    if nG > 0:
        # least-squares initialisation: [I G_u';G_u -eEq*I] [w;nu] = [F_u'lambda-f_u;0]
        n0 = nU + nG
        WW0: List[Dict[int, int]] = [dict() for _ in range(n0)]
        for i in range(nU):
            WW0[i][i] = one
        for k in range(nG):
            r = nU + k
            for i, d in G_u[k].items():
                WW0[r][i] = d
                WW0[i][r] = d
            WW0[r][r] = tab.neg(eEq)
        b0: List[Id] = []
        for i in range(nU):
            Ftl = tab.dot((F_u[k][i], lam[k]) for k in range(nF) if i in F_u[k])
            b0.append(tab.add([(1, Ftl), (-1, f_u.get(i))]))
        b0 += [None] * nG
        fac0 = ldlmod.ldl(tab, WW0, ordering)
        wnu0 = ldlmod.solve(tab, fac0, b0)
        zero = tab.append_unique(it.I_load, (0.0,))
        code.declare_copy("initDualEqX__", nu, [zero if x is None else x for x in wnu0[nU:]])
        code.declare_get("getNorminf_G__", [[tab.max_abs(G)]])

    # ---- step sizes (:223-229) ---------------------------------------------------------------
    aP = tab.new_set("alphaPrimal__")
    aDe = tab.new_set("alphaDualEq__")
    aDi = tab.new_set("alphaDualIneq__")
    code.declare_set("setAlphaPrimal__", [aP])
    code.declare_set("setAlphaDualEq__", [aDe])
    code.declare_set("setAlphaDualIneq__", [aDi])

    # ---- KKT matrix (:234-240) ------------------------------------------------------------------
    n = nU + nG + nF
    WW: List[Dict[int, int]] = [dict() for _ in range(n)]
    WW11: List[Dict[int, int]] = [dict() for _ in range(nU)]
    for i in range(nU):
        for j, h in Lf_uu[i].items():
            WW11[i][j] = h
        WW11[i][i] = tab.add([(1, Lf_uu[i].get(i)), (1, eU)])
        WW[i].update(WW11[i])
    for k in range(nG):
        r = nU + k
        for i, d in G_u[k].items():
            WW[r][i] = d
            WW[i][r] = d
        WW[r][r] = tab.neg(eEq)
    for k in range(nF):
        r = nU + nG + k
        for i, d in F_u[k].items():
            x = tab.neg(tab.mul([lam[k], d]))
            WW[r][i] = x
            WW[i][r] = x
        WW[r][r] = tab.neg(tab.mul([F[k], lam[k]]))
    fac = ldlmod.ldl(tab, WW, ordering)
    out["factor"] = fac

    # inertia (:244-248)
    zero = tab.append_unique(it.I_load, (0.0,))
    if useInertia:   # only reachable from the driver when useInertia!=0 (ipmPD_CSsolver.c:408-419)
        npos = tab.add((1, tab.unary(it.I_heaviside, d)) for d in fac.D)
        nneg = tab.add((1, tab.unary(it.I_heaviside, tab.neg(d))) for d in fac.D)
        code.declare_get("getHessInertia__", [[zero if npos is None else npos], [zero if nneg is None else nneg]])

    # ---- Newton direction (:250-258) ----------------------------------------------------------------
    b: List[Id] = [tab.neg(g) for g in Lf_u]
    b += [tab.neg(Gk) for Gk in G]
    b += [tab.add([(1, tab.mul([lam[k], F[k]])), (-1, mu)]) for k in range(nF)]
    dx = ldlmod.solve(tab, fac, b)
    out["dx"] = dx
    resid = []
    for i in range(n):
        Wdx = tab.dot((w, dx[j]) for j, w in sorted(WW[i].items()) if dx[j] is not None)
        resid.append(tab.add([(1, Wdx), (-1, b[i])]))
    derr = tab.max_abs(resid)
    code.declare_get("getDirectionError__", [[zero if derr is None else derr]])

    # ---- updates (:260-281) ----------------------------------------------------------------------------
    dU, dNu, dLam = dx[:nU], dx[nU:nU + nG], dx[nU + nG:]
    newU = [tab.add([(1, u[i]), (1, tab.mul([aP, dU[i]]))]) for i in range(nU)]
    newNu = [tab.add([(1, nu[k]), (1, tab.mul([aDe, dNu[k]]))]) for k in range(nG)]
    W11dU = [tab.dot((w, dU[j]) for j, w in sorted(WW11[i].items()) if dU[j] is not None) for i in range(nU)]
    curvature = tab.dot((dU[i], W11dU[i]) for i in range(nU) if dU[i] is not None and W11dU[i] is not None)
    code.declare_get("getCurvature__", [[zero if curvature is None else curvature]])
    if nF > 0:
        newLam = [tab.mul([lam[k], tab.add([(1, one), (1, tab.mul([aDi, dLam[k]]))])]) for k in range(nF)]
        FudU = [tab.dot((d, dU[i]) for i, d in sorted(F_u[k].items()) if dU[i] is not None) for k in range(nF)]
        maxAlphaPrimal = tab.clp(list(zip(F, FudU)))
        maxAlphaDual = tab.clp([(one, d) for d in dLam])
        big = tab.append_unique(it.I_load, (1.7976931348623157e308,))
        code.declare_get("getMaxAlphas_s__", [[big if maxAlphaPrimal is None else maxAlphaPrimal],
                                              [big if maxAlphaDual is None else maxAlphaDual]])
        newF = tab.substitute(F, dict(zip(u, newU)))
        code.declare_get("getMinF_s__", [[tab.minmax(newF, True)]])
    else:
        newLam = []
    code.declare_copy("updatePrimalDual__", list(u) + list(nu) + list(lam), newU + newNu + newLam)

    out.update(dict(F=F, f=f, Lf_u=Lf_u, Lf_uu=Lf_uu, WW=WW, b=b, mu=mu, nU=nU, nG=nG, nF=nF))
    return out


def build_ipm_pd(code: CSparse, u: List[int], f: Id, F: List[Id], G: List[Id],
                 lam: List[int], nu: List[int], *, scaleInequalities: bool = True,
                 scaleCost: float = 0.0, addEye2Hessian: bool = True, useInertia: bool = False,
                 skipAffine: bool = True, smallerNewtonMatrix: bool = False, ordering=None) -> Dict[str, object]:
    """``lib/ipmPD_CS.m:74-466``: the builder behind ``ipmPD_CSsolver.c``'s full option set -- the
    plain (not "times lambda") KKT matrix with additive dual update, optionally the affine-scaling
    (Mehrotra) predictor (``skipAffine=false``: ``getMaxAlphas_a__``, ``getMinF_a__``, ``getRho__``,
    :313-315,422-425) and the reduced Newton system (``smallerNewtonMatrix=true`` :254-351, lambda
    eliminated).  ``cmex2optimizeCS`` as shipped hard-wires ``ipmPD_CStimesLambda`` (cmex2optimizeCS.m:83-93),
    so this builder is reached through the generator option ``algorithm="ipmPD_CS"``.

        large:  WW = [ Lf_uu + eU*I   G_u'      -F_u'          ]      b_s = [ -Lf_u ; -G ; F - mu./lambda (+ Mehrotra) ]
                     [ G_u           -eEq*I      0             ]
                     [ -F_u           0         -diag(F./lambda)]
        small:  WW = [ Lf_uu + F_u' diag(lambda./F) F_u + eU*I   G_u' ;  G_u   -eEq*I ]
    """
    tab: InstructionTable = code.tab
    nU, nG, nF = len(u), len(G), len(F)
    if any(x is None for x in F) or any(x is None for x in G):
        raise ValueError("constraints that do not involve any variable or parameter")
    if not addEye2Hessian:
        raise NotImplementedError("addEye2Hessian=false: LDL without pivoting needs a non-zero diagonal")
    out: Dict[str, object] = {}
    one = tab.const(1.0)
    zero = tab.append_unique(it.I_load, (0.0,))

    # ---- scaling (:74-95) ------------------------------------------------------------------------
    if nF > 0 and scaleInequalities:
        scale4Ineq = [tab.new_set(f"scale4Ineq__[{k}]") for k in range(nF)]
        code.declare_copy("scaleIneq__", scale4Ineq, [tab.unary(it.I_abs, tab.div(one, Fk)) for Fk in F])
        F = [tab.mul([s, Fk]) for s, Fk in zip(scale4Ineq, F)]
    if scaleCost > 0:
        scale4Cost = tab.new_set("scale4Cost__")
        code.declare_copy("scaleCost__", [scale4Cost], [tab.unary(it.I_abs, tab.div(tab.const(scaleCost), f))])
        f = tab.mul([scale4Cost, f])
        code.declare_get("getScale4Cost__", [[scale4Cost]])
    code.declare_get("getJ__", [[f]])                                                     # :99
    eU = tab.new_set("addEye2HessianU__")                                                 # :113-123
    eEq = tab.new_set("addEye2HessianEq__")
    code.declare_set("setAddEye2HessianU__", [eU])
    code.declare_set("setAddEye2HessianEq__", [eEq])

    # ---- Lagrangian and derivatives (:127-190) -----------------------------------------------------
    D = Differentiator(tab, u)
    pos = {v: k for k, v in enumerate(u)}

    def grad_row(x: Id) -> Dict[int, int]:
        return {pos[v]: d for v, d in D.grad(x).items()}

    f_u = grad_row(f)
    F_u = [grad_row(Fk) for Fk in F]
    G_u = [grad_row(Gk) for Gk in G]
    gap = tab.dot(zip(lam, F)) if nF else None
    Lf = tab.add([(1, f), (-1, gap)] + ([(1, tab.dot(zip(nu, G)))] if nG else []))
    Lf_u_d = grad_row(Lf)
    Lf_u: List[Id] = [Lf_u_d.get(i) for i in range(nU)]
    Lf_uu: List[Dict[int, int]] = [grad_row(g) for g in Lf_u]
    for i in range(nU):                       # the lower triangle defines the symmetric matrix (sparsity_ldl.m)
        for j in list(Lf_uu[i]):
            if j > i:
                if i in Lf_uu[j]:
                    Lf_uu[i][j] = Lf_uu[j][i]
                else:
                    Lf_uu[j][i] = Lf_uu[i][j]
    code.declare_get("getNorminf_Grad__", [[tab.max_abs(Lf_u)]])                          # :187

    # ---- barrier variable, dual initialisation (:193-228) ----------------------------------------------
    mu = None
    if nF > 0:
        mu = tab.new_set("mu__")
        code.declare_set("setMu__", [mu])
        code.declare_copy("initDualIneq__", lam, [tab.div(mu, Fk) for Fk in F])
        code.declare_get("getGapMinFMinLambda__", [[gap], [tab.minmax(F, True)], [tab.minmax(lam, True)]])
    if nG > 0:
        n0 = nU + nG
        WW0: List[Dict[int, int]] = [dict() for _ in range(n0)]
        for i in range(nU):
            WW0[i][i] = one
        for k in range(nG):
            r = nU + k
            for i, d in G_u[k].items():
                WW0[r][i] = d
                WW0[i][r] = d
            WW0[r][r] = tab.neg(eEq)
        b0: List[Id] = []
        for i in range(nU):
            Ftl = tab.dot((F_u[k][i], lam[k]) for k in range(nF) if i in F_u[k])
            b0.append(tab.add([(1, Ftl), (-1, f_u.get(i))]))
        b0 += [None] * nG
        fac0 = ldlmod.ldl(tab, WW0, ordering)
        wnu0 = ldlmod.solve(tab, fac0, b0)
        code.declare_copy("initDualEqX__", nu, [zero if x is None else x for x in wnu0[nU:]])
        code.declare_get("getNorminf_G__", [[tab.max_abs(G)]])
    aP = tab.new_set("alphaPrimal__")                                                     # :230-236
    aDe = tab.new_set("alphaDualEq__")
    aDi = tab.new_set("alphaDualIneq__")
    code.declare_set("setAlphaPrimal__", [aP])
    code.declare_set("setAlphaDualEq__", [aDe])
    code.declare_set("setAlphaDualIneq__", [aDi])

    def matvec_rows(rows: List[Dict[int, int]], x: List[Id]) -> List[Id]:
        return [tab.dot((w, x[j]) for j, w in sorted(r.items()) if x[j] is not None) for r in rows]

    def Ft_times(y: List[Id]) -> List[Id]:      # F_u' * y
        return [tab.dot((F_u[k][i], y[k]) for k in range(nF) if i in F_u[k] and y[k] is not None) for i in range(nU)]

    big = tab.append_unique(it.I_load, (1.7976931348623157e308,))
    WW11: List[Dict[int, int]] = [dict() for _ in range(nU)]
    if smallerNewtonMatrix:                                                               # :242-351
        n = nU + nG
        lamF = [tab.div(lam[k], F[k]) for k in range(nF)]                 # lambda./F
        LPG = [{i: tab.mul([lamF[k], d]) for i, d in F_u[k].items()} for k in range(nF)]   # diag(lambda./F) F_u
        FtLPG: List[Dict[int, list]] = [dict() for _ in range(nU)]
        for k in range(nF):
            for i, di in F_u[k].items():
                for j, lj in LPG[k].items():
                    FtLPG[i].setdefault(j, []).append((di, lj))
        for i in range(nU):
            cols = set(Lf_uu[i]) | set(FtLPG[i]) | {i}
            for j in cols:
                terms = [(1, Lf_uu[i].get(j)), (1, tab.dot(FtLPG[i].get(j, [])))]
                if i == j:
                    terms.append((1, eU))
                WW11[i][j] = tab.add(terms)
        muF = [tab.div(mu, F[k]) for k in range(nF)]
        ffu = [tab.add([(-1, f_u.get(i)), (-1, tab.dot((G_u[k][i], nu[k]) for k in range(nG) if i in G_u[k]))])
               for i in range(nU)]                                        # -f_u - G_u'nu
    else:                                                                                 # :353-466
        n = nU + nG + nF
        for i in range(nU):
            for j, h in Lf_uu[i].items():
                WW11[i][j] = h
            WW11[i][i] = tab.add([(1, Lf_uu[i].get(i)), (1, eU)])
    WW: List[Dict[int, int]] = [dict() for _ in range(n)]
    for i in range(nU):
        WW[i].update(WW11[i])
    for k in range(nG):
        r = nU + k
        for i, d in G_u[k].items():
            WW[r][i] = d
            WW[i][r] = d
        WW[r][r] = tab.neg(eEq)
    if not smallerNewtonMatrix:
        for k in range(nF):
            r = nU + nG + k
            for i, d in F_u[k].items():
                x = tab.neg(d)
                WW[r][i] = x
                WW[i][r] = x
            WW[r][r] = tab.neg(tab.div(F[k], lam[k]))
    fac = ldlmod.ldl(tab, WW, ordering)
    out["factor"] = fac
    if useInertia:                                                                        # :270-274,371-375
        npos = tab.add((1, tab.unary(it.I_heaviside, d)) for d in fac.D)
        nneg = tab.add((1, tab.unary(it.I_heaviside, tab.neg(d))) for d in fac.D)
        code.declare_get("getHessInertia__", [[zero if npos is None else npos], [zero if nneg is None else nneg]])

    negG = [tab.neg(Gk) for Gk in G]
    mehrotra: List[Id] = [None] * nF
    if not skipAffine and nF > 0:
        # ---- affine-scaling direction (:288-318 small, :388-417 large) ----------------------------------
        if smallerNewtonMatrix:
            b_a = ffu + negG
        else:
            b_a = [tab.neg(g) for g in Lf_u] + negG + list(F)
        dx_a = ldlmod.solve(tab, fac, b_a)
        dU_a = dx_a[:nU]
        FudU_a = [tab.dot((d, dU_a[i]) for i, d in sorted(F_u[k].items()) if dU_a[i] is not None) for k in range(nF)]
        if smallerNewtonMatrix:
            dLam_a = [tab.add([(-1, tab.dot((l, dU_a[i]) for i, l in sorted(LPG[k].items()) if dU_a[i] is not None)),
                               (-1, lam[k])]) for k in range(nF)]
        else:
            dLam_a = dx_a[nU + nG:]
        newU_a = [tab.add([(1, u[i]), (1, tab.mul([aP, dU_a[i]]))]) for i in range(nU)]
        newLam_a = [tab.add([(1, lam[k]), (1, tab.mul([aDi, dLam_a[k]]))]) for k in range(nF)]
        newF_a = tab.substitute(F, dict(zip(u, newU_a)))
        mAP = tab.clp(list(zip(F, FudU_a)))
        mAD = tab.clp(list(zip(lam, dLam_a)))
        rho = tab.div(tab.dot(zip(newF_a, newLam_a)), gap)
        code.declare_get("getMaxAlphas_a__", [[big if mAP is None else mAP], [big if mAD is None else mAD]])
        code.declare_get("getMinF_a__", [[tab.minmax(newF_a, True)]])
        code.declare_get("getRho__", [[zero if rho is None else rho]])
        den = F if smallerNewtonMatrix else lam
        mehrotra = [tab.div(tab.mul([FudU_a[k], dLam_a[k]]), den[k]) for k in range(nF)]

    # ---- combined direction ------------------------------------------------------------------------
    if smallerNewtonMatrix:
        FtmuF = Ft_times(muF)
        FtM = Ft_times(mehrotra)
        b = [tab.add([(1, ffu[i]), (1, FtmuF[i]), (-1, FtM[i])]) for i in range(nU)] + negG
    else:
        b = [tab.neg(g) for g in Lf_u] + negG
        b += [tab.add([(1, F[k]), (1, mehrotra[k]), (-1, tab.div(mu, lam[k]))]) for k in range(nF)]
    dx = ldlmod.solve(tab, fac, b)
    out["dx"] = dx
    resid = []
    for i in range(n):
        Wdx = tab.dot((w, dx[j]) for j, w in sorted(WW[i].items()) if dx[j] is not None)
        resid.append(tab.add([(1, Wdx), (-1, b[i])]))
    derr = tab.max_abs(resid)
    code.declare_get("getDirectionError__", [[zero if derr is None else derr]])
    dU, dNu = dx[:nU], dx[nU:nU + nG]
    newU = [tab.add([(1, u[i]), (1, tab.mul([aP, dU[i]]))]) for i in range(nU)]
    newNu = [tab.add([(1, nu[k]), (1, tab.mul([aDe, dNu[k]]))]) for k in range(nG)]
    W11dU = matvec_rows(WW11, dU)
    curvature = tab.dot((dU[i], W11dU[i]) for i in range(nU) if dU[i] is not None and W11dU[i] is not None)
    code.declare_get("getCurvature__", [[zero if curvature is None else curvature]])
    newLam: List[Id] = []
    if nF > 0:
        FudU = [tab.dot((d, dU[i]) for i, d in sorted(F_u[k].items()) if dU[i] is not None) for k in range(nF)]
        if smallerNewtonMatrix:
            dLam = [tab.add([(1, muF[k]), (-1, tab.dot((l, dU[i]) for i, l in sorted(LPG[k].items()) if dU[i] is not None)),
                             (-1, lam[k]), (-1, mehrotra[k])]) for k in range(nF)]
        else:
            dLam = dx[nU + nG:]
        newLam = [tab.add([(1, lam[k]), (1, tab.mul([aDi, dLam[k]]))]) for k in range(nF)]
        mAP = tab.clp(list(zip(F, FudU)))
        mAD = tab.clp(list(zip(lam, dLam)))
        code.declare_get("getMaxAlphas_s__", [[big if mAP is None else mAP], [big if mAD is None else mAD]])
        newF = tab.substitute(F, dict(zip(u, newU)))
        code.declare_get("getMinF_s__", [[tab.minmax(newF, True)]])
    for k, x in enumerate(newLam):
        if x is None:
            raise ValueError("structurally zero dual update")
    code.declare_copy("updatePrimalDual__", list(u) + list(nu) + list(lam), newU + newNu + newLam)
    out.update(dict(F=F, f=f, Lf_u=Lf_u, Lf_uu=Lf_uu, WW=WW, b=b, mu=mu, nU=nU, nG=nG, nF=nF))
    return out
