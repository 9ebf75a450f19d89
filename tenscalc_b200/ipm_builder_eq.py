"""Symbolic construction of everything the equilibrium driver ``ipmPDeq_CSsolver`` gets/sets/copies.

Restates ``lib/ipmPDeqlat_CS.m`` (the builder ``cmex2equilibriumLatentCS`` calls, lib/cmex2equilibriumLatentCS.m:363-388)
for the LARGE Newton matrix (``smallerNewtonMatrix=false``, the generator's default, :563-769): two players

    P1:  min_u  f(u,d,x)  s.t. Fu >= 0, Gu = 0          P2:  min_d  g(u,d,x)  s.t. Fd >= 0, Gd = 0
    latent variables x defined by H(u,d,x) = 0 (as many equations as latent variables)

with Lagrangians Lf = f - P1lambda*Fu + P1nu*Gu + P1xnu*H, Lg = g - P2lambda*Fd + P2nu*Gd + P2xnu*H (:191-208), stacked
z=[u;d;x], F=[Fu;Fd], lambda=[P1lambda;P2lambda], G=[Gu;Gd;H], nu=[P1nu;P1xnu;P2nu;P2xnu] (:108-128) and the
(non-symmetric) Newton system (:568-605)

    WW = [ Lf_uz  Lf_un  Lf_ul ]        b_s = [ -Lf_u ; -Lg_d ; -Lf_x ; -Lg_x ; -G ; -F + mu./lambda ]
         [ Lg_dz  Lg_dn  Lg_dl ]
         [ Lf_xz  Lf_xn  Lf_xl ]        unknowns dx = [ dz ; dnu ; dlambda ]
         [ Lg_xz  Lg_xn  Lg_xl ]
         [ G_z    0      0     ]
         [ F_z    0      diag(F./lambda) ]

factored by the fixed-pattern LU of ``lu.py`` (``useLDL=false``, the default) or, with ``useLDL=true``, by the LDL' of
the normal equations WW'WW (:645-651).  Function names are the reference's ``*__`` names.
"""
from __future__ import annotations

from typing import Dict, List

from . import itable as it
from . import ldl as ldlmod
from . import lu as lumod
from .csparse import CSparse
from .differentiate import Differentiator
from .itable import InstructionTable, Id


def build_ipm_eqlat(code: CSparse, u: List[int], d: List[int], x: List[int], f: Id, g: Id,
                    Fu: List[Id], Fd: List[Id], Gu: List[Id], Gd: List[Id], H: List[Id],
                    P1lam: List[int], P2lam: List[int], P1nu: List[int], P2nu: List[int],
                    P1xnu: List[int], P2xnu: List[int], *, scaleInequalities: bool = True, scaleCost: float = 0.0,
                    addEye2Hessian: bool = False, skipAffine: bool = True, useLDL: bool = False,
                    smallerNewtonMatrix: bool = False, ordering=None) -> Dict[str, object]:
    tab: InstructionTable = code.tab
    nU, nD, nX = len(u), len(d), len(x)
    nZ = nU + nD + nX
    nGu, nGd, nH = len(Gu), len(Gd), len(H)
    nG = nGu + nGd + nH
    nFu, nFd = len(Fu), len(Fd)
    nF = nFu + nFd
    if nH != nX:
        raise ValueError("number of latent variables must match the number of latent constraints "
                         "(cmex2equilibriumLatentCS.m:299-301)")
    if any(e is None for e in Fu + Fd + Gu + Gd + H):
        raise ValueError("constraints that do not involve any variable or parameter")
    one = tab.const(1.0)
    zero = tab.append_unique(it.I_load, (0.0,))
    big = tab.append_unique(it.I_load, (1.7976931348623157e308,))
    out: Dict[str, object] = {}

    # ---- scaling (:74-102) --------------------------------------------------------------------------
    if scaleInequalities and nF > 0:
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
    if scaleCost > 0:
        s4c = tab.new_set("scale4Cost__")
        code.declare_copy("scaleCost__", [s4c], [tab.unary(it.I_abs, tab.div(tab.const(scaleCost), f))])
        f, g = tab.mul([s4c, f]), tab.mul([s4c, g])
        code.declare_get("getScale4Cost__", [[s4c]])
    code.declare_get("getfg__", [[f], [g]])                                                  # :106

    # ---- stacked variables (:108-128) ------------------------------------------------------------------
    F = list(Fu) + list(Fd)
    lam = list(P1lam) + list(P2lam)
    G = list(Gu) + list(Gd) + list(H)
    nu = list(P1nu) + list(P1xnu) + list(P2nu) + list(P2xnu)
    nNu = len(nu)
    z = list(u) + list(d) + list(x)
    gap = tab.dot(zip(lam, F)) if nF else None

    mu = None
    if nF > 0:                                                                               # :140-148
        mu = tab.new_set("mu__")
        code.declare_set("setMu__", [mu])
        code.declare_get("getGapMinFMinLambda__", [[gap], [tab.minmax(F, True)], [tab.minmax(lam, True)]])
    if nG > 0:
        code.declare_get("getNorminf_G__", [[tab.max_abs(G)]])                               # :149-151
    if nu:                                                                                   # :157-175
        code.declare_copy("initDualEq__", nu, [one] * nNu)
    if nF > 0:                                                                               # :176-188
        code.declare_copy("initDualIneq__", lam, [tab.div(mu, e) for e in F])

    # ---- Lagrangians (:191-208) ---------------------------------------------------------------------------
    Lf = tab.add([(1, f)] + ([(-1, tab.dot(zip(P1lam, Fu)))] if nFu else []) + ([(1, tab.dot(zip(P1nu, Gu)))] if nGu else [])
                 + ([(1, tab.dot(zip(P1xnu, H)))] if nH else []))
    Lg = tab.add([(1, g)] + ([(-1, tab.dot(zip(P2lam, Fd)))] if nFd else []) + ([(1, tab.dot(zip(P2nu, Gd)))] if nGd else [])
                 + ([(1, tab.dot(zip(P2xnu, H)))] if nH else []))
    aP = tab.new_set("alphaPrimal__")                                                        # :210-216
    aDe = tab.new_set("alphaDualEq__")
    aDi = tab.new_set("alphaDualIneq__")
    code.declare_set("setAlphaPrimal__", [aP])
    code.declare_set("setAlphaDualEq__", [aDe])
    code.declare_set("setAlphaDualIneq__", [aDi])
    if addEye2Hessian:                                                                       # :220-228 (unused by the large matrix)
        e1 = tab.new_set("addEye2Hessian1__")
        e2 = tab.new_set("addEye2Hessian2__")
        code.declare_set("setAddEye2Hessian1__", [e1])
        code.declare_set("setAddEye2Hessian2__", [e2])

    # ---- derivatives (:231-284): everything with respect to w = [z; nu; lambda], the unknowns' order ---------
    w = z + nu + lam
    pos = {v: k for k, v in enumerate(w)}
    D = Differentiator(tab, w)

    def grad_row(e: Id) -> Dict[int, int]:
        return {} if e is None else {pos[v]: dd for v, dd in D.grad(e).items()}

    def grad_wrt(e: Id, vars_: List[int]) -> List[Id]:
        r = grad_row(e)
        return [r.get(pos[v]) for v in vars_]

    Lf_u, Lg_d = grad_wrt(Lf, u), grad_wrt(Lg, d)
    Lf_x, Lg_x = grad_wrt(Lf, x), grad_wrt(Lg, x)
    norms = [tab.max_abs(Lf_u)] + ([tab.max_abs(Lf_x)] if nX else []) + [tab.max_abs(Lg_d)] + ([tab.max_abs(Lg_x)] if nX else [])
    ng = tab.add([(1, e) for e in norms if e is not None])
    code.declare_get("getNorminf_Grad__", [[zero if ng is None else ng]])                    # :238,247

    residuals = Lf_u + Lg_d + Lf_x + Lg_x
    if any(r is None for r in residuals):
        raise ValueError("a stationarity condition is structurally zero: some optimisation variable does not enter its "
                         "player's Lagrangian")
    neg = tab.neg
    F_z = [{c: v for c, v in grad_row(e).items() if c < nZ} for e in F]
    if smallerNewtonMatrix:
        # ---- small matrix (:314-446): unknowns [z; nu]; dlambda = -lambda - LFF dz + mu./F is eliminated, ------
        # LFF = diag(lambda./F) F_z (:304).  Row of a stationarity condition r: [r_z - r_lambda LFF, r_nu]; then [G_z, 0]
        n = nZ + nNu
        lamF = [tab.div(lam[k], F[k]) for k in range(nF)]
        LFF = [{c: tab.mul([lamF[k], v]) for c, v in F_z[k].items()} for k in range(nF)]
        r_l: List[Dict[int, int]] = []              # d residual / d lambda (Lf_ul, Lg_dl, Lf_xl, Lg_xl)
        WW = []
        for i, r in enumerate(residuals):
            g_ = grad_row(r)
            rl = {c - nZ - nNu: v for c, v in g_.items() if c >= nZ + nNu}
            r_l.append(rl)
            row: Dict[int, int] = {c: v for c, v in g_.items() if nZ <= c < nZ + nNu}
            cols = sorted({c for c in g_ if c < nZ} | {c for k in rl for c in LFF[k]})
            for c in cols:
                corr = tab.dot((rl[k], LFF[k][c]) for k in sorted(rl) if c in LFF[k])
                e = tab.add(([(1, g_[c])] if c in g_ else []) + ([(-1, corr)] if corr is not None else []))
                if e is not None:
                    row[c] = e
            if nX == 0 and addEye2Hessian and i < nZ:                                        # :362-365 (no latent variables)
                row[i] = tab.add([(1, row[i]), (1, e1)]) if i in row else e1
            WW.append(row)
        for j, e in enumerate(G):
            row = {c: v for c, v in grad_row(e).items() if c < nZ}
            if nX == 0 and addEye2Hessian:
                row[nZ + j] = neg(e2)
            WW.append(row)
        if len(WW) != n:
            raise AssertionError("Newton matrix is not square")
        # right-hand sides (:338-361,383-403): the gradients of f + nu'G WITHOUT the inequality terms, which enter
        # through the eliminated dlambda
        ff = tab.add([(1, f)] + ([(1, tab.dot(zip(P1nu, Gu)))] if nGu else []) + ([(1, tab.dot(zip(P1xnu, H)))] if nX else []))
        gg = tab.add([(1, g)] + ([(1, tab.dot(zip(P2nu, Gd)))] if nGd else []) + ([(1, tab.dot(zip(P2xnu, H)))] if nX else []))
        grads = grad_wrt(ff, u) + grad_wrt(gg, d) + grad_wrt(ff, x) + grad_wrt(gg, x)
        b_a = [neg(e) for e in grads] + [neg(e) for e in G]
    else:
        # ---- large matrix (:563-600): rows in the reference's order; every row is the gradient of one residual with
        # respect to w, except that the last block is diag(F./lambda) instead of the derivative with respect to lambda
        n = nZ + nNu + nF
        WW = [grad_row(r) for r in residuals]
        for e in G:
            WW.append({k: v for k, v in grad_row(e).items() if k < nZ})                      # [G_z, 0, 0]
        for k, e in enumerate(F):
            row = {c: v for c, v in grad_row(e).items() if c < nZ}                           # [F_z, 0, diag(F./lambda)]
            row[nZ + nNu + k] = tab.div(e, lam[k])
            WW.append(row)
        if len(WW) != n:
            raise AssertionError("Newton matrix is not square")
        b_a = [neg(r) for r in residuals] + [neg(e) for e in G] + [neg(e) for e in F]

    # ---- factorisation (:645-654) ---------------------------------------------------------------------------
    def symmetrise(M: List[Dict[int, int]], rhs: List[Id]):
        cols: List[Dict[int, int]] = [dict() for _ in range(n)]           # cols[j] = {row: instr}
        for r, row in enumerate(M):
            for c, v in row.items():
                cols[c][r] = v
        MtM: List[Dict[int, int]] = [dict() for _ in range(n)]
        for i in range(n):
            for j in range(i + 1):
                common = sorted(set(cols[i]) & set(cols[j]))
                if common:
                    v = tab.dot((cols[i][k], cols[j][k]) for k in common)
                    if v is not None:
                        MtM[i][j] = v
                        MtM[j][i] = v
        Mtb = [tab.dot((cols[i][k], rhs[k]) for k in sorted(cols[i]) if rhs[k] is not None) for i in range(n)]
        return MtM, Mtb

    WWs = WW
    if useLDL:
        WWs, _ = symmetrise(WW, b_a)
        fac = ldlmod.ldl(tab, WWs, ordering)
        fsolve = lambda rhs: ldlmod.solve(tab, fac, symmetrise(WW, rhs)[1])      # noqa: E731
    else:
        fac = lumod.lu(tab, WW, ordering)
        fsolve = lambda rhs: lumod.solve(tab, fac, rhs)                          # noqa: E731
    out["factor"] = fac

    def times_row(row: Dict[int, int], v: List[Id]) -> Id:
        return tab.dot((m, v[c]) for c, m in sorted(row.items()) if v[c] is not None)

    def stepped(vars_: List[int], dv: List[Id], alpha: int) -> List[Id]:
        return [tab.add([(1, a), (1, tab.mul([alpha, b]))]) for a, b in zip(vars_, dv)]

    def lam_step(dZ_: List[Id], with_mu: bool, meh: List[Id]) -> List[Id]:
        """small matrix: dlambda = -lambda - LFF dz [+ mu./F] [- Mehrotra] (:466,516-524)"""
        outl = []
        for k in range(nF):
            t = times_row(LFF[k], dZ_)
            outl.append(tab.add([(-1, lam[k])] + ([(-1, t)] if t is not None else [])
                                + ([(1, tab.div(mu, F[k]))] if with_mu else []) + ([(-1, meh[k])] if meh[k] is not None else [])))
        return outl

    def minus_rl(vec: List[Id], sign: int) -> List[Id]:
        """sign * [Lf_ul; Lg_dl; Lf_xl; Lg_xl] vec, zero for the G rows (:358-361,484-494)"""
        rows = [tab.dot((r_l[i][k], vec[k]) for k in sorted(r_l[i]) if vec[k] is not None) for i in range(len(residuals))]
        return [None if e is None else (e if sign > 0 else neg(e)) for e in rows] + [None] * nG

    def plus(a: List[Id], b: List[Id]) -> List[Id]:
        return [tab.add(([(1, x)] if x is not None else []) + ([(1, y)] if y is not None else [])) for x, y in zip(a, b)]

    mehrotra: List[Id] = [None] * nF
    if not skipAffine and nF > 0:                                                            # :656-713 / :433-497
        dx_a = fsolve(b_a)
        dZ_a = dx_a[:nZ]
        dLam_a = lam_step(dZ_a, False, mehrotra) if smallerNewtonMatrix else dx_a[nZ + nNu:]
        newZ_a = stepped(z, dZ_a, aP)
        newLam_a = stepped(lam, dLam_a, aDi)
        newF_a = tab.substitute(F, dict(zip(z, newZ_a)))
        FzdZ_a = [times_row(F_z[k], dZ_a) for k in range(nF)]
        mAP = tab.clp(list(zip(F, FzdZ_a)))
        mAD = tab.clp(list(zip(lam, dLam_a)))
        # Mehrotra correction: (F_z dz).*dlambda./lambda in the large system (:697), ./F in the small one (:481)
        mehrotra = [tab.div(tab.mul([FzdZ_a[k], dLam_a[k]]), F[k] if smallerNewtonMatrix else lam[k]) for k in range(nF)]
        rho = tab.div(tab.dot(zip(newF_a, newLam_a)), gap)
        code.declare_get("getMaxAlphas_a__", [[big if mAP is None else mAP], [big if mAD is None else mAD]])
        code.declare_get("getMinF_a__", [[tab.minmax(newF_a, True)]])
        code.declare_get("getRho__", [[zero if rho is None else rho]])

    # ---- combined direction (:716-768 / :500-560) ---------------------------------------------------------------
    if smallerNewtonMatrix:
        b_s = list(b_a)
        if nF > 0:
            b_s = plus(b_s, minus_rl([tab.div(mu, F[k]) for k in range(nF)], -1))
            if any(m is not None for m in mehrotra):
                b_s = plus(b_s, minus_rl(mehrotra, +1))
    else:
        b_s = [neg(r) for r in residuals] + [neg(e) for e in G]
        b_s += [tab.add([(-1, F[k]), (-1, mehrotra[k]), (1, tab.div(mu, lam[k]))]) for k in range(nF)]
    dx = fsolve(b_s)
    out["dx"] = dx
    if useLDL:                                   # WW and b_s were replaced by the symmetrised ones (:647-650)
        Ws, bs = symmetrise(WW, b_s)
    else:
        Ws, bs = WW, b_s
    resid = [tab.add([(1, times_row(Ws[i], dx))] + ([(-1, bs[i])] if bs[i] is not None else [])) for i in range(n)]
    derr = tab.max_abs(resid)
    code.declare_get("getDirectionError__", [[zero if derr is None else derr]])              # :726
    dZ, dNu = dx[:nZ], dx[nZ:nZ + nNu]
    dLam = (lam_step(dZ, True, mehrotra) if nF else []) if smallerNewtonMatrix else dx[nZ + nNu:]
    newZ = stepped(z, dZ, aP)
    newNu = stepped(nu, dNu, aDe)
    newLam: List[Id] = []
    if nF > 0:
        FzdZ = [times_row(F_z[k], dZ) for k in range(nF)]
        mAP = tab.clp(list(zip(F, FzdZ)))
        mAD = tab.clp(list(zip(lam, dLam)))
        newLam = stepped(lam, dLam, aDi)
        newF = tab.substitute(F, dict(zip(z, newZ)))
        code.declare_get("getMaxAlphas_s__", [[big if mAP is None else mAP], [big if mAD is None else mAD]])
        code.declare_get("getMinF_s__", [[tab.minmax(newF, True)]])
    for e in newZ + newNu + newLam:
        if e is None:
            raise ValueError("structurally zero update")
    # the reference lists the destinations as u,d,x,P1nu,P2nu,P1xnu,P2xnu,P1lambda,P2lambda (:800-829); a copy is a set of
    # independent scalar moves, so the order [z; nu; lambda] used here is equivalent
    code.declare_copy("updatePrimalDual__", z + nu + lam, newZ + newNu + newLam)
    out.update(dict(F=F, f=f, g=g, WW=WW, b=b_s, mu=mu, nZ=nZ, nU=nU, nD=nD, nX=nX, nG=nG, nF=nF, nNu=nNu))
    return out
