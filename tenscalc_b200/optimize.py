"""``cmex2optimizeCS`` -- generate a batched interior-point solver for

        minimise  objective(u)   s.t.  G(u)=0,  F(u)>=0

Mirrors the name/value surface of the reference generator
(``lib/cmex2optimizeCS.m:1-52``; defaults ``lib/private/parameters4all.m`` and
``parameters4optimize.m``) and its sequence of declarations (:156-355): sets for
parameters and optimisation variables, ``parseConstraints`` (private/parseConstraints.m:72-132),
primal/dual packing (private/packVariables.m:38-70), ``initPrimal__``, the
``ipmPD_CStimesLambda`` builder, the ``#define`` table handed to the driver (:303-331)
and ``getOutputs``.

The returned :class:`GeneratedSolver` plays the role of the generated class: calling it
creates an object with ``setP_<name>/setV_<name>/solve/getOutputs`` working on a whole
batch of instances on one GPU (``runtime.BatchedSolver``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import tensor as T
from .csparse import CSparse, CompiledProblem
from .ipm_builder import build_ipm_pd, build_ipm_timeslambda
from .tensor import Constraint, Tcalculus


@dataclass
class GeneratedSolver:
    classname: str
    problem: CompiledProblem
    parameters: List[Tuple[str, Tuple[int, ...]]]
    variables: List[Tuple[str, Tuple[int, ...]]]
    outputs: List[Tuple[str, Tuple[int, ...]]]
    output_is_struct: bool
    defines: Dict[str, object]
    info: Dict[str, object] = field(default_factory=dict)

    def __call__(self, batch: int = 1, device: Optional[int] = None, backend: str = "compiled", **module_opts):
        from .runtime import BatchedSolver
        return BatchedSolver(self, batch, device, backend, **module_opts)


def solver_defines(nU: int, nG: int, nF: int, *, gradTolerance: float = 1e-4, equalTolerance: float = 1e-4,
                   desiredDualityGap: float = 1e-5, alphaMin: float = 1e-7, alphaMax: float = 1.0,
                   coupledAlphas: bool = True, muFactorAggressive: float = .2, muFactorConservative: float = .95,
                   skipAffine: bool = True, delta: int = 3, scaleInequalities: bool = True, scaleCost: float = 0.0,
                   scaleEqualities: bool = False, addEye2Hessian: bool = True, adjustAddEye2Hessian: bool = True,
                   useInertia: bool = False, addEye2HessianUtolerance: float = 1e-6, solverVerboseLevel: int = 1,
                   smallerNewtonMatrix: bool = False, allowSave: bool = False, **ignored) -> Dict[str, object]:
    """the #define table cmex2optimizeCS hands to the C driver (lib/cmex2optimizeCS.m:303-331; defaults
    lib/private/parameters4all.m, parameters4optimize.m)"""
    return dict(
        nU=int(nU), nG=int(nG), nF=int(nF),
        gradTolerance=float(gradTolerance), addEye2HessianUtolerance=float(addEye2HessianUtolerance),
        equalTolerance=float(equalTolerance), desiredDualityGap=float(desiredDualityGap),
        scaleCost=int(scaleCost != 0), scaleInequalities=int(bool(scaleInequalities)),
        scaleEqualities=int(bool(scaleEqualities)),
        alphaMin=float(alphaMin), alphaMax=float(alphaMax), coupledAlphas=int(bool(coupledAlphas)),
        setAddEye2Hessian=int(bool(addEye2Hessian)), adjustAddEye2Hessian=int(bool(adjustAddEye2Hessian)),
        useInertia=int(bool(useInertia)), muFactorAggressive=float(muFactorAggressive),
        muFactorConservative=float(muFactorConservative), delta=int(delta), skipAffine=int(bool(skipAffine)),
        smallerNewtonMatrix=int(bool(smallerNewtonMatrix)), useLDL=1, useUmfpack=0, allowSave=int(bool(allowSave)),
        debugConvergence=0, profiling=0, verboseLevel=int(min(solverVerboseLevel, 1)))


def _name_of(v: Tcalculus) -> str:
    if not v.is_variable():
        raise ValueError("parameters and optimizationVariables must be Tvariables")
    return v.name


def cmex2optimizeCS(*, classname: str, objective: Tcalculus,
                    optimizationVariables: Sequence[Tcalculus],
                    constraints: Sequence[Constraint] = (),
                    outputExpressions: Union[Sequence[Tcalculus], Dict[str, Tcalculus]] = (),
                    parameters: Sequence[Tcalculus] = (),
                    gradTolerance: float = 1e-4, equalTolerance: float = 1e-4,
                    desiredDualityGap: float = 1e-5, alphaMin: float = 1e-7, alphaMax: float = 1.0,
                    coupledAlphas: bool = True, muFactorAggressive: float = .2,
                    muFactorConservative: float = .95, skipAffine: bool = True, delta: int = 3,
                    scaleInequalities: bool = True, scaleCost: float = 0.0,
                    scaleEqualities: bool = False, addEye2Hessian: bool = True,
                    adjustAddEye2Hessian: bool = True, useInertia: bool = False,
                    addEye2HessianUtolerance: float = 1e-6, useLDL: bool = True,
                    smallerNewtonMatrix: bool = False, useUmfpack: bool = False,
                    sensitivityVariables: Sequence[Tcalculus] = (),
                    solverVerboseLevel: int = 1, ordering=None, algorithm: str = "ipmPD_CStimesLambda",
                    allowSave: bool = False, **ignored) -> GeneratedSolver:
    # the reference's C generator hard-wires algorithm 3 and rejects the options below
    # (cmex2optimizeCS.m:83-93, ipmPD_CStimesLambda.m:31-48)
    # ``algorithm`` selects the symbolic builder: "ipmPD_CStimesLambda" is what the shipped generator always
    # uses (its `switch 3`, cmex2optimizeCS.m:83-93, which also forces skipAffine=true); "ipmPD_CS" is the
    # builder of the same file's cases 1/2 (ipmPD_CS.m) with the affine-scaling (Mehrotra) step and the
    # reduced Newton matrix that ipmPD_CSsolver.c supports (:576-665)
    if algorithm not in ("ipmPD_CStimesLambda", "ipmPD_CS"):
        raise ValueError("algorithm must be 'ipmPD_CStimesLambda' or 'ipmPD_CS'")
    if sensitivityVariables:
        raise NotImplementedError("sensitivityVariables: not implemented by ipmPD_CStimesLambda")
    if not useLDL or useUmfpack:
        raise NotImplementedError("useLDL=false/useUmfpack=true: not implemented by ipmPD_CStimesLambda")
    if algorithm == "ipmPD_CStimesLambda":
        if smallerNewtonMatrix:
            raise NotImplementedError("smallerNewtonMatrix=true: not implemented by ipmPD_CStimesLambda")
        skipAffine = True

    tab = T.table()
    code = CSparse(tab)
    if objective.numel() != 1:
        raise ValueError("objective must be a scalar")

    # ---- sets for parameters and optimisation variables (:156-195) -----------------------
    par_info = []
    for p in parameters:
        nm = _name_of(p)
        code.declare_set(f"setP_{nm}", p.flat(), p.shape, kind="parameter")
        par_info.append((nm, p.shape))
    # The symbolic variables themselves become the packed primal vector u (the reference
    # substitutes them by slices of a new variable `x_`, packVariables.m:38-70); setV_*
    # writes separate storage that initPrimal__ copies into u (:216-218).
    u: List[int] = []
    u0: List[int] = []
    var_info = []
    for v in optimizationVariables:
        nm = _name_of(v)
        ids = v.flat()
        store = [tab.new_set(f"{nm}_init[{k}]") for k in range(len(ids))]
        code.declare_set(f"setV_{nm}", store, v.shape, kind="variable")
        u += ids
        u0 += store
        var_info.append((nm, v.shape))
    if len(set(u)) != len(u):
        raise ValueError("optimizationVariables must be distinct")
    code.declare_copy("initPrimal__", u, u0)

    # ---- constraints -> G (==0), F (>=0) and dual variables (parseConstraints.m:72-132) -----
    G: List[Optional[int]] = []
    F: List[Optional[int]] = []
    for c in constraints:
        if not isinstance(c, Constraint):
            raise ValueError("constraint is not a Tcalculus comparison")
        (G if c.kind == "iszero" else F).extend(c.expr.flat())
    nu = [tab.new_set(f"nu__[{k}]") for k in range(len(G))]
    lam = [tab.new_set(f"lambda__[{k}]") for k in range(len(F))]

    if algorithm == "ipmPD_CS":
        info = build_ipm_pd(code, u, objective.e[()], F, G, lam, nu,
                            scaleInequalities=scaleInequalities, scaleCost=scaleCost,
                            addEye2Hessian=addEye2Hessian, useInertia=useInertia, skipAffine=skipAffine,
                            smallerNewtonMatrix=smallerNewtonMatrix, ordering=ordering)
    else:
        info = build_ipm_timeslambda(code, u, objective.e[()], F, G, lam, nu,
                                     scaleInequalities=scaleInequalities, scaleCost=scaleCost,
                                     addEye2Hessian=addEye2Hessian, useInertia=useInertia,
                                     ordering=ordering)

    # ---- outputs (:337-355) --------------------------------------------------------------------
    is_struct = isinstance(outputExpressions, dict)
    items = list(outputExpressions.items()) if is_struct else \
        [(f"y{k + 1}", e) for k, e in enumerate(outputExpressions)]
    out_info = []
    srcs, shapes, names = [], [], []
    for nm, e in items:
        e = Tcalculus._lift(e)
        srcs.append(e.flat())
        shapes.append(e.shape)
        names.append(nm)
        out_info.append((nm, e.shape))
    # ---- allowSave (cmex2optimizeCS.m:267,327; ipmPD_CStimesLambda.m:288-290 declareSave(code,WW,'saveWW__',...)) ----
    # The reference's saveWW__ writes the values of WW's structural non-zeros to `<class>_WW.values` (their
    # subscripts to `<class>_WW.subscripts` at generation time).  A batch has no file per instance: saveWW__ is a
    # copy of those values into save variables that getOutputs returns as one more, last, output `WW_values__`
    # (subscripts: info["WW_subscripts"], column-major order like the reference's tensors).
    ww_subs = None
    if allowSave:
        WW = info["WW"]
        entries = sorted((c, r, x) for r, row in enumerate(WW) for c, x in row.items() if x is not None)
        ww_subs = np.array([(r, c) for c, r, _ in entries], dtype=np.int64).reshape(-1, 2)
        store = [tab.new_set(f"WW_saved__[{k}]") for k in range(len(entries))]
        code.declare_copy("saveWW__", store, [x for _, _, x in entries])
        srcs.append(store)
        shapes.append((len(store),))
        names.append("WW_values__")
        out_info.append(("WW_values__", (len(store),)))
    code.declare_get("getOutputs", srcs, shapes, names)

    defines = solver_defines(len(u), len(G), len(F), gradTolerance=gradTolerance, equalTolerance=equalTolerance,
                             desiredDualityGap=desiredDualityGap, alphaMin=alphaMin, alphaMax=alphaMax,
                             coupledAlphas=coupledAlphas, muFactorAggressive=muFactorAggressive,
                             muFactorConservative=muFactorConservative, skipAffine=skipAffine, delta=delta,
                             scaleInequalities=scaleInequalities, scaleCost=scaleCost, scaleEqualities=scaleEqualities,
                             addEye2Hessian=addEye2Hessian, adjustAddEye2Hessian=adjustAddEye2Hessian,
                             useInertia=useInertia, addEye2HessianUtolerance=addEye2HessianUtolerance,
                             solverVerboseLevel=solverVerboseLevel, smallerNewtonMatrix=smallerNewtonMatrix,
                             allowSave=allowSave)
    problem = code.compile(defines)
    fac = info["factor"]
    from .ordering import DEFAULT as default_ordering
    from .ldl import dense_trailing_block as ldl_dense_block
    gen_info = dict(ordering=(ordering if isinstance(ordering, str) else "custom") if ordering is not None
                    else default_ordering,
                    nnzL=fac.nnz(), kkt=fac.n, n_instructions=sum(len(g) for g in problem.groups),
                    n_groups=len(problem.groups), n_slots=problem.n_slots,
                    dense_block=ldl_dense_block(fac))
    if ww_subs is not None:
        gen_info["WW_subscripts"] = ww_subs
    return GeneratedSolver(classname, problem, par_info, var_info, out_info, is_struct, defines, gen_info)


# ------------------------------------------------------------------------------------------------
# Nash equilibria with latent variables: cmex2equilibriumLatentCS -> ipmPDeq_CSsolver
# ------------------------------------------------------------------------------------------------
def equilibrium_defines(nU: int, nD: int, nX: int, nG: int, nF: int, nNu: int, *, gradTolerance: float = 1e-4,
                        equalTolerance: float = 1e-4, desiredDualityGap: float = 1e-5, alphaMin: float = 1e-7,
                        alphaMax: float = 1.0, coupledAlphas: bool = True, muFactorAggressive: float = .2,
                        muFactorConservative: float = .95, skipAffine: bool = True, delta: int = 3,
                        scaleInequalities: bool = True, scaleCost: float = 0.0, scaleEqualities: bool = False,
                        addEye2Hessian: bool = False, adjustAddEye2Hessian: bool = False,
                        addEye2Hessian1tolerance: float = 1e-6, useLDL: bool = False, smallerNewtonMatrix: bool = False,
                        solverVerboseLevel: int = 1, **ignored) -> Dict[str, object]:
    """the #define table cmex2equilibriumLatentCS hands to ipmPDeq_CSsolver.c (lib/cmex2equilibriumLatentCS.m:417-448;
    defaults lib/private/parameters4equilibrium.m, parameters4all.m).  ``driver`` (not a reference define) tells the
    device image which solver loop to run; nU/nG/nF/addEye2HessianUtolerance are the names the device driver reads."""
    return dict(
        nZ=int(nU + nD + nX), nU=int(nU), nD=int(nD), nX=int(nX), nG=int(nG), nF=int(nF), nNu=int(nNu),
        gradTolerance=float(gradTolerance), addEye2Hessian1tolerance=float(addEye2Hessian1tolerance),
        addEye2HessianUtolerance=float(addEye2Hessian1tolerance),
        equalTolerance=float(equalTolerance), desiredDualityGap=float(desiredDualityGap),
        scaleCost=int(scaleCost != 0), scaleInequalities=int(bool(scaleInequalities)),
        scaleEqualities=int(bool(scaleEqualities)),
        alphaMin=float(alphaMin), alphaMax=float(alphaMax), coupledAlphas=int(bool(coupledAlphas)),
        setAddEye2Hessian=int(bool(addEye2Hessian)), adjustAddEye2Hessian=int(bool(adjustAddEye2Hessian)),
        useInertia=0, muFactorAggressive=float(muFactorAggressive),
        muFactorConservative=float(muFactorConservative), delta=int(delta), skipAffine=int(bool(skipAffine)),
        smallerNewtonMatrix=int(bool(smallerNewtonMatrix)), useLDL=int(bool(useLDL)), useUmfpack=0, allowSave=0,
        debugConvergence=0, profiling=0, verboseLevel=int(min(solverVerboseLevel, 1)), driver="ipmPDeq_CSsolver")


def cmex2equilibriumLatentCS(*, classname: str, P1objective: Tcalculus, P2objective: Tcalculus,
                             P1optimizationVariables: Sequence[Tcalculus], P2optimizationVariables: Sequence[Tcalculus],
                             latentVariables: Sequence[Tcalculus] = (), P1constraints: Sequence[Constraint] = (),
                             P2constraints: Sequence[Constraint] = (), latentConstraints: Sequence[Constraint] = (),
                             outputExpressions: Union[Sequence[Tcalculus], Dict[str, Tcalculus]] = (),
                             parameters: Sequence[Tcalculus] = (), ordering=None, **options) -> GeneratedSolver:
    """``cmex2equilibriumLatentCS`` (lib/cmex2equilibriumLatentCS.m:1-596): generates the solver of the two-player
    Nash equilibrium with latent variables described in ``ipm_builder_eq``; the generated class has the usual
    ``setP_*/setV_*/solve/getOutputs`` surface and its ``solve`` runs ``ipmPDeq_CSsolver`` (lib/ipmPDeq_CSsolver.c).
    Options as in the reference (``addEye2Hessian``, ``adjustAddEye2Hessian`` and ``useLDL`` default to false)."""
    from .ipm_builder_eq import build_ipm_eqlat
    if options.get("adjustAddEye2Hessian") and not options.get("addEye2Hessian"):
        raise ValueError("adjustAddEye2Hessian needs addEye2Hessian")
    if options.get("useUmfpack"):
        raise NotImplementedError("useUmfpack=true (atomic factorisation) is not built")
    tab = T.table()
    code = CSparse(tab)
    if P1objective.numel() != 1 or P2objective.numel() != 1:
        raise ValueError("objectives must be scalars")
    par_info = []
    for p in parameters:
        nm = _name_of(p)
        code.declare_set(f"setP_{nm}", p.flat(), p.shape, kind="parameter")
        par_info.append((nm, p.shape))
    var_info = []

    def pack(vars_: Sequence[Tcalculus]):
        ids, store = [], []
        for v in vars_:
            nm = _name_of(v)
            e = v.flat()
            st = [tab.new_set(f"{nm}_init[{k}]") for k in range(len(e))]
            code.declare_set(f"setV_{nm}", st, v.shape, kind="variable")
            ids += e
            store += st
            var_info.append((nm, v.shape))
        return ids, store

    u, u0 = pack(P1optimizationVariables)                     # :230-246 / packVariables.m
    d, d0 = pack(P2optimizationVariables)
    x, x0 = pack(latentVariables)
    if len(set(u + d + x)) != len(u + d + x):
        raise ValueError("optimisation and latent variables must be distinct")
    code.declare_copy("initPrimal__", u + d + x, u0 + d0 + x0)   # :303

    def parse(cons, what):
        Gs: List[Optional[int]] = []
        Fs: List[Optional[int]] = []
        for c in cons:
            if not isinstance(c, Constraint):
                raise ValueError(f"{what} constraint is not a Tcalculus comparison")
            (Gs if c.kind == "iszero" else Fs).extend(c.expr.flat())
        return Gs, Fs

    Gu, Fu = parse(P1constraints, "P1")                       # :252-262
    Gd, Fd = parse(P2constraints, "P2")
    H, Herr = parse(latentConstraints, "latent")
    if Herr:
        raise ValueError("latent constraints cannot be inequalities (cmex2equilibriumLatentCS.m:273-276)")
    mk = lambda nm, n: [tab.new_set(f"{nm}[{k}]") for k in range(n)]     # noqa: E731
    P1nu, P1lam = mk("P1nu_", len(Gu)), mk("P1lambda_", len(Fu))
    P2nu, P2lam = mk("P2nu_", len(Gd)), mk("P2lambda_", len(Fd))
    P1xnu, P2xnu = mk("P1xnu_", len(H)), mk("P2xnu_", len(H))
    keys = ("scaleInequalities", "scaleCost", "addEye2Hessian", "skipAffine", "useLDL", "smallerNewtonMatrix")
    info = build_ipm_eqlat(code, u, d, x, P1objective.e[()], P2objective.e[()], Fu, Fd, Gu, Gd, H, P1lam, P2lam, P1nu,
                           P2nu, P1xnu, P2xnu, ordering=ordering, **{k: options[k] for k in keys if k in options})
    is_struct = isinstance(outputExpressions, dict)
    items = list(outputExpressions.items()) if is_struct else [(f"y{k + 1}", e) for k, e in enumerate(outputExpressions)]
    out_info, srcs, shapes, names = [], [], [], []
    for nm, e in items:
        e = Tcalculus._lift(e)
        srcs.append(e.flat())
        shapes.append(e.shape)
        names.append(nm)
        out_info.append((nm, e.shape))
    code.declare_get("getOutputs", srcs, shapes, names)
    defines = equilibrium_defines(len(u), len(d), len(x), info["nG"], info["nF"], info["nNu"], **options)
    problem = code.compile(defines)
    fac = info["factor"]
    gen_info = dict(ordering="lu threshold pivoting" if not options.get("useLDL") else "symamd (normal equations)",
                    nnzL=fac.nnz(), kkt=fac.n, n_instructions=sum(len(g) for g in problem.groups),
                    n_groups=len(problem.groups), n_slots=problem.n_slots, driver="ipmPDeq_CSsolver")
    return GeneratedSolver(classname, problem, par_info, var_info, out_info, is_struct, defines, gen_info)


# ------------------------------------------------------------------------------------------------
# min-max problems: cmex2minmaxCS / class2minmaxCS -> ipmPDminmax_CSsolver
# ------------------------------------------------------------------------------------------------
def minmax_defines(nU: int, nD: int, nG: int, nF: int, nGd: int, nFd: int, *, gradTolerance: float = 1e-4,
                   equalTolerance: float = 1e-4, desiredDualityGap: float = 1e-5, alphaMin: float = 1e-7,
                   alphaMax: float = 1.0, coupledAlphas: bool = True, muFactorAggressive: float = .2,
                   muFactorConservative: float = .95, scaleInequalities: bool = True, scaleCost: float = 0.0,
                   addEye2Hessian: bool = True, adjustAddEye2Hessian: bool = False,
                   addEye2HessianUtolerance: float = 1e-6, addEye2HessianDtolerance: float = 1e-6,
                   solverVerboseLevel: int = 1, **ignored) -> Dict[str, object]:
    """the properties the generated minmax class hands to its solver (lib/class2minmaxCS.m; defaults
    lib/private/parameters4minmax.m, parameters4all.m)"""
    return dict(
        nZ=int(nU + nD), nU=int(nU), nD=int(nD), nG=int(nG), nF=int(nF),
        mpUdesired=int(nU + nGd + nFd), mnDdesired=int(nD),       # ipmPDminmax_CSsolver.m:68-69
        gradTolerance=float(gradTolerance), addEye2HessianUtolerance=float(addEye2HessianUtolerance),
        addEye2HessianDtolerance=float(addEye2HessianDtolerance), equalTolerance=float(equalTolerance),
        desiredDualityGap=float(desiredDualityGap), scaleCost=int(scaleCost != 0),
        scaleInequalities=int(bool(scaleInequalities)), scaleEqualities=0, alphaMin=float(alphaMin),
        alphaMax=float(alphaMax), coupledAlphas=int(bool(coupledAlphas)), setAddEye2Hessian=int(bool(addEye2Hessian)),
        adjustAddEye2Hessian=int(bool(adjustAddEye2Hessian)), useInertia=1,
        muFactorAggressive=float(muFactorAggressive), muFactorConservative=float(muFactorConservative), delta=3,
        skipAffine=1, smallerNewtonMatrix=0, useLDL=1, useUmfpack=0, allowSave=0, debugConvergence=0, profiling=0,
        verboseLevel=int(min(solverVerboseLevel, 1)), driver="ipmPDminmax_CSsolver")


def cmex2minmaxCS(*, classname: str, objective: Tcalculus, minOptimizationVariables: Sequence[Tcalculus],
                  maxOptimizationVariables: Sequence[Tcalculus], minConstraints: Sequence[Constraint] = (),
                  maxConstraints: Sequence[Constraint] = (),
                  outputExpressions: Union[Sequence[Tcalculus], Dict[str, Tcalculus]] = (),
                  parameters: Sequence[Tcalculus] = (), ordering=None, **options) -> GeneratedSolver:
    """``cmex2minmaxCS`` / ``class2minmaxCS`` (lib/cmex2minmaxCS.m, lib/class2minmaxCS.m): min over the first set of
    variables of the max over the second; the generated class's ``solve`` runs ``ipmPDminmax_CSsolver``
    (lib/ipmPDminmax_CSsolver.m -- the reference's C file of that name is stale, SURVEY.md Note A, so the MATLAB
    solver is the specification and parity with it is UNPINNED: no oracle of the reference's own can be run here).
    ``solve`` takes three regularisers ``addEye2Hessian=[U, D, Eq]``."""
    from .ipm_builder_minmax import build_ipm_minmax
    tab = T.table()
    code = CSparse(tab)
    if objective.numel() != 1:
        raise ValueError("objective must be a scalar")
    par_info = []
    for p in parameters:
        nm = _name_of(p)
        code.declare_set(f"setP_{nm}", p.flat(), p.shape, kind="parameter")
        par_info.append((nm, p.shape))
    var_info = []

    def pack(vars_):
        ids, store = [], []
        for v in vars_:
            nm = _name_of(v)
            e = v.flat()
            st = [tab.new_set(f"{nm}_init[{k}]") for k in range(len(e))]
            code.declare_set(f"setV_{nm}", st, v.shape, kind="variable")
            ids += e
            store += st
            var_info.append((nm, v.shape))
        return ids, store

    u, u0 = pack(minOptimizationVariables)
    d, d0 = pack(maxOptimizationVariables)
    if len(set(u + d)) != len(u + d):
        raise ValueError("optimisation variables must be distinct")
    code.declare_copy("initPrimal__", u + d, u0 + d0)

    def parse(cons):
        Gs: List[Optional[int]] = []
        Fs: List[Optional[int]] = []
        for c in cons:
            if not isinstance(c, Constraint):
                raise ValueError("constraint is not a Tcalculus comparison")
            (Gs if c.kind == "iszero" else Fs).extend(c.expr.flat())
        return Gs, Fs

    Gu, Fu = parse(minConstraints)
    Gd, Fd = parse(maxConstraints)
    mk = lambda nm, n: [tab.new_set(f"{nm}[{k}]") for k in range(n)]     # noqa: E731
    nuU, lamU, nuD, lamD = mk("nuU_", len(Gu)), mk("lambdaU_", len(Fu)), mk("nuD_", len(Gd)), mk("lambdaD_", len(Fd))
    keys = ("scaleInequalities", "scaleCost", "addEye2Hessian")
    info = build_ipm_minmax(code, u, d, objective.e[()], Fu, Fd, Gu, Gd, lamU, lamD, nuU, nuD, ordering=ordering,
                            **{k: options[k] for k in keys if k in options})
    is_struct = isinstance(outputExpressions, dict)
    items = list(outputExpressions.items()) if is_struct else [(f"y{k + 1}", e) for k, e in enumerate(outputExpressions)]
    out_info, srcs, shapes, names = [], [], [], []
    for nm, e in items:
        e = Tcalculus._lift(e)
        srcs.append(e.flat())
        shapes.append(e.shape)
        names.append(nm)
        out_info.append((nm, e.shape))
    code.declare_get("getOutputs", srcs, shapes, names)
    defines = minmax_defines(len(u), len(d), info["nG"], info["nF"], info["nGd"], info["nFd"], **options)
    problem = code.compile(defines)
    fac = info["factor"]
    from .ordering import DEFAULT as default_ordering
    gen_info = dict(ordering=(ordering if isinstance(ordering, str) else "custom") if ordering is not None else default_ordering,
                    nnzL=fac.nnz(), kkt=fac.n, n_instructions=sum(len(g) for g in problem.groups),
                    n_groups=len(problem.groups), n_slots=problem.n_slots, driver="ipmPDminmax_CSsolver")
    return GeneratedSolver(classname, problem, par_info, var_info, out_info, is_struct, defines, gen_info)
