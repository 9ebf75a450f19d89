"""Problem definitions of the BASELINE configurations, transcribed from the reference's
example scripts into the Python front end (same variables, constraints, options).

* ``sls``            examples/sls.m:136-166 (box-constrained least squares; README.md:166-187)
* ``l1l2``           examples/l1l2estimationCS.m:137-183
* ``mpc_dcmotor``    examples/mpcmhe/mpc_dcmotor.m:52-142
* ``mpc_quadcopter`` examples/mpcmhe/mpc_quadcopter.m:86-182

Each builder returns a :class:`~tenscalc_b200.optimize.GeneratedSolver` plus the solve
arguments the script uses (``mu0``, ``maxIter``).  ``synthetic_inputs`` draws the
seeded parameter/initial-value batches described in SURVEY.md section 8(d).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import tensor as T
from .optimize import GeneratedSolver, cmex2optimizeCS
from .tensor import (Tconstant, Tones, Tvariable, Tzeros, horzcat, norm2, repmat, reshape, sqrt,
                     tmax, tmin, tsCross, tsDerivative, tsDerivative2, tsIntegral, tsum, vertcat)


def sls(N: int = 100, n: int = 8, **opts) -> Tuple[GeneratedSolver, dict]:
    T.clear()
    A = Tvariable("A", [N, n])
    b = Tvariable("b", N)
    x = Tvariable("x", n)
    y = A @ x - b
    J = norm2(y) / N
    gen = cmex2optimizeCS(classname="tmp_minslsc2", objective=J, optimizationVariables=[x],
                          constraints=[x >= 0, x <= .05], outputExpressions=[J, x],
                          parameters=[A, b], **opts)
    return gen, dict(mu0=1.0, maxIter=30)


def l1l2(N: int = 200, **opts) -> Tuple[GeneratedSolver, dict]:
    T.clear()
    measurement = Tvariable("measurement", N)
    dt1 = Tvariable("dt1", N - 1)
    weight2acceleration = Tvariable("weight2acceleration", ())
    position = Tvariable("position", N)
    weight1acceleration = Tvariable("weight1acceleration", ())
    weight1noise = Tvariable("weight1noise", ())
    noise1 = Tvariable("noise1", N)
    acceleration1 = Tvariable("acceleration1", N - 2)
    noise1abs = Tvariable("noise1abs", N)
    acceleration1abs = Tvariable("acceleration1abs", N - 2)

    velocity = (position[1:] - position[:-1]) * dt1
    acceleration = (velocity[1:] - velocity[:-1]) * dt1[:-1]
    noise2 = measurement - position - noise1
    acceleration2 = acceleration - acceleration1
    J = (norm2(noise2) + weight2acceleration * norm2(acceleration2)
         + weight1noise * tsum(noise1abs, 0) + weight1acceleration * tsum(acceleration1abs, 0))
    o = dict(scaleCost=1, desiredDualityGap=1e-3)
    o.update(opts)
    gen = cmex2optimizeCS(
        classname="tmpC_minl1l2", objective=J,
        optimizationVariables=[position, noise1, acceleration1, noise1abs, acceleration1abs],
        constraints=[noise1 <= noise1abs, noise1 >= -noise1abs,
                     acceleration1 <= acceleration1abs, acceleration1 >= -acceleration1abs],
        outputExpressions=[J, position, noise1, acceleration1, noise1abs, acceleration1abs],
        parameters=[measurement, dt1, weight1acceleration, weight2acceleration, weight1noise], **o)
    return gen, dict(mu0=1.0, maxIter=100)


def mpc_dcmotor(T_: int = 30, **opts) -> Tuple[GeneratedSolver, dict]:
    T.clear()
    nX, nU = 2, 1
    Ts = Tvariable("Ts", ())
    x = Tvariable("x", [nX, T_])
    u = Tvariable("u", [nU, T_ - 1])
    xinit = Tvariable("xinit", [nX, 1])
    p = Tvariable("p", ())
    k = Tvariable("k", ())
    A = vertcat(horzcat(Tzeros(1, 1), Tones(1, 1)), horzcat(Tzeros(1, 1), reshape(p, (1, 1))))
    B = vertcat(Tzeros(1, 1), reshape(k, (1, 1)))
    dynamics = [x[:, 1:] == x[:, :-1] + Ts * (A @ x[:, :-1] + B @ u),
                x[:, 0:1] == xinit]
    min_x = Tvariable("min_x", [nX, 1])
    max_x = Tvariable("max_x", [nX, 1])
    min_u = Tvariable("min_u", [nU, 1])
    max_u = Tvariable("max_u", [nU, 1])
    constraints = [x[:, 1:] >= repmat(min_x, [1, T_ - 1]),
                   x[:, 1:] <= repmat(max_x, [1, T_ - 1]),
                   u >= repmat(min_u, [1, T_ - 1]),
                   u <= repmat(max_u, [1, T_ - 1])]
    ref = Tvariable("ref", [1, T_])
    lambda_u = Tvariable("lambda_u", ())
    Jx2 = tsIntegral(tsum((x[0:1, :] - ref) ** 2, 0), Ts)
    Ju2 = tsIntegral(tsum(u ** 2, 0), Ts)
    J = Jx2 + lambda_u * Ju2
    uWarm = horzcat(u[:, 1:], Tzeros(nU, 1))
    xWarm = x[:, list(range(1, T_)) + [T_ - 1]]
    xWarm = tmax(xWarm, repmat(min_x + .05 * (max_x - min_x), [1, T_]))
    xWarm = tmin(xWarm, repmat(max_x - .05 * (max_x - min_x), [1, T_]))
    uWarm = tmax(uWarm, repmat(min_u + .05 * (max_u - min_u), [1, T_ - 1]))
    uWarm = tmin(uWarm, repmat(max_u - .05 * (max_u - min_u), [1, T_ - 1]))
    outputs = dict(J=J, Jx2=Jx2, Ju2=Ju2, t=Tconstant(np.arange(T_)) * Ts, u=u, x=x, ref=ref,
                   xWarm=xWarm, uWarm=uWarm)
    gen = cmex2optimizeCS(classname="tmp_mpc_dcmotor", objective=J, optimizationVariables=[u, x],
                          constraints=dynamics + constraints, outputExpressions=outputs,
                          parameters=[Ts, p, k, xinit, ref, min_x, max_x, min_u, max_u, lambda_u],
                          addEye2Hessian=True, adjustAddEye2Hessian=True, scaleInequalities=True, **opts)
    return gen, dict(mu0=1e-3, maxIter=100)


def mpc_quadcopter(T_: int = 100, **opts) -> Tuple[GeneratedSolver, dict]:
    T.clear()
    Ts = Tvariable("Ts", ())
    p = Tvariable("p", [3, T_])
    u = Tvariable("u", [3, T_])
    pinit = Tvariable("pinit", [3, 1])
    vinit = Tvariable("vinit", [3, 1])
    b_drag = Tvariable("b_drag", ())
    g = Tconstant([[0], [0], [9.8]])
    min_thrust = Tvariable("min_thrust", ())
    max_thrust = Tvariable("max_thrust", ())
    min_altitude = Tvariable("min_altitude", ())
    v = tsDerivative(p, Ts)
    a = tsDerivative2(p, Ts)
    dynamics = [a == -b_drag * v + repmat(g, [1, T_]) + u,
                p[:, 0:1] == pinit,
                v[:, 0:1] == vinit]
    m_thrust = sqrt(tsum(u ** 2, 0))                       # size [T]
    R_ib3 = u / repmat(reshape(m_thrust, [1, T_]), [3, 1])
    R_ib2 = tsCross(R_ib3, vertcat(Tzeros(2, T_), Tones(1, T_)))
    R_ib1 = tsCross(R_ib2, R_ib3)
    positive2 = Tvariable("positive2", [T_])
    constraints = [m_thrust >= min_thrust,
                   max_thrust - m_thrust == positive2,
                   positive2 > 0,
                   p[2, :] < -min_altitude]
    pdesired = Tvariable("pdesired", [3, 1])
    vdesired = Tvariable("vdesired", [3, 1])
    lambda_v = Tvariable("lambda_v", ())
    lambda_thrust = Tvariable("lambda_thrust", ())
    Jp2 = tsIntegral(tsum((p - repmat(pdesired, [1, T_])) ** 2, 0), Ts)
    Jv2 = tsIntegral(tsum((v - repmat(vdesired, [1, T_])) ** 2, 0), Ts)
    Jthrust = tsIntegral(m_thrust, Ts)
    J = Jp2 + lambda_v * Jv2 + lambda_thrust * Jthrust
    outputs = dict(J=J, Jp2=Jp2, Jv2=Jv2, Jthrust=Jthrust, t=Tconstant(np.arange(T_)) * Ts, u=u,
                   m_thrust=m_thrust, p=p, v=v, a=a, R_ib1=R_ib1, R_ib2=R_ib2, R_ib3=R_ib3,
                   positive2=positive2)
    gen = cmex2optimizeCS(classname="tmp_quadcopter", objective=J,
                          optimizationVariables=[p, u, positive2],
                          constraints=dynamics + constraints, outputExpressions=outputs,
                          parameters=[Ts, pinit, vinit, pdesired, vdesired, b_drag, min_thrust,
                                      max_thrust, min_altitude, lambda_v, lambda_thrust],
                          addEye2Hessian=True, adjustAddEye2Hessian=True, scaleInequalities=True, **opts)
    return gen, dict(mu0=1e-3, maxIter=1000)


def _cw_sin(x):
    """sin as a user-defined componentwise function with first and second derivative (Tcalculus.m:1851-1853)"""
    import math
    from .tensor import componentwise
    return componentwise(x, math.sin, "sin(%s)",
                         lambda z: componentwise(z, math.cos, "cos(%s)", lambda w: -componentwise(w, math.sin, "sin(%s)")))


def isa_probe(N: int = 60, cw: bool = False, **opts) -> Tuple[GeneratedSolver, dict]:
    """Not a reference example: a small smooth bounded problem whose cost, gradient and Hessian together
    contain every unary opcode of the scalar ISA (lib/instructionTypes.m:8-117, C forms
    instructionsTableUTHash.c:1347-1539) and sums of products longer than one emitted line (:1019-1026,
    ``split_long`` off), so that the parity tests execute all of them on the device against the reference's
    own emitter.  Non-differentiable opcodes (abs, sign, round, ceil, floor, relu, heaviside) act on
    parameters only.  ``cw=True`` writes the sine term as a user-defined ``componentwise`` function with its two
    derivatives (the example of lib/@Tcalculus/Tcalculus.m:1851-1853: ``I_componentwise`` instructions carrying
    the C texts ``sin(%s)`` / ``cos(%s)``) -- the same problem, so the same iterates."""
    from .tensor import _compose, abs_, atan, cos, cube, exp, log, relu, sin, sqr, srelu
    from . import itable as it
    T.clear()
    T.table().split_long = False
    T.table().sqrt_by_division = False       # differentiate sqrt to I_Dsqrt / I_DDsqrt like the reference's table
    n = 10
    x = Tvariable("x", n)
    c = Tvariable("c", n)
    A = Tvariable("A", [N, n])
    b = Tvariable("b", N)
    X = [x[i] for i in range(n)]
    C = [c[i] for i in range(n)]
    step = (relu(C[9] - 1) + _compose(C[0] - 1, it.I_heaviside) + .1 * abs_(C[1] - 1)
            + .1 * _compose(C[2] - 1, it.I_sign) + .1 * _compose(3 * C[3], it.I_round)
            + .1 * _compose(3 * C[4], it.I_ceil) + .1 * _compose(3 * C[5], it.I_floor))
    terms = [exp(X[0]) - C[0] * X[0],
             srelu(X[1]) + sqr(X[1] - C[1]),
             X[2] * atan(X[2]) + sqr(X[2] - C[2]),
             C[3] * X[3] - log(X[3] + 3),
             cube(X[4]) + 10 * sqr(X[4] - C[4]),
             C[5] * X[5] - cos(X[5]),
             (_cw_sin(X[6]) if cw else sin(X[6])) + 2 * sqr(X[6]),
             C[7] * X[7] - sqrt(X[7] + 3),
             1 / (X[8] + 3) + sqr(X[8]),
             sqr(X[9]) * (1 + step)]
    J = norm2(A @ x - b) / N
    for t in terms:
        J = J + t
    o = dict(opts)
    gen = cmex2optimizeCS(classname="tmp_isaprobe", objective=J, optimizationVariables=[x],
                          constraints=[x >= -1, x <= 1], outputExpressions=[J, x, step],
                          parameters=[c, A, b], **o)
    return gen, dict(mu0=1.0, maxIter=60)


# ------------------------------------------------------------------------------------------------
# zero-sum games of examples/TCgames.m, generated with cmex2equilibriumLatentCS -> ipmPDeq_CSsolver
# ------------------------------------------------------------------------------------------------
def matrix_game(N1: int = 10, N2: int = 6, **opts) -> Tuple[GeneratedSolver, dict]:
    """mixed saddle point of a matrix game, examples/TCgames.m:29-98 (N1=50, N2=20 there)"""
    from .optimize import cmex2equilibriumLatentCS
    T.clear()
    A1 = Tvariable("A1", [N1, N2])
    u = Tvariable("u", N1)
    d = Tvariable("d", N2)
    J1 = (u @ A1) @ d
    gen = cmex2equilibriumLatentCS(classname="tmp_matrixsaddle", P1objective=J1, P2objective=-J1,
                                   P1optimizationVariables=[u], P2optimizationVariables=[d],
                                   P1constraints=[tsum(u, 0) == 1, u >= 0], P2constraints=[tsum(d, 0) == 1, d >= 0],
                                   outputExpressions=[u, d, J1], parameters=[A1], **opts)
    return gen, dict(mu0=1.0, maxIter=30)


def qsaddle(N: int = 20, latent: bool = False, **opts) -> Tuple[GeneratedSolver, dict]:
    """saddle point of a quadratic zero-sum cost with bounded policies, examples/TCgames.m:104-147, and the same
    problem written with a latent (state) variable x = u + d, :155-214.  The constant 2 of the script is the
    parameter ``target`` here, so that a batch holds different instances."""
    from .optimize import cmex2equilibriumLatentCS
    T.clear()
    u = Tvariable("u", N)
    d = Tvariable("d", N)
    target = Tvariable("target", N)
    if latent:
        x = Tvariable("x", N)
        Ju = norm2(x - target) - 10 * norm2(d)
        gen = cmex2equilibriumLatentCS(classname="tmp_qsaddlestate", P1objective=Ju, P2objective=-Ju,
                                       P1optimizationVariables=[u], P2optimizationVariables=[d], latentVariables=[x],
                                       P1constraints=[u >= -1, u <= 1], P2constraints=[d >= -1, d <= 1],
                                       latentConstraints=[x == u + d], outputExpressions=[u, d, x, Ju],
                                       parameters=[target], **opts)
    else:
        Ju = norm2(u + d - target) - 10 * norm2(d)
        gen = cmex2equilibriumLatentCS(classname="tmp_qsaddle", P1objective=Ju, P2objective=-Ju,
                                       P1optimizationVariables=[u], P2optimizationVariables=[d],
                                       P1constraints=[u >= -1, u <= 1], P2constraints=[d >= -1, d <= 1],
                                       outputExpressions=[u, d, Ju], parameters=[target], **opts)
    return gen, dict(mu0=1.0, maxIter=30)


def qsaddle_latent(N: int = 20, **opts) -> Tuple[GeneratedSolver, dict]:
    return qsaddle(N, latent=True, **opts)


def minmax_box(N: int = 4, latent: bool = False, **opts) -> Tuple[GeneratedSolver, dict]:
    """cases 5 and 5.5 of examples/minmaxTest.m:84-103, vectorised over N coordinates and with the offset as a
    parameter: min_u max_d |u+d+a|^2 - 2|d+2|^2, -2<=u<=2, -1<=d<=1 (5.5: through x == u+d, x a maximiser variable);
    class2minmaxCS options of the script (:106-118): addEye2Hessian, adjustAddEye2Hessian, no scaling"""
    from .optimize import cmex2minmaxCS
    T.clear()
    u = Tvariable("u", N)
    d = Tvariable("d", N)
    a = Tvariable("a", N)
    o = dict(scaleCost=0, addEye2Hessian=True, adjustAddEye2Hessian=True, scaleInequalities=False)
    o.update(opts)
    if latent:
        x = Tvariable("x", N)
        J = norm2(x) - 2 * norm2(d + 2)
        gen = cmex2minmaxCS(classname="tmp_minmaxstate", objective=J, minOptimizationVariables=[u],
                            maxOptimizationVariables=[d, x], minConstraints=[u >= -2, u <= 2],
                            maxConstraints=[d >= -1, d <= 1, x == u + d + a], outputExpressions=dict(u=u, d=d, x=x, J=J),
                            parameters=[a], **o)
    else:
        J = norm2(u + d + a) - 2 * norm2(d + 2)
        gen = cmex2minmaxCS(classname="tmp_minmax", objective=J, minOptimizationVariables=[u],
                            maxOptimizationVariables=[d], minConstraints=[u >= -2, u <= 2],
                            maxConstraints=[d >= -1, d <= 1], outputExpressions=dict(u=u, d=d, J=J), parameters=[a], **o)
    return gen, dict(mu0=1.0, maxIter=30, addEye2Hessian=(1e-9, 1e-9, 1e-9))


def minmax_box_latent(N: int = 4, **opts) -> Tuple[GeneratedSolver, dict]:
    return minmax_box(N, latent=True, **opts)


BUILDERS = dict(sls=sls, l1l2=l1l2, minmax_box=minmax_box, minmax_box_latent=minmax_box_latent, mpc_dcmotor=mpc_dcmotor, mpc_quadcopter=mpc_quadcopter, isa_probe=isa_probe,
                matrix_game=matrix_game, qsaddle=qsaddle, qsaddle_latent=qsaddle_latent)


# ------------------------------------------------------------------------------------------------
# synthetic batches; arrays carry the batch as the LAST dimension
# ------------------------------------------------------------------------------------------------
def synthetic_inputs(name: str, gen: GeneratedSolver, batch: int, seed: int = 0) -> Tuple[Dict, Dict]:
    rng = np.random.Generator(np.random.PCG64(seed))
    B = batch
    if name == "sls":
        N, n = dict(gen.parameters)["A"]
        P = dict(A=rng.random((N, n, B)), b=rng.random((N, B)))
        V = dict(x=.02 * rng.random((n, B)))
    elif name == "l1l2":
        N = dict(gen.parameters)["measurement"][0]
        t = np.cumsum(np.ceil(rng.random((N, B)) * 3), axis=0)     # irregular sample times
        true_pos = 5 * np.sin(np.abs(t - 100) / 5)
        meas = np.round(true_pos + rng.standard_normal((N, B)))
        outl = rng.random((N, B)) < .1
        meas[outl] = np.round(10 * rng.standard_normal(int(outl.sum())))
        P = dict(measurement=meas, dt1=1. / np.diff(t, axis=0),
                 weight1acceleration=np.ones(B), weight2acceleration=10 * np.ones(B),
                 weight1noise=np.ones(B))
        V = dict(position=rng.random((N, B)), noise1=rng.random((N, B)),
                 acceleration1=rng.random((N - 2, B)), noise1abs=10 * np.ones((N, B)),
                 acceleration1abs=10 * np.ones((N - 2, B)))
    elif name == "mpc_dcmotor":
        T_ = dict(gen.parameters)["ref"][1]
        Ts = .1
        xinit = np.stack([rng.uniform(-.35, .35, B), rng.uniform(-.25, .25, B)])[:, None, :]
        t0 = rng.uniform(0, 4 * np.pi, B)
        tt = t0[None, :] + Ts * np.arange(T_)[:, None]
        ref = (-.35 * np.sign(np.sin(.5 * tt)))[None, :, :]
        ones = np.ones(B)
        P = dict(Ts=Ts * ones, p=-2 * ones, k=ones, xinit=xinit, ref=ref,
                 min_x=np.array([-.4, -.3])[:, None, None] * ones, max_x=np.array([.4, .3])[:, None, None] * ones,
                 min_u=-ones[None, None, :], max_u=ones[None, None, :], lambda_u=ones / 50)
        V = dict(x=xinit + .01 * rng.random((2, T_, B)), u=.01 * rng.random((1, T_ - 1, B)))
    elif name == "mpc_quadcopter":
        T_ = dict(gen.variables)["p"][1]
        Ts, b_drag = .02, .1
        ones = np.ones(B)
        pinit = rng.normal(0, .5, (3, B))
        pinit[2] = np.minimum(pinit[2], -.2)
        vinit = rng.normal(0, .3, (3, B))
        pdes = np.array([0, 5., -2.5])[:, None] * ones
        P = dict(Ts=Ts * ones, pinit=pinit[:, None, :], vinit=vinit[:, None, :],
                 pdesired=pdes[:, None, :], vdesired=np.zeros((3, 1, B)), b_drag=b_drag * ones,
                 min_thrust=5 * ones, max_thrust=20 * ones, min_altitude=-.1 * ones,
                 lambda_v=.05 * ones, lambda_thrust=.05 * ones)
        # primal initialisation (mpc_quadcopter.m:234-247): cubic through p(t1), p(t1)+Ts v,
        # p(t1)+2Ts v and the partial goal at t_T (not-a-knot spline over 4 points = the
        # interpolating cubic)
        pfinal = pinit + (pdes - pinit) / 10
        tk = np.array([1, 2, 3, T_]) * Ts
        tq = np.arange(1, T_ + 1) * Ts
        vals = np.stack([pinit, pinit + Ts * vinit, pinit + 2 * Ts * vinit, pfinal])      # [4,3,B]
        Vm = np.vander(tk, 4)
        coef = np.linalg.solve(Vm, vals.reshape(4, -1))
        p = (np.vander(tq, 4) @ coef).reshape(T_, 3, B).transpose(1, 0, 2)
        v = np.empty_like(p)
        v[:, 0] = (-1.5 * p[:, 0] + 2 * p[:, 1] - .5 * p[:, 2]) / Ts
        v[:, 1:-1] = -.5 / Ts * (p[:, :-2] - p[:, 2:])
        v[:, -1] = (.5 * p[:, -3] - 2 * p[:, -2] + 1.5 * p[:, -1]) / Ts
        a = np.empty_like(p)
        a[:, 1:-1] = (p[:, :-2] - 2 * p[:, 1:-1] + p[:, 2:]) / Ts ** 2
        a[:, 0] = a[:, 1]
        a[:, -1] = a[:, -2]
        u = a + b_drag * v - np.array([0, 0, 9.8])[:, None, None]
        V = dict(p=p, u=u, positive2=np.ones((T_, B)))
    elif name == "matrix_game":
        N1, N2 = dict(gen.parameters)["A1"]
        P = dict(A1=rng.random((N1, N2, B)))
        V = dict(u=np.full((N1, B), 1. / N1), d=np.full((N2, B), 1. / N2))
    elif name in ("qsaddle", "qsaddle_latent"):
        N = dict(gen.parameters)["target"][0]
        P = dict(target=rng.uniform(0, 3, (N, B)))
        u0 = rng.uniform(-.5, .5, (N, B))
        d0 = rng.uniform(-.5, .5, (N, B))
        V = dict(u=u0, d=d0)
        if name == "qsaddle_latent":
            V["x"] = u0 + d0
    elif name in ("minmax_box", "minmax_box_latent"):
        N = dict(gen.parameters)["a"][0]
        P = dict(a=rng.uniform(-1, 3, (N, B)))
        V = dict(u=np.full((N, B), .1), d=np.full((N, B), .1))
        if name == "minmax_box_latent":
            V["x"] = V["u"] + V["d"] + P["a"]
    elif name == "isa_probe":
        N, n = dict(gen.parameters)["A"]
        P = dict(c=rng.uniform(.5, 1.5, (n, B)), A=rng.random((N, n, B)), b=rng.random((N, B)))
        V = dict(x=.1 * rng.random((n, B)))
    else:
        raise KeyError(name)
    return P, V
