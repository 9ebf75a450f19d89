"""Dense least-squares path (BASELINE config 5: N=4000, n=256, batch 4096).

CPU (oracle pinning): ``oracle/dense_ls.c`` -- the reference's own ipmPD_CSsolver.c around hand-written dense
versions of the functions ipmPD_CStimesLambda.m declares -- is checked, built for N=100, n=8, against the solver
the reference GENERATES for examples/sls.m (oracle/_ref: reference emitter + reference driver):
every ``*__`` function agrees to rounding along a solver trajectory, and whole solves give the same status, the
same iteration count +-1 and the same solution.

GPU (-m gpu): the CUDA kernels through the C ABI (tc_dense_*) against that oracle: identical status, iterations
within +-1, solution and cost within 1e-8 relative (BASELINE.json north_star) -- at the pinning size, at a
middle size and at the full config-5 size.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, generated
from tenscalc_b200 import problems

HAVE_REF = os.path.isdir("/root/reference")


def _dense_oracle(N, n, lo, hi):
    from oracle.dense_oracle import DenseOracle
    o = DenseOracle(N, n, lo, hi, "ref")
    if not os.path.exists(o.so) and not HAVE_REF:
        o = DenseOracle(N, n, lo, hi, "port")
    return o


def test_dense_oracle_functions_match_reference_generated_functions(tmp_path):
    """function-level pin: getNorminf_Grad__, getGapMinFMinLambda__, getCurvature__, getDirectionError__,
    getMaxAlphas_s__, getMinF_s__ of the hand-written dense oracle against the reference-emitted code of the
    generated sls solver, called in the driver's order along three Newton steps"""
    from oracle.oracle import Oracle
    gen, args = generated("sls")
    o = Oracle(gen, "ref")
    if not os.path.exists(o.so[:-3] + ".c") and not HAVE_REF:
        pytest.skip("needs the generated C of the reference oracle (build container)")
    o.build()
    g_so, d_so = str(tmp_path / "gen.so"), str(tmp_path / "dense.so")
    subprocess.check_call(["gcc", "-O1", "-w", "-shared", "-fPIC", o.so[:-3] + ".c", "-o", g_so, "-lm"])
    drv = [] if HAVE_REF else ["-DDRIVER_PORT"]
    subprocess.check_call(["gcc", "-O1", "-w", "-shared", "-fPIC", "-DDN=100", "-DDn=8", "-DLO=0.0", "-DHI=0.05"] + drv +
                          [os.path.join(ROOT, "oracle", "dense_ls.c"), "-o", d_so, "-lm"])
    g, d = ctypes.CDLL(g_so), ctypes.CDLL(d_so)
    dp = ctypes.POINTER(ctypes.c_double)
    P, V = problems.synthetic_inputs("sls", gen, 3, seed=4)

    def call(lib, name, *vals):
        a = [ctypes.c_double(v) for v in vals]
        getattr(lib, name)(*[ctypes.byref(x) for x in a])
        return np.array([x.value for x in a])

    for k in range(3):
        A = np.ascontiguousarray(P["A"][:, :, k].reshape(-1, order="F"))
        b = np.ascontiguousarray(P["b"][:, k])
        x0 = np.ascontiguousarray(V["x"][:, k])
        for lib, pre in ((g, "tmp_minslsc2_set_"), (d, "dense_set_")):
            for nm, arr in (("A", A), ("b", b), ("x", x0)):
                getattr(lib, pre + nm)(arr.ctypes.data_as(dp))
            call(lib, "setAddEye2HessianU__", 1e-9)
            call(lib, "setAddEye2HessianEq__", 1e-9)
            lib.initPrimal__(); lib.scaleIneq__(); call(lib, "setMu__", 1.0); lib.initDualIneq__()
        for step in range(3):
            for name, nout in (("getNorminf_Grad__", 1), ("getGapMinFMinLambda__", 3), ("getCurvature__", 1),
                               ("getMaxAlphas_s__", 2)):
                a, c = call(g, name, *[0.0] * nout), call(d, name, *[0.0] * nout)
                assert np.allclose(a, c, rtol=1e-11, atol=0), (k, step, name, a, c)
            # the direction error is pure rounding noise (1e-13 on O(1e3) data): both must be negligible
            assert call(g, "getDirectionError__", 0.0)[0] < 1e-10 and call(d, "getDirectionError__", 0.0)[0] < 1e-10
            for lib in (g, d):
                call(lib, "setAlphaPrimal__", .5)
            a, c = call(g, "getMinF_s__", 0.0), call(d, "getMinF_s__", 0.0)
            assert np.allclose(a, c, rtol=1e-12, atol=0)
            for lib in (g, d):
                call(lib, "setAlphaDualIneq__", .5)
                lib.updatePrimalDual__()
                call(lib, "setMu__", .5 ** (step + 1))


def test_dense_oracle_solves_like_the_reference_generated_solver():
    from oracle.oracle import Oracle
    gen, args = generated("sls")
    o = Oracle(gen, "ref")
    if not os.path.exists(o.so) and not HAVE_REF:
        pytest.skip("compiled reference library not present")
    B = 64
    P, V = problems.synthetic_inputs("sls", gen, B, seed=4)
    st, it, out, _ = o.solve_batch(P, V, args["mu0"], args["maxIter"])
    st2, it2, J2, x2, _ = _dense_oracle(100, 8, 0.0, .05).solve_batch(P["A"], P["b"], V["x"], args["mu0"], args["maxIter"])
    assert np.array_equal(st, st2)
    assert np.abs(it.astype(int) - it2.astype(int)).max() <= 1
    # examples/sls.m's data (A ~ U(0,1)) is ill-conditioned: two roundings of the same arithmetic part by up to
    # 1e-6 at equal iteration counts, and a solve that stops one iteration earlier is 1e-4 away
    eq = it == it2
    assert eq.mean() > .8
    scale = np.abs(out["y2"]).max()
    assert np.abs(x2 - out["y2"])[:, eq].max() / scale < 1e-5
    assert np.abs(x2 - out["y2"]).max() / scale < 1e-3
    assert np.abs(J2 - out["y1"]).max() / np.abs(out["y1"]).max() < 1e-3


def test_dense_trailing_block_detection():
    """the symbolic LDL' of the box-constrained least-squares KKT matrix ends in a completely dense n x n block
    (what routes such a problem to the dense kernels); the banded MPC factor does not"""
    gen, _ = generated("sls")
    assert gen.info["dense_block"] >= 8
    gen2, _ = generated("mpc_dcmotor")
    assert gen2.info["dense_block"] < 16


def test_flop_model_matches_survey_anchor():
    """dense LDL 256^3/3 = 5.6e6 flop per iteration, SYRK 2.6e8 flop once"""
    from tenscalc_b200.dense import gram_flops, newton_step_flops
    assert 2.5e8 < gram_flops(4000, 256) < 2.8e8
    assert 5.5e6 < newton_step_flops(256) < 6.5e6


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("N,n,B,lo,hi,kind", [(100, 8, 256, 0.0, .05, "sls"), (400, 32, 128, -1.0, 1.0, "c5"),
                                               (4000, 256, 16, -1.0, 1.0, "c5")])
def test_dense_gpu_parity(N, n, B, lo, hi, kind):
    from tenscalc_b200.dense import DenseLSSolver, synthetic_inputs
    if kind == "sls":
        gen, args = generated("sls")
        P, V = problems.synthetic_inputs("sls", gen, B, seed=9)
        A, b, x0, mu0, maxIter = P["A"], P["b"], V["x"], args["mu0"], args["maxIter"]
    else:
        A, b, x0 = synthetic_inputs(N, n, B, seed=9)
        mu0, maxIter = 1.0, 50
    st_o, it_o, J_o, x_o, _ = _dense_oracle(N, n, lo, hi).solve_batch(A, b, x0, mu0, maxIter)
    obj = DenseLSSolver(N, n, B, lo, hi)
    obj.setP_A(A); obj.setP_b(b); obj.setV_x(x0)
    st, it, t = obj.solve(mu0, maxIter, -1)
    J, x = obj.getOutputs()
    launches, dmma = obj.counters()
    obj.close()
    assert np.array_equal(st, st_o), f"status mismatch at {np.nonzero(st != st_o)[0][:10]}"
    dit = np.abs(it.astype(int) - it_o.astype(int))
    assert dit.max() <= 1
    ok = (st_o == 0) & (dit == 0)
    assert ok.mean() > (.8 if kind == "sls" else .99)
    tol = 1e-5 if kind == "sls" else 1e-8        # sls data is ill-conditioned (see the CPU test above)
    assert np.abs(x - x_o)[:, ok].max() / np.abs(x_o).max() <= tol
    assert (np.abs(J - J_o)[ok] / np.abs(J_o[ok])).max() <= tol
    assert launches >= 2 and (dmma > 0 or n <= 32)
    print(f"dense N={N} n={n} B={B}: iters {it.min()}..{it.max()} equal={int((dit == 0).sum())}/{B} "
          f"x err {np.abs(x - x_o)[:, ok].max() / np.abs(x_o).max():.2e} batch time {t * 1e3:.2f} ms dmma flops {dmma}")


@pytest.mark.gpu
def test_dense_set_b_alone_recomputes_and_bad_sizes_are_rejected():
    from tenscalc_b200.dense import DenseLSSolver, synthetic_inputs
    from tenscalc_b200.runtime import TenscalcError
    N, n, B = 400, 32, 32
    A, b, x0 = synthetic_inputs(N, n, B, seed=2)
    o = _dense_oracle(N, n, -1.0, 1.0)
    obj = DenseLSSolver(N, n, B, -1.0, 1.0)
    obj.setP_A(A); obj.setP_b(b); obj.setV_x(x0)
    obj.solve(1.0, 50, -1)
    b2 = b[:, ::-1].copy()
    obj.setP_b(b2)                                   # only b changes: the Gram groups are re-evaluated lazily
    st, it, _ = obj.solve(1.0, 50, -1)
    J, x = obj.getOutputs()
    st_o, it_o, J_o, x_o, _ = o.solve_batch(A, b2, x0, 1.0, 50)
    assert np.array_equal(st, st_o) and np.abs(it.astype(int) - it_o.astype(int)).max() <= 1
    same = it == it_o
    assert np.abs(x - x_o)[:, same].max() / np.abs(x_o).max() <= 1e-8
    obj.close()
    with pytest.raises(TenscalcError):
        DenseLSSolver(100, 12, 4, 0.0, 1.0)          # n must be 8 * 2^k <= 256
