"""CPU ORACLE (test infrastructure): build every per-problem oracle library used by the tests
and by bench.py's cpu_baseline / --impl reference legs.  Run here (needs /root/reference for
kind="ref"); the results in oracle/_ref/ travel to the GPU box."""
import sys
import time
from concurrent.futures import ProcessPoolExecutor

AFFINE = {"algorithm": "ipmPD_CS", "skipAffine": False}           # Mehrotra predictor (ipmPD_CSsolver.c:576-665, ipmPD_CS.m)
AFFINE_SMALL = {**AFFINE, "smallerNewtonMatrix": True}            # + reduced Newton matrix (ipmPD_CS.m:242-351)
CONFIGS = [("sls", {}), ("mpc_dcmotor", {}), ("l1l2", {}), ("mpc_quadcopter", {}), ("isa_probe", {}),
           ("mpc_dcmotor", {"allowSave": True}),                   # saveWW__ events of the drivers (ipmPD_CSsolver.c:330-332,561-566,...)
           ("isa_probe", {"cw": True}),                            # user-defined componentwise function (I_componentwise)
           ("mpc_quadcopter", {"T_": 20, "useInertia": True}),    # inertia-controlled regularisation (ipmPD_CSsolver.c:408-519)
           ("sls", AFFINE), ("mpc_dcmotor", AFFINE), ("mpc_dcmotor", AFFINE_SMALL),
           ("mpc_quadcopter", {"T_": 20, **AFFINE})]
# ipmPDeq_CSsolver (equilibria): the reference's own driver is the only oracle (kind "ref")
EQ_CONFIGS = [("matrix_game", {}), ("qsaddle", {}), ("qsaddle_latent", {}), ("qsaddle", {"useLDL": True}),
              ("matrix_game", {"skipAffine": False}),
              ("qsaddle", {"addEye2Hessian": True, "adjustAddEye2Hessian": True}),   # regulariser control of the eq driver (:318-374)
              # reduced Newton matrix of the equilibrium builder (ipmPDeqlat_CS.m:314-561): dlambda eliminated
              ("qsaddle", {"smallerNewtonMatrix": True}), ("qsaddle_latent", {"smallerNewtonMatrix": True}),
              ("matrix_game", {"smallerNewtonMatrix": True, "skipAffine": False})]


# extra builds of the benchmark workload: the toolbox's default -O0 and -O2 beside the headline -O1
# (bench.py cpu_baseline.other_opt_levels), and the libm-sensitivity probe (bench.py parity.libm_noise_floor)
EXTRA = [("mpc_quadcopter", {}, "-O0", ""), ("mpc_quadcopter", {}, "-O2", ""), ("mpc_quadcopter", {}, "-O1", "crpow")]


# ipmPDminmax_CSsolver (min-max): PARITY UNPINNED port of the MATLAB solver, both kinds include it
MM_CONFIGS = [("minmax_box", {}), ("minmax_box_latent", {}), ("minmax_box", {"N": 16})]


def _one(args):
    name, kw, kind, opt, variant = args
    from tenscalc_b200 import problems
    from oracle.oracle import Oracle
    gen, _ = problems.BUILDERS[name](**kw)
    t = time.time()
    so = Oracle(gen, kind, opt, variant=variant).build()
    return name, kind + opt + variant, so, time.time() - t


# dense least squares (config 5, oracle/dense_ls.c): (N, n, lo, hi) of tests/test_dense.py and bench.py --workload dense_ls
DENSE = [(100, 8, 0.0, .05), (400, 32, -1.0, 1.0), (4000, 256, -1.0, 1.0)]


def build_all(kinds=("ref", "port"), configs=CONFIGS, workers=4, extra=True):
    from oracle.dense_oracle import DenseOracle
    for N, n, lo, hi in DENSE:
        for k in kinds:
            print("oracle dense_ls", N, n, k, DenseOracle(N, n, lo, hi, k).build())
    jobs = [(n, kw, k, "-O1", "") for n, kw in configs for k in kinds]
    if "ref" in kinds:
        jobs += [(n, kw, "ref", "-O1", "") for n, kw in EQ_CONFIGS + MM_CONFIGS]
    if extra:
        jobs += [(n, kw, kinds[0], opt, var) for n, kw, opt, var in EXTRA]
    with ProcessPoolExecutor(workers) as ex:
        for name, kind, so, dt in ex.map(_one, jobs):
            print(f"oracle {name:16s} {kind:12s} {dt:6.1f}s  {so}")


if __name__ == "__main__":
    kinds = tuple(sys.argv[1:]) or ("ref",)
    build_all(kinds)
