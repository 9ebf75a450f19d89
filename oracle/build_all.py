"""CPU ORACLE (test infrastructure): build every per-problem oracle library used by the tests
and by bench.py's cpu_baseline / --impl reference legs.  Run here (needs /root/reference for
kind="ref"); the results in oracle/_ref/ travel to the GPU box."""
import sys
import time
from concurrent.futures import ProcessPoolExecutor

CONFIGS = [("sls", {}), ("mpc_dcmotor", {}), ("l1l2", {}), ("mpc_quadcopter", {}),
           ("mpc_quadcopter", {"T_": 20, "useInertia": True})]   # inertia-controlled regularisation (ipmPD_CSsolver.c:408-519)


def _one(args):
    name, kw, kind = args
    from tenscalc_b200 import problems
    from oracle.oracle import Oracle
    gen, _ = problems.BUILDERS[name](**kw)
    t = time.time()
    so = Oracle(gen, kind).build()
    return name, kind, so, time.time() - t


def build_all(kinds=("ref", "port"), configs=CONFIGS, workers=4):
    jobs = [(n, kw, k) for n, kw in configs for k in kinds]
    with ProcessPoolExecutor(workers) as ex:
        for name, kind, so, dt in ex.map(_one, jobs):
            print(f"oracle {name:16s} {kind:5s} {dt:6.1f}s  {so}")


if __name__ == "__main__":
    kinds = tuple(sys.argv[1:]) or ("ref",)
    build_all(kinds)
