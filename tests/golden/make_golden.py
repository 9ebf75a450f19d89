"""Generates the golden fixtures of this directory from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

For every BASELINE problem a small seeded batch is solved by the reference oracle of
oracle/oracle.py kind="ref": the reference's unmodified lib/ipmPD_CSsolver.c driving code
printed by the reference's own writeCinstructionsC (lib/csparse/instructionsTableUTHash.c),
compiled with gcc -O1.  Inputs (C-ABI layout [batch, n]) and results (status, iter, outputs)
are stored as <name>.npz; tests compare the restated oracle (port) on CPU and the CUDA path on
the GPU box against them (neither needs /root/reference at test time).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = dict(sls=24, mpc_dcmotor=24, l1l2=6, mpc_quadcopter=6, isa_probe=24,
             matrix_game=24, qsaddle=24, qsaddle_latent=24)     # ipmPDeq_CSsolver (examples/TCgames.m)
SEED = 2024


def main():
    from tenscalc_b200 import problems
    from oracle.oracle import Oracle, pack_inputs
    only = sys.argv[1:]
    for name, B in CASES.items():
        if only and name not in only:
            continue
        gen, args = problems.BUILDERS[name]()
        P, V = problems.synthetic_inputs(name, gen, B, seed=SEED)
        st, it, out, _ = Oracle(gen, "ref").solve_batch(P, V, args["mu0"], args["maxIter"])
        Pp, Vp = pack_inputs(gen, P, V, B)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), P=Pp, V=Vp, status=st, iter=it,
                            mu0=args["mu0"], maxIter=args["maxIter"], seed=SEED,
                            **{"out_" + k: v for k, v in out.items()})
        print(name, "status", np.unique(st), "iter", it.min(), it.max())


if __name__ == "__main__":
    main()
