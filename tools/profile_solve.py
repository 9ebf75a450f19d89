#!/usr/bin/env python
"""One short batched solve for ncu captures (profiles/): `--max-iter` bounds the Newton
iterations so that a full-occupancy launch replays in seconds.

    ncu --set full --clock-control none --import-source on -k regex:tc_solve -c 1 \
        -o gpurun_out/prof python tools/profile_solve.py --workload mpc_quadcopter --batch 75776 --max-iter 2
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mpc_quadcopter")
    ap.add_argument("--batch", type=int, default=75776)
    ap.add_argument("--max-iter", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--backend", default="compiled", choices=["compiled", "tape"])
    ap.add_argument("--lockstep", type=int, default=0, help="warps per CTA walking the code together (0: default)")
    ap.add_argument("--assume-global", type=int, default=0, help="1: __isGlobal assumption -> LDG/STG")
    ap.add_argument("--prefetch", type=int, default=-1, help="0: no next-chunk L2 prefetch (-1: default)")
    ap.add_argument("--resident", type=int, default=0, help="cap on resident CTAs per SM (0: register limit)")
    ap.add_argument("--driver", default="", help="warp | lanes (default: emitter default)")
    ap.add_argument("--tiles", type=int, default=0, help="instances per lane (0: default)")
    ap.add_argument("--merge", type=int, default=0, help="chunks per emitted function (0: default)")
    ap.add_argument("--divcall", type=int, default=-1, help="1: IEEE division out of line, 0: inline (-1: default)")
    ap.add_argument("--max-loads", type=int, default=0, help="scratchbook operands per chunk (0: default)")
    ap.add_argument("--warps-per-sm", type=int, default=0, help="resident warps per SM = register budget (0: default)")
    ap.add_argument("--mopts", default="", help="further emitter options, k=v,k=v (integers)")
    args = ap.parse_args()
    from tenscalc_b200 import problems
    gen, sargs = problems.BUILDERS[args.workload]()
    P, V = problems.synthetic_inputs(args.workload, gen, args.batch, seed=7)
    opts = {k: v for k, v in dict(lockstep=args.lockstep, max_loads=args.max_loads,
                                      warps_per_sm=args.warps_per_sm, merge=args.merge, tiles=args.tiles,
                                      assume_global=args.assume_global, resident=args.resident).items() if v}
    for kv in filter(None, args.mopts.split(",")):
        k, v = kv.split("=")
        opts[k] = int(v) if v.lstrip("-").isdigit() else v
    if args.driver:
        opts["driver"] = args.driver
    if args.prefetch >= 0:
        opts["prefetch"] = args.prefetch
    if args.divcall >= 0:
        opts["divcall"] = args.divcall
    obj = gen(args.batch, device=0, backend=args.backend, **opts)
    for n, _ in gen.parameters:
        getattr(obj, "setP_" + n)(P[n])
    for _ in range(args.repeat):
        for n, _ in gen.variables:
            getattr(obj, "setV_" + n)(V[n])
        t = time.time()
        st, it, sec = obj.solve(sargs["mu0"], args.max_iter or sargs["maxIter"], -1)
        print(f"{args.workload}[{args.backend} {opts}] B={args.batch} maxIter={args.max_iter}: kernel {sec * 1e3:.1f} ms, "
              f"wall {1e3 * (time.time() - t):.1f} ms, newton steps {int(it.sum())}, "
              f"status histogram {dict(zip(*np.unique(st, return_counts=True)))}, iter percentiles "
              f"50/90/99/99.9/max {[int(x) for x in np.percentile(it, [50, 90, 99, 99.9, 100])]}, "
              f"mean warp-max {it[:len(it) // 32 * 32].reshape(-1, 32).max(1).mean():.1f}")
    obj.close()


if __name__ == "__main__":
    main()
