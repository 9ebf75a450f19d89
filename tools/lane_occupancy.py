#!/usr/bin/env python
"""Lane occupancy of the solve kernel per dependency group: instance-executions (tc_group_counts)
against warp-executions (tc_group_warp_counts), weighted by the bytes each group issues.  Tells how
much of a batch's cost is idle lanes (warp-max iterations, re-factorisations for a few lanes)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mpc_quadcopter")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--driver", default="")
    args = ap.parse_args()
    from tenscalc_b200 import problems
    from tenscalc_b200.emit_cuda import emit_module
    gen, sargs = problems.BUILDERS[args.workload]()
    P, V = problems.synthetic_inputs(args.workload, gen, args.batch, seed=7)
    opts = dict(driver=args.driver) if args.driver else {}
    obj = gen(args.batch, device=0, **opts)
    for n, _ in gen.parameters:
        getattr(obj, "setP_" + n)(P[n])
    for n, _ in gen.variables:
        getattr(obj, "setV_" + n)(V[n])
    obj.group_counts(reset=True); obj.group_warp_counts(reset=True)
    st, it, sec = obj.solve(sargs["mu0"], sargs["maxIter"], -1)
    inst = obj.group_counts().astype(float)
    warp = obj.group_warp_counts().astype(float)
    obj.close()
    per_group = emit_module(gen)[1]["per_group"]
    bytes_g = np.array([8.0 * (g["loads"] + g["stores"]) for g in per_group])
    cost_lane = inst * bytes_g                 # bytes issued for lanes that needed the group
    cost_warp = warp * 32 * bytes_g            # bytes-equivalent of the warp time spent
    print(f"{args.workload} driver={args.driver or 'default'} B={args.batch}: kernel {sec * 1e3:.1f} ms, "
          f"newton steps {int(it.sum())}, mean iter {it.mean():.1f}, mean warp-max "
          f"{it[:len(it) // 32 * 32].reshape(-1, 32).max(1).mean():.1f}")
    print(f"overall lane occupancy (byte-weighted) {cost_lane.sum() / cost_warp.sum():.3f}")
    order = np.argsort(-cost_warp)[:12]
    print("group  bytes/exec   inst-exec/instance  warp-exec/warp  occupancy  share of warp time")
    nw = args.batch / 32
    for g in order:
        print(f"{g:5d} {bytes_g[g]:11.0f} {inst[g] / args.batch:19.2f} {warp[g] / nw:15.2f} "
              f"{inst[g] / max(warp[g] * 32, 1):10.3f} {cost_warp[g] / cost_warp.sum():10.3f}")


if __name__ == "__main__":
    main()
