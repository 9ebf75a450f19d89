#!/usr/bin/env python
"""Level-set profile of the dependency DAG of the largest dependency groups (the KKT factorisation and the
triangular solves): critical path and width per level.  Answers whether ONE instance can keep a warp or a CTA
busy (north_star's "each warp or CTA owns one or a few instances"): with W lanes working on one instance a
group takes at least  sum over levels of ceil(width/W) * (cost of the level's longest instruction)  steps, and at
least the critical path however many lanes help.

    python tools/level_profile.py [workload ...]        -> profiles/r02_level_profile.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def profile(name, top=4):
    from tenscalc_b200 import problems, itable as it
    gen, _ = problems.BUILDERS[name]()
    pr = gen.problem
    tab = pr.tab
    out = [f"{name}: {len(pr.groups)} groups, {sum(len(g) for g in pr.groups)} instructions, KKT {gen.info['kkt']}, nnz(L) {gen.info['nnzL']}"]
    for g in sorted(range(len(pr.groups)), key=lambda g: -len(pr.groups[g]))[:top]:
        members = pr.groups[g]
        pos = {i: k for k, i in enumerate(members)}
        level = np.zeros(len(members), dtype=np.int64)          # longest path in instructions
        depth = np.zeros(len(members), dtype=np.int64)          # longest path in multiply-adds (terms), +20 per division
        cost = np.zeros(len(members), dtype=np.int64)
        for k, i in enumerate(members):
            t, ops = tab.types[i], tab.operands[i]
            c = max(1, len(ops) // 2)
            if t in (it.I_div, it.I_plus_minus_dot_div, it.I_minus_dot_div, it.I_inv):
                c += 20                                          # FP64 division ~ 20 dependent FMA slots
            cost[k] = c
            lv = dp = 0
            for o in ops:
                q = pos.get(o)
                if q is not None:
                    lv = max(lv, level[q] + 1)
                    dp = max(dp, depth[q])
            level[k] = lv
            depth[k] = dp + c
        nlev = int(level.max()) + 1
        width = np.bincount(level, minlength=nlev)
        work = int(cost.sum())
        lvl_cost = np.zeros(nlev, dtype=np.int64)
        np.maximum.at(lvl_cost, level, cost)
        line = (f"  group {g:3d}: {len(members):6d} instr, work {work:7d} op-slots, {nlev:5d} levels, critical path "
                f"{int(depth.max()):6d} op-slots (parallelism work/cp = {work / depth.max():5.1f}), width mean {width.mean():6.1f} "
                f"median {int(np.median(width))} max {int(width.max())}")
        out.append(line)
        for W in (32, 128, 256):
            steps = int((np.ceil(width / W) * lvl_cost).sum())
            out.append(f"             {W:3d} lanes on one instance, level by level: >= {steps:7d} op-slots per execution "
                       f"(speed-up over one lane {work / steps:4.1f}x of {W}x)")
    return out


def main():
    names = sys.argv[1:] or ["mpc_quadcopter", "l1l2"]
    lines = []
    for n in names:
        lines += profile(n)
    text = "\n".join(lines)
    print(text)
    os.makedirs("profiles", exist_ok=True)
    with open(os.path.join("profiles", "r02_level_profile.txt"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
